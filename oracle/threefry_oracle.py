"""ORACLE (test infrastructure): jax 0.4.30's threefry PRNG restated in plain Python integers, for the head draw of
`iSDQN.best_action` — `jax.random.randint(key, (), 0, K)` (slimdqn/networks/isdqn.py:127-135).

JAX is a pinned dependency of the reference (setup.cfg:20-21) that is absent from this image and from /root/reference, so
this follows its published algorithm:
  threefry2x32     jax/_src/prng.py `_threefry2x32_lowering` / Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"
                   (Threefry-2x32, 20 rounds) — pinned by the Random123 known-answer vectors in tests/test_threefry.py
  split            jax/_src/prng.py `threefry_split`: threefry_2x32(key, iota(2 * num)) reshaped (num, 2)
  random_bits      jax/_src/prng.py `threefry_random_bits` (jax_threefry_partitionable = False, the 0.4.30 default)
  randint          jax/_src/random.py `_randint`: two 32-bit draws from the two halves of a split, combined multiply-shift
PARITY UNPINNED for split / random_bits / randint (no JAX here to generate vectors); the block function is pinned.
"""
from __future__ import annotations

M = 0xFFFFFFFF


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (32 - r))) & M


def threefry2x32(k0: int, k1: int, x0: int, x1: int):
    rot = ((13, 15, 26, 6), (17, 29, 16, 24))
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0 = (x0 + ks[0]) & M
    x1 = (x1 + ks[1]) & M
    for g in range(5):
        for r in rot[g & 1]:
            x0 = (x0 + x1) & M
            x1 = _rotl(x1, r)
            x1 ^= x0
        x0 = (x0 + ks[(g + 1) % 3]) & M
        x1 = (x1 + ks[(g + 2) % 3] + g + 1) & M
    return x0, x1


def threefry_2x32_counts(key, counts):
    """jax's threefry_2x32(keypair, count): the flattened counts (padded to an even length) are split into two halves that
    form the two lanes; the outputs of the two lanes are concatenated."""
    n = len(counts)
    c = list(counts) + ([0] if n % 2 else [])
    half = len(c) // 2
    out0, out1 = [], []
    for i in range(half):
        y0, y1 = threefry2x32(key[0], key[1], c[i], c[half + i])
        out0.append(y0)
        out1.append(y1)
    return (out0 + out1)[:n]


def split(key, num: int = 2):
    w = threefry_2x32_counts(key, list(range(2 * num)))
    return [(w[2 * i], w[2 * i + 1]) for i in range(num)]


def random_bits32(key) -> int:
    return threefry_2x32_counts(key, [0])[0]


def randint(key, minval: int, maxval: int) -> int:
    k1, k2 = split(key, 2)
    higher, lower = random_bits32(k1), random_bits32(k2)
    span = (maxval - minval) & M
    mult = (2**16) % span
    mult = ((mult * mult) & M) % span
    off = ((((higher % span) * mult) & M) + (lower % span)) & M
    return minval + off % span


def uniform(key) -> float:
    """jax.random.uniform(key) for float32: (bits >> 9 | 0x3F800000) viewed as a float in [1, 2), minus 1."""
    import struct

    bits = (random_bits32(key) >> 9) | 0x3F800000
    return struct.unpack("<f", struct.pack("<I", bits))[0] - 1.0
