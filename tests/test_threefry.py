"""CPU: the head draw of `best_action` — jax.random.randint restated (isdqn_threefry_randint, host C) against the Python
restatement in oracle/threefry_oracle.py, and the Threefry-2x32 block function against the Random123 known answers."""
import ctypes as C

import numpy as np

from isdqn_b200 import _lib
from oracle import threefry_oracle as T

# Salmon et al. (Random123) known-answer tests for threefry2x32_20; the same three vectors are in jax's own test-suite
KAT = [
    ((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6B200159, 0x99BA4EFE)),
    ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
    ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0)),
]


def test_threefry2x32_known_answers():
    lib = _lib.load()
    for key, ctr, want in KAT:
        assert T.threefry2x32(*key, *ctr) == want
        out = (C.c_uint32 * 2)()
        lib.isdqn_threefry2x32(key[0], key[1], ctr[0], ctr[1], out)
        assert (out[0], out[1]) == want


def test_randint_matches_the_restatement_and_is_uniform():
    lib = _lib.load()
    rng = np.random.default_rng(0)
    counts = np.zeros(9, dtype=np.int64)
    for _ in range(3000):
        k0, k1 = (int(x) for x in rng.integers(0, 2**32, 2, dtype=np.uint64))
        for span in (1, 2, 9, 49, 1000):
            got = lib.isdqn_threefry_randint(k0, k1, 0, span)
            assert got == T.randint((k0, k1), 0, span) and 0 <= got < span
        counts[lib.isdqn_threefry_randint(k0, k1, 0, 9)] += 1
    assert counts.min() > 3000 / 9 * 0.75  # every head is drawn
    assert lib.isdqn_threefry_randint(1, 2, -5, -2) == T.randint((1, 2), -5, -2)


def test_raw_key_forms():
    from isdqn_b200.networks.isdqn import _raw_key

    assert _raw_key(7) == (0, 7)  # jax.random.PRNGKey(7) with x64 disabled
    assert _raw_key(np.array([3, 4], dtype=np.uint32)) == (3, 4)
    assert _raw_key(np.array([[3, 4]], dtype=np.uint32)) == (3, 4)
    a, b = _raw_key("seed"), _raw_key("seed")
    assert a == b and a != _raw_key("other")


def test_split_and_uniform_match_the_restatement():
    lib = _lib.load()
    rng = np.random.default_rng(1)
    for _ in range(500):
        k0, k1 = (int(x) for x in rng.integers(0, 2**32, 2, dtype=np.uint64))
        for num in (2, 3, 5):
            out = (C.c_uint32 * (2 * num))()
            lib.isdqn_threefry_split(k0, k1, num, out)
            assert [(out[2 * i], out[2 * i + 1]) for i in range(num)] == T.split((k0, k1), num)
        u = lib.isdqn_threefry_uniform(k0, k1)
        assert u == np.float32(T.uniform((k0, k1))) and 0.0 <= u < 1.0


def test_select_action_follows_the_reference_control_flow():
    """utils.py:8-15 on the host: explore iff uniform(k_u) <= eps, random action from k_a, greedy call with k_g."""
    from isdqn_b200.sample_collection import utils as U

    seen = []

    def best_action(params, state, key):
        seen.append(tuple(int(x) for x in key))
        return np.int32(7)

    for seed in range(200):
        ku, ka, kg = T.split((0, seed), 3)
        eps = 0.5
        got = U.select_action(best_action, None, None, seed, 9, lambda n: eps, 0)
        if T.uniform(ku) <= eps:
            assert int(got) == T.randint(ka, 0, 9)
        else:
            assert int(got) == 7 and seen[-1] == kg
    sched = U.linear_schedule(1.0, 0.01, 1000)
    assert sched(0) == 1.0 and abs(sched(500) - 0.505) < 1e-12 and abs(sched(5000) - 0.01) < 1e-12
