"""ORACLE (test infrastructure): pure-Python restatement of the NumPy `Generator(PCG64)` draws that
`slimdqn/sample_collection/samplers.py:17,43,110` consume.

The algorithm lives in a third-party dependency of the reference (numpy==2.2.1, setup.cfg:22; this image has
numpy 2.3.5 — same algorithm).  Restated from numpy's published sources
(numpy/random/src/pcg64/pcg64.h: XSL-RR 128/64; numpy/random/src/distributions/distributions.c:
`buffered_bounded_lemire_uint32`, `random_standard_uniform` = (next64 >> 11) * 2**-53,
`random_uniform` = low + range * u) and PINNED against numpy itself in `tests/test_pcg64.py`
(numpy travels to the GPU box, so this pin is re-checked everywhere).
"""
from __future__ import annotations

import numpy as np

MASK64 = (1 << 64) - 1
MASK128 = (1 << 128) - 1
PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645  # PCG_DEFAULT_MULTIPLIER_128


class PCG64Oracle:
    """State mirror of `np.random.PCG64`: (state, inc, has_uint32, uinteger)."""

    def __init__(self, state: int, inc: int, has_uint32: int = 0, uinteger: int = 0) -> None:
        self.state = state & MASK128
        self.inc = inc & MASK128
        self.has_uint32 = int(has_uint32)
        self.uinteger = int(uinteger)

    @classmethod
    def from_seed(cls, seed) -> "PCG64Oracle":
        # SeedSequence hashing stays NumPy's: only the stream arithmetic is restated.
        st = np.random.PCG64(seed).state
        return cls(st["state"]["state"], st["state"]["inc"], st["has_uint32"], st["uinteger"])

    def numpy_state(self) -> dict:
        return {
            "bit_generator": "PCG64",
            "state": {"state": self.state, "inc": self.inc},
            "has_uint32": self.has_uint32,
            "uinteger": self.uinteger,
        }

    # -- raw streams -------------------------------------------------------------------------------
    def next64(self) -> int:
        self.state = (self.state * PCG_MULT + self.inc) & MASK128
        hi, lo = self.state >> 64, self.state & MASK64
        x = hi ^ lo
        rot = self.state >> 122
        return ((x >> rot) | (x << ((-rot) & 63))) & MASK64

    def next32(self) -> int:
        if self.has_uint32:
            self.has_uint32 = 0
            return self.uinteger
        v = self.next64()
        self.has_uint32 = 1
        self.uinteger = v >> 32
        return v & 0xFFFFFFFF

    def next_double(self) -> float:
        return (self.next64() >> 11) * (1.0 / 9007199254740992.0)

    # -- Generator methods used by the samplers ---------------------------------------------------------
    def integers(self, n: int, size: int) -> np.ndarray:
        """`Generator.integers(n, size=size)` (int64 out) for 1 <= n <= 2**32."""
        out = np.empty(size, dtype=np.int64)
        rng = n - 1
        if rng == 0:
            out[:] = 0  # consumes no randomness
            return out
        if rng == 0xFFFFFFFF:
            for i in range(size):
                out[i] = self.next32()
            return out
        rng_excl = rng + 1
        for i in range(size):
            m = self.next32() * rng_excl
            leftover = m & 0xFFFFFFFF
            if leftover < rng_excl:
                threshold = (0x100000000 - rng_excl) % rng_excl
                while leftover < threshold:
                    m = self.next32() * rng_excl
                    leftover = m & 0xFFFFFFFF
            out[i] = m >> 32
        return out

    def uniform(self, low: float, high: float, size: int) -> np.ndarray:
        """`Generator.uniform(low, high, size)`: low + (high-low) * next_double, one next64 per element."""
        rng = np.float64(high) - np.float64(low)
        out = np.empty(size, dtype=np.float64)
        for i in range(size):
            out[i] = np.float64(low) + rng * np.float64(self.next_double())
        return out
