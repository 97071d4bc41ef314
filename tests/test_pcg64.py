"""CPU: the PCG64 / Lemire / uniform restatement against numpy's Generator (the third-party code the
reference's samplers call: samplers.py:17,43,110)."""
import numpy as np
import pytest

from oracle.pcg64_oracle import PCG64Oracle


@pytest.mark.parametrize("seed", [0, 1, 12345, 2**40 + 7])
def test_raw_streams(seed):
    g = np.random.Generator(np.random.PCG64(seed))
    o = PCG64Oracle.from_seed(seed)
    want = g.bit_generator.random_raw(64)
    got = np.array([o.next64() for _ in range(64)], dtype=np.uint64)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("n", [1, 2, 3, 10, 37, 1000, 999_983, 1_000_000, 2**31 - 1, 2**31 + 11, 2**32 - 5, 2**32 - 1])
def test_integers_bit_exact(n):
    g = np.random.default_rng(7)
    o = PCG64Oracle.from_seed(7)
    for size in (1, 5, 32, 33, 1000):
        np.testing.assert_array_equal(o.integers(n, size), g.integers(n, size=size))
    assert o.numpy_state() == g.bit_generator.state


def test_integers_rejection_path_is_exercised():
    # n just above 2**31: ~half of the draws are rejected, so the redraw loop and the buffered halves matter
    n = 2**31 + 11
    g = np.random.default_rng(3)
    o = PCG64Oracle.from_seed(3)
    before = o.state
    got = o.integers(n, 4096)
    np.testing.assert_array_equal(got, g.integers(n, size=4096))
    assert o.numpy_state() == g.bit_generator.state
    assert o.state != before


def test_uniform_bit_exact_and_interleaving():
    g = np.random.default_rng(11)
    o = PCG64Oracle.from_seed(11)
    for root in (1.0, 2.5, 924742.82967363915, 1e-7):
        np.testing.assert_array_equal(o.uniform(0.0, root, 257), g.uniform(0.0, root, size=257))
        # interleave with 32-bit draws: uniform must not disturb the buffered half
        np.testing.assert_array_equal(o.integers(1000, 3), g.integers(1000, size=3))
    assert o.numpy_state() == g.bit_generator.state
