"""GPU: device samplers against numpy's Generator (uniform) and the oracle (prioritized): draws, keys and the RNG
state afterwards are bit-exact; reference tests/test_samplers.py restated."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def samplers():
    from isdqn_b200.sample_collection import samplers as s

    return s


@pytest.mark.parametrize("n", [1, 2, 3, 10, 37, 1000, 999_983, 1_000_000])
def test_uniform_draws_match_numpy(samplers, n):
    s = samplers.UniformSamplingDistribution(seed=7)
    g = np.random.default_rng(7)
    for k in range(n):
        s.add(k * 3 + 1)  # key != index
    for size in (1, 5, 32, 33, 1000, 1025, 5000):
        want = g.integers(n, size=size) * 3 + 1
        got = s.sample(size)
        assert got.dtype == np.int32
        np.testing.assert_array_equal(got, want.astype(np.int32))
        assert s._rng_key.bit_generator.state == g.bit_generator.state


def test_uniform_rejection_heavy(samplers):
    """Drives the kernel directly with n just above 2**30 (25 % rejections) to exercise the scan/compaction."""
    import torch

    from isdqn_b200 import _lib

    lib = _lib.load()
    s = samplers.UniformSamplingDistribution(seed=3)
    g = np.random.default_rng(3)
    for n in (2**30 + 12345, 2**31 - 1, 3 * 2**29 + 1):
        for size in (1, 31, 4096, 70_000):
            out = torch.empty(size, dtype=torch.int32, device="cuda")
            _lib.check(lib.isdqn_sample_uniform(s._d_rng.data_ptr(), n, size, None, 1, out.data_ptr(), None, None, _lib.stream_ptr()))
            np.testing.assert_array_equal(out.cpu().numpy().astype(np.int64), g.integers(n, size=size))
            s._pull_rng_state()
            assert s._rng_key.bit_generator.state == g.bit_generator.state


def test_uniform_many_cta_variant(samplers):
    """isdqn_sample_uniform_ws (two passes over the draw positions, whole GPU) gives numpy's draws and generator state:
    large draws with few rejections (the fast path), 25-50 % rejections (more than the margin: the single-CTA fallback
    runs), and a moderate rejection rate that still fits the margin."""
    import torch

    from isdqn_b200 import _lib

    lib = _lib.load()
    s = samplers.UniformSamplingDistribution(seed=11)
    g = np.random.default_rng(11)
    ws = torch.zeros(int(lib.isdqn_sample_uniform_workspace_bytes()), dtype=torch.uint8, device="cuda")
    cases = [(1_000_000, 65_536), (999_983, 8192), (1_048_577, 200_001), (37, 65_536), (2**30 + 12345, 70_000),
             (2**31 - 1, 9000), (2**32 // 3 + 1, 100_000), (1_000_000, 1_048_576), (2**26 + 5, 65_536)]
    for n, size in cases:
        out = torch.empty(size, dtype=torch.int32, device="cuda")
        _lib.check(lib.isdqn_sample_uniform_ws(s._d_rng.data_ptr(), n, size, None, 1, out.data_ptr(), None, None, ws.data_ptr(),
                                               ws.numel(), _lib.stream_ptr()))
        np.testing.assert_array_equal(out.cpu().numpy().astype(np.int64), g.integers(n, size=size), err_msg=f"n={n} size={size}")
        s._pull_rng_state()
        assert s._rng_key.bit_generator.state == g.bit_generator.state, (n, size)
    # through the sampler class (keys + slots) at the throughput shape
    s2 = samplers.UniformSamplingDistribution(seed=5)
    g2 = np.random.default_rng(5)
    for k in range(50_000):
        s2.add(k * 2 + 7)
    for size in (65_536, 32, 10_000):
        want = g2.integers(50_000, size=size) * 2 + 7
        d_index, d_key, d_slot = s2.sample_device(size, 50_001)
        np.testing.assert_array_equal(d_key.cpu().numpy(), want.astype(np.int32))
        np.testing.assert_array_equal(d_slot.cpu().numpy(), (want % 50_001).astype(np.int32))
        s2._pull_rng_state()
        assert s2._rng_key.bit_generator.state == g2.bit_generator.state


def test_swap_remove_key_maps(samplers):
    s = samplers.UniformSamplingDistribution(seed=0)
    g = np.random.default_rng(0)
    for k in range(10):
        s.add(k)
    s.sample(4), g.integers(10, size=4)
    s.add(10)
    s.remove(0)
    assert 0 not in s._key_to_index and s._index_to_key[0] == 10 and s._key_to_index[10] == 0
    idx = g.integers(10, size=64)
    np.testing.assert_array_equal(s.sample(64), np.asarray([s._index_to_key[i] for i in idx], dtype=np.int32))


def test_prioritized_reference_test(samplers):  # reference tests/test_samplers.py:16-35
    sampler = samplers.PrioritizedSamplingDistribution(seed=0, max_capacity=10)
    for key, priority in zip([0, 1, 2, 3, 4], [1.0, 2.0, 3.0, 4.0, 0.0]):
        sampler.add(key, priority=priority)
    np.testing.assert_array_less(sampler.sample(5), 4)
    sampler.update(keys=np.array([2, 3]), priorities=np.array([0.0, 0.0]))
    np.testing.assert_array_less(sampler.sample(5), 2)
    sampler.remove(0)
    np.testing.assert_array_almost_equal(sampler.sample(5), 1)


def test_prioritized_matches_oracle(samplers):
    from oracle.samplers_oracle import PrioritizedSamplingOracle

    cap = 1000
    a = samplers.PrioritizedSamplingDistribution(seed=5, max_capacity=cap, priority_exponent=0.6)
    b = PrioritizedSamplingOracle(5, cap, 0.6)
    rng = np.random.default_rng(1)
    key = 0
    for step in range(1500):
        p = float(abs(rng.standard_normal())) if rng.random() > 0.05 else 0.0
        a.add(key, priority=p)
        b.add(key, priority=p)
        key += 1
        if key > cap - 10:
            a.remove(key - (cap - 10))
            b.remove(key - (cap - 10))
        if step % 97 == 96:
            ka, kb = a.sample(64), b.sample(64)
            np.testing.assert_array_equal(ka, kb)
            newp = np.abs(rng.standard_normal(64))
            a.update(ka, newp)
            b.update(kb, newp)
    assert a._sum_tree._nodes.tobytes() == b.tree._nodes.tobytes()
    assert a._index_to_key == b.index_to_key
    assert a._rng_key.bit_generator.state == b.rng.numpy_state()
    a._flush_maps()  # sample()/sample_device() do this; _draw_device is the raw launch
    d_idx, d_key, d_slot, d_t = a._draw_device(4096, 77, want_targets=True)
    want_t = b.rng.uniform(0.0, b.tree.root, 4096)
    np.testing.assert_array_equal(d_t.cpu().numpy(), want_t)
    want_idx = b.tree.query(want_t)
    np.testing.assert_array_equal(d_idx.cpu().numpy(), want_idx)
    np.testing.assert_array_equal(d_key.cpu().numpy(), np.asarray([b.index_to_key[i] for i in want_idx], dtype=np.int32))
    np.testing.assert_array_equal(d_slot.cpu().numpy(), d_key.cpu().numpy() % 77)


def test_prioritized_empty_tree_mirrors_reference_bug(samplers):
    s = samplers.PrioritizedSamplingDistribution(seed=0, max_capacity=10)
    s.add(0, priority=0.0)
    with pytest.raises(AttributeError):  # samplers.py:105-108 (`.keys` on an ndarray)
        s.sample(3)
