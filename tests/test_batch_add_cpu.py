"""CPU: the batched forms of `ReplayBuffer.add` — the closed-form accumulator run and the sampler run — against the
one-at-a-time code paths they replace (which the golden fixtures pin against the unmodified reference)."""
import numpy as np
import pytest

from isdqn_b200.sample_collection.accumulator import NStepAccumulator


class _Committer:
    def __init__(self):
        self.next = 0

    def __call__(self, fr):
        fr.frame_id = self.next
        self.next += 1
        return fr.frame_id


def _drive_single(S, n, gamma, acts, rews, terms, ends):
    c = _Committer()
    acc = NStepAccumulator(S, n, gamma, c)
    out = []
    for i in range(len(acts)):
        for rec in acc.accumulate(("obs", i), acts[i], rews[i], bool(terms[i]), bool(ends[i])):
            out.append((rec[0].copy(), rec[1], rec[2], rec[3]))
    return out, acc, c


def _drive_batched(S, n, gamma, acts, rews, terms, ends):
    c = _Committer()
    acc = NStepAccumulator(S, n, gamma, c)
    out = []
    i, N = 0, len(acts)
    n_runs = 0
    while i < N:
        m = acc.steady_run(terms, ends, i)
        if m == 0:
            for rec in acc.accumulate(("obs", i), acts[i], rews[i], bool(terms[i]), bool(ends[i])):
                out.append((rec[0].copy(), rec[1], rec[2], rec[3]))
            i += 1
            continue
        n_runs += 1
        first = c.next
        c.next += m  # what ReplayBuffer._commit_many does
        refs, a, r = acc.accumulate_run(first, acts[i : i + m], rews[i : i + m], bool(ends[i + m - 1]))
        for j in range(m):
            out.append((refs[j].copy(), a[j], r[j], False))
        i += m
    return out, acc, c, n_runs


@pytest.mark.parametrize("S,n,gamma", [(4, 1, 0.99), (4, 3, 0.5), (1, 1, 0.99), (2, 2, 0.97), (3, 5, 0.9)])
def test_accumulate_run_equals_per_transition(S, n, gamma):
    rng = np.random.default_rng(S * 100 + n)
    N = 3000
    acts = rng.integers(0, 9, N)
    rews = rng.integers(-1, 2, N).astype(np.float64) * rng.random(N)
    u = rng.random(N)
    terms = u < 0.01
    ends = terms | (u < 0.02)
    a, acc_a, ca = _drive_single(S, n, gamma, acts, rews, terms, ends)
    b, acc_b, cb, n_runs = _drive_batched(S, n, gamma, acts, rews, terms, ends)
    assert n_runs > 5
    assert len(a) == len(b) and ca.next == cb.next
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x[0], y[0])
        assert int(x[1]) == int(y[1])
        assert np.float64(x[2]).tobytes() == np.float64(y[2]).tobytes()  # bit-exact n-step return
        assert bool(x[3]) == bool(y[3])
    # the trajectory the next transition will see is the same
    assert [f.frame_id for f in acc_a.trajectory] == [f.frame_id for f in acc_b.trajectory]
    assert [int(f.action) for f in acc_a.trajectory] == [int(f.action) for f in acc_b.trajectory]
    assert [float(f.reward) for f in acc_a.trajectory] == [float(f.reward) for f in acc_b.trajectory]


def test_accumulate_run_int_rewards_like_the_reference_tests():
    N = 200
    acts = np.arange(N)
    rews = np.arange(N)  # Python/NumPy ints: `reward * gamma ** k` promotes to float in both paths
    terms = np.zeros(N, bool)
    ends = np.zeros(N, bool)
    a, _, _ = _drive_single(4, 3, 0.5, acts, rews, terms, ends)
    b, _, _, _ = _drive_batched(4, 3, 0.5, acts, rews, terms, ends)
    assert [float(x[2]) for x in a] == [float(y[2]) for y in b]


# --------------------------------------------------------------------------------------------- sampler runs
class _FakeTree:
    def __init__(self, capacity):
        self._depth = int(np.ceil(np.log2(capacity))) + 1
        self._first_leaf_offset = 2 ** (self._depth - 1) - 1

        class _N:
            def __init__(s, n):
                s.n = n

            def numel(s):
                return s.n

        self._d_nodes = _N(2**self._depth - 1)
        self.ops = []

    def set(self, indices, values):
        if isinstance(indices, (int, np.integer)):
            indices = np.asarray([indices], np.int32)
        if isinstance(values, (int, float, np.floating)):
            values = np.asarray([values], np.float64)
        self._enqueue(np.asarray(indices, np.int32), np.asarray(values, np.float64))

    def _enqueue(self, idx, val):
        self.ops.append((idx.tolist(), val.tolist()))

    def _enqueue_ops(self, idx, val, lens):
        o = 0
        for m in lens.tolist():
            self.ops.append((idx[o : o + m].tolist(), val[o : o + m].tolist()))
            o += m


def _singles(ops):
    """Ops with the merged fill sets (distinct ascending leaves, no leaf tag) split back into single sets: node by node
    np.add.at performs the same additions in the same order either way."""
    out = []
    for idx, val in ops:
        if len(idx) >= 2 and all(v > -1.0 for v in val) and all(a < b for a, b in zip(idx, idx[1:])):
            out.extend(([i], [v]) for i, v in zip(idx, val))
        else:
            out.append((idx, val))
    return out


def _bare_sampler(cls, capacity, exponent=1.0):
    s = object.__new__(cls)
    s._key_to_index, s._index_to_key, s._patches = {}, [], {}
    s._max_capacity, s._priority_exponent = capacity, exponent
    s._sum_tree = _FakeTree(capacity)
    return s


@pytest.mark.parametrize("capacity,exponent", [(7, 1.0), (37, 0.6), (3, 1.0), (5, 1.0)])
def test_sampler_run_equals_add_remove_calls(capacity, exponent):
    from isdqn_b200.sample_collection.samplers import PrioritizedSamplingDistribution, UniformSamplingDistribution

    rng = np.random.default_rng(capacity)
    for cls in (UniformSamplingDistribution, PrioritizedSamplingDistribution):
        one = _bare_sampler(cls, capacity, exponent)
        run = _bare_sampler(cls, capacity, exponent)
        add_count = oldest = 0
        for _ in range(12):
            m = int(rng.integers(1, 2 * capacity + 3))
            m = min(m, capacity)
            prios = np.abs(rng.standard_normal(m)) + 1e-3
            prios[rng.random(m) < 0.2] = 0.0
            # one at a time (the order of ReplayBuffer.add, replay_buffer.py:190-196)
            a, o = add_count, oldest
            for j in range(m):
                if cls is PrioritizedSamplingDistribution:
                    one.add(a, float(prios[j]))
                else:
                    one.add(a)
                a += 1
                if a > capacity:
                    one.remove(o)
                    o += 1
            evict_from = min(max(capacity - add_count, 0), m)
            run._add_remove_run(add_count, m, oldest, evict_from, prios.tolist() if cls is PrioritizedSamplingDistribution else None)
            add_count, oldest = a, o
            assert one._index_to_key == run._index_to_key
            assert one._key_to_index == run._key_to_index
            assert one._patches == run._patches
            if cls is PrioritizedSamplingDistribution:
                assert _singles(one._sum_tree.ops) == _singles(run._sum_tree.ops)
                assert all(len(i) <= 1024 for i, _ in run._sum_tree.ops)
                one._sum_tree.ops.clear()
                run._sum_tree.ops.clear()


def test_sampler_run_max_priority_tag():
    from isdqn_b200 import _lib
    from isdqn_b200.sample_collection.samplers import PrioritizedSamplingDistribution

    s = _bare_sampler(PrioritizedSamplingDistribution, 5)
    s._add_remove_run(0, 3, 0, 3, "max")
    assert _singles(s._sum_tree.ops) == [([0], [_lib.SUMTREE_TAG_MAX]), ([1], [_lib.SUMTREE_TAG_MAX]), ([2], [_lib.SUMTREE_TAG_MAX])]
