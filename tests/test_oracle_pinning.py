"""CPU: the oracle restatements against the golden fixtures produced from the UNMODIFIED reference
(oracle/make_golden.py) and against the reference's own known-answer tests (tests/test_sum_tree.py,
tests/test_samplers.py, tests/test_replay_buffer.py of the reference, restated on the oracle API)."""
import os

import numpy as np
import pytest

from oracle.replay_oracle import ReplayOracle
from oracle.samplers_oracle import PrioritizedSamplingOracle, UniformSamplingOracle
from oracle.sum_tree_oracle import SumTreeOracle
from tests import scenarios as S

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


class OracleAdapter(S.Adapter):
    def __init__(self, sc):
        if sc.sampler == "uniform":
            sampler = UniformSamplingOracle(sc.seed)
        else:
            sampler = PrioritizedSamplingOracle(sc.seed, sc.capacity, sc.priority_exponent)
        self.rb = ReplayOracle(sampler, sc.batch, sc.capacity, sc.stack, sc.horizon, sc.gamma)

    def add(self, obs, action, reward, terminal, episode_end, priority):
        kw = {} if priority is None else {"priority": priority}
        self.rb.add(obs, action, reward, terminal, episode_end, **kw)

    def add_count(self):
        return self.rb.add_count

    def sample(self):
        return tuple(self.rb.sample())

    def sample_keys(self, size):
        return self.rb.sampler.sample(size)

    def update(self, keys, priorities):
        self.rb.update(keys, priorities=priorities)

    def memory_keys(self):
        return list(self.rb.memory.keys())

    def index_to_key(self):
        return list(self.rb.sampler.index_to_key)

    def tree_nodes(self):
        return self.rb.sampler.tree._nodes.copy() if hasattr(self.rb.sampler, "tree") else None


@pytest.mark.parametrize("sc", S.SCENARIOS, ids=lambda s: s.name)
def test_replay_oracle_matches_reference_fixture(sc):
    want = dict(np.load(os.path.join(GOLDEN, f"replay_{sc.name}.npz")))
    got = S.run_scenario(sc, OracleAdapter(sc))
    S.compare_results(got, want, where=sc.name)


@pytest.mark.parametrize("tt", S.TREE_TRACES, ids=lambda t: t.name)
def test_sumtree_oracle_matches_reference_fixture(tt):
    want = dict(np.load(os.path.join(GOLDEN, f"sumtree_{tt.name}.npz")))
    got = S.run_tree_trace(tt, SumTreeOracle(tt.capacity))
    S.compare_results(got, want, where=tt.name)


# ---- the reference's known-answer tests (reference tests/test_sum_tree.py:16-136) on the oracle -------------
def test_ref_sum_tree_known_answers():
    with pytest.raises(AssertionError):
        SumTreeOracle(-1)
    tree = SumTreeOracle(100)
    with pytest.raises(AssertionError):
        tree.set(0, -1)
    t1 = SumTreeOracle(1)
    t1.set(0, 1.5)
    assert t1.root == 1.5
    tree.set(0, 1.0)
    assert tree.get(0) == 1.0
    leaf = tree._first_leaf_offset
    while leaf > 0:
        leaf //= 2
        assert tree._nodes[leaf] == 1.0
    tree = SumTreeOracle(100)
    tree.set(np.array([1, 1, 1, 2, 2], dtype=np.int32), np.array([3.0, 3.0, 3.0, 4.0, 4.0], dtype=np.float32))
    assert tree.get(1) == 3.0 and tree.get(2) == 4.0 and tree.root == 7.0
    with pytest.raises(ValueError):
        SumTreeOracle(100).query(1.0)
    tree = SumTreeOracle(100)
    tree.set(5, 1.0)
    assert tree.query(0.99) == 5
    tree = SumTreeOracle(4)
    tree.set(np.array([0, 1, 2, 3], dtype=np.int32), np.array([0.5, 1.0, 0.5, 0.5], dtype=np.float32))
    assert tree.root == 2.5 and tree._depth == 3 and tree._nodes.size == 7
    np.testing.assert_array_equal(tree.query(np.array([1.5, 1.0])), np.array([2, 1], np.int32))
    tree.set(0, 0.25)
    assert tree.root == 2.25
    assert tree.query(0.249) == 0 and tree.query(0.5) == 1 and tree.query(1.25) == 2
    tree = SumTreeOracle(8)
    tree.set(np.arange(8, dtype=np.int32), np.ones((8,), dtype=np.float32))
    assert tree.root == 8.0 and tree._depth == 4 and tree._nodes.size == 15
    np.testing.assert_array_equal(tree.query(np.arange(8, dtype=np.int32)), np.arange(8, dtype=np.int32))
    tree = SumTreeOracle(100)
    tree.set(0, 0)
    assert tree.max_recorded_priority == 1
    for i in range(1, 32):
        tree.set(i, i)
        assert tree.max_recorded_priority == i


def test_ref_prioritized_sampler_known_answers():  # reference tests/test_samplers.py:16-35
    s = PrioritizedSamplingOracle(seed=0, max_capacity=10)
    for key, p in zip([0, 1, 2, 3, 4], [1.0, 2.0, 3.0, 4.0, 0.0]):
        s.add(key, priority=p)
    assert (s.sample(5) < 4).all()
    s.update(keys=np.array([2, 3]), priorities=np.array([0.0, 0.0]))
    assert (s.sample(5) < 2).all()
    s.remove(0)
    np.testing.assert_array_almost_equal(s.sample(5), 1)


def test_ref_replay_key_mapping_known_answers():  # reference tests/test_replay_buffer.py:135-203
    cap, B = 10, 32
    rb = ReplayOracle(UniformSamplingOracle(0), B, cap, stack_size=1, update_horizon=1, gamma=0.99)
    for i in range(cap + 1):
        rb.add(np.full((84, 84), i), i, i, False, False)
    for i in range(cap):
        assert rb.sampler.key_to_index[i] == i and rb.sampler.index_to_key[i] == i
    rb.add(np.full((84, 84), cap + 1), cap + 1, cap + 1, False, False)
    assert 0 not in rb.sampler.key_to_index and rb.sampler.index_to_key[0] != 0
    assert rb.sampler.index_to_key[rb.sampler.key_to_index[cap]] == cap
    idx = np.random.default_rng(seed=0).integers(len(rb.sampler.index_to_key), size=B)
    keys = [rb.sampler.index_to_key[i] for i in idx]
    batch = rb.sample()
    for i, key in enumerate(keys):
        np.testing.assert_array_equal(batch.state[i], np.full((84, 84), key)[..., None])
        np.testing.assert_array_equal(batch.next_state[i], np.full((84, 84), key + 1)[..., None])
        assert batch.action[i] == key and batch.reward[i] == key and batch.is_terminal[i] == 0
