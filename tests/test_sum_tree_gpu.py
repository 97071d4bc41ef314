"""GPU: the CUDA sum tree through the reference-shaped API: the reference's own known-answer tests
(reference tests/test_sum_tree.py), the golden traces (bit-exact float64 heap), randomized checks vs the oracle."""
import os

import numpy as np
import pytest

from tests import scenarios as S

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def sum_tree():
    from isdqn_b200.sample_collection import sum_tree as st

    return st


def test_negative_capacity_raises(sum_tree):
    with pytest.raises(AssertionError):
        sum_tree.SumTree(capacity=-1)


def test_negative_value_raises(sum_tree):
    with pytest.raises(AssertionError):
        sum_tree.SumTree(capacity=100).set(0, -1)


def test_set_small_capacity(sum_tree):
    tree = sum_tree.SumTree(capacity=1)
    tree.set(0, 1.5)
    assert tree.root == 1.5


def test_set_and_get_value(sum_tree):
    tree = sum_tree.SumTree(capacity=100)
    tree.set(0, 1.0)
    assert tree.get(0) == 1.0
    leaf_index = tree._first_leaf_offset
    nodes = tree._nodes
    while leaf_index > 0:
        leaf_index = leaf_index // 2
        assert nodes[leaf_index] == 1.0


def test_set_and_get_values_vectorized(sum_tree):
    tree = sum_tree.SumTree(capacity=100)
    tree.set(np.array([1, 2], dtype=np.int32), np.array([3.0, 4.0], dtype=np.float32))
    assert tree.get(1) == 3.0 and tree.get(2) == 4.0 and tree.root == 7.0


def test_set_with_duplicates(sum_tree):
    tree = sum_tree.SumTree(capacity=100)
    tree.set(np.array([1, 1, 1, 2, 2], dtype=np.int32), np.array([3.0, 3.0, 3.0, 4.0, 4.0], dtype=np.float32))
    assert tree.get(1) == 3.0 and tree.get(2) == 4.0 and tree.root == 7.0


def test_capacity_greater_than_requested(sum_tree):
    assert sum_tree.SumTree(capacity=100)._nodes.size >= 100


def test_query_empty_tree(sum_tree):
    with pytest.raises(ValueError):
        sum_tree.SumTree(capacity=100).query(1.0)


def test_query_value(sum_tree):
    tree = sum_tree.SumTree(capacity=100)
    tree.set(5, 1.0)
    assert tree.query(0.99) == 5


def test_query_values_vectorized_and_update(sum_tree):
    tree = sum_tree.SumTree(capacity=4)
    tree.set(np.array([0, 1, 2, 3], dtype=np.int32), np.array([0.5, 1.0, 0.5, 0.5], dtype=np.float32))
    assert tree.root == 2.5 and tree._depth == 3 and tree._nodes.size == 7
    out = tree.query(np.array([1.5, 1.0]))
    assert out.dtype == np.int32
    np.testing.assert_array_equal(out, np.array([2, 1], np.int32))
    tree.set(0, 0.25)
    assert tree.root == 2.25
    assert tree.query(0.249) == 0 and tree.query(0.5) == 1 and tree.query(1.25) == 2


def test_query_values_vectorized_large_tree(sum_tree):
    tree = sum_tree.SumTree(capacity=8)
    tree.set(np.arange(8, dtype=np.int32), np.ones((8,), dtype=np.float32))
    assert tree.root == 8.0 and tree._depth == 4 and tree._nodes.size == 15
    np.testing.assert_array_equal(tree.query(np.arange(8, dtype=np.int32)), np.arange(8, dtype=np.int32))


def test_max_recorded_priority(sum_tree):
    tree = sum_tree.SumTree(capacity=100)
    tree.set(0, 0)
    assert tree.max_recorded_priority == 1
    for i in range(1, 32):
        tree.set(i, i)
        assert tree.max_recorded_priority == i


def test_power_of_two_capacity_overflow_is_an_index_error(sum_tree):
    # leaf index == capacity only exists when the leaves were padded
    tree = sum_tree.SumTree(capacity=8)
    with pytest.raises(IndexError):
        tree.set(8, 1.0)


@pytest.mark.parametrize("tt", S.TREE_TRACES, ids=lambda t: t.name)
def test_golden_traces_bit_exact(sum_tree, tt):
    want = dict(np.load(os.path.join(GOLDEN, f"sumtree_{tt.name}.npz")))
    got = S.run_tree_trace(tt, sum_tree.SumTree(tt.capacity))
    S.compare_results(got, want, where=tt.name)


def test_large_set_and_large_query_vs_oracle(sum_tree):
    from oracle.sum_tree_oracle import SumTreeOracle

    cap = 50_000
    rng = np.random.default_rng(99)
    a, b = sum_tree.SumTree(cap), SumTreeOracle(cap)
    for m in (8192, 5000, 1025, 1024, 33):
        idx = rng.integers(0, cap, m).astype(np.int32)
        idx[rng.integers(0, m, m // 4)] = idx[0]
        val = np.abs(rng.standard_normal(m)) * 10.0
        a.set(idx, val)
        b.set(idx, val)
        assert a._nodes.tobytes() == b._nodes.tobytes(), f"heap differs after set of {m}"
    with pytest.raises(Exception):
        a.set(np.zeros(9000, dtype=np.int32), np.ones(9000))
    targets = rng.random(200_000) * float(b.root)  # staged-top-levels path
    np.testing.assert_array_equal(a.query(targets), b.query(targets))
    targets = rng.random(100) * float(b.root)
    np.testing.assert_array_equal(a.query(targets), b.query(targets))


def test_tagged_swap_remove_matches_get_then_set(sum_tree):
    from oracle.sum_tree_oracle import SumTreeOracle

    a, b = sum_tree.SumTree(37), SumTreeOracle(37)
    rng = np.random.default_rng(5)
    for i in range(30):
        v = float(abs(rng.standard_normal()))
        a.set(i, v)
        b.set(i, v)
    for last in range(29, 20, -1):
        hole = int(rng.integers(0, last))
        a._enqueue(np.asarray([hole, last], np.int32), np.asarray([-(1.0 + last), 0.0]))
        b.set(np.asarray([hole, last], np.int32), np.asarray([b.get(last), 0.0]))
    assert a._nodes.tobytes() == b._nodes.tobytes()
    assert a.max_recorded_priority == b.max_recorded_priority


@pytest.mark.parametrize("path", ["direct", "queued"])
def test_batch_sized_sets_warp_kernel_bit_exact(sum_tree, path):
    """Sets of 3..32 entries (the batch-32 priority update) run on one warp: heap, maximum priority and first-duplicate-wins
    must equal the oracle's, through the direct launch (`_set_now`) and inside a queue of ops, mixed with the one- and
    two-entry ops of add / evict and with larger sets."""
    from oracle.sum_tree_oracle import SumTreeOracle

    cap = 5000
    a, b = sum_tree.SumTree(cap), SumTreeOracle(cap)
    rng = np.random.default_rng(9)
    a.set(7, 12345.5)  # a large recorded maximum that later, smaller sets must not lower
    b.set(7, 12345.5)
    for it in range(300):
        m = int(rng.choice([1, 2, 3, 5, 16, 31, 32, 33, 200]))
        idx = rng.integers(0, cap, m).astype(np.int32)
        if m > 2 and rng.random() < 0.6:
            idx[rng.integers(0, m, max(1, m // 3))] = idx[0]  # duplicates: the first one wins
        if rng.random() < 0.3:
            idx = (idx // 64 * 64 + rng.integers(0, 4, m)).astype(np.int32)  # shared ancestors close to the leaves
        val = np.abs(rng.standard_normal(m)) * float(rng.choice([1e-3, 1.0, 50.0]))
        val[rng.random(m) < 0.1] = 0.0
        if path == "direct" and 2 < m <= 32:
            a._set_now(idx, val)
        else:
            a.set(idx, val)
        b.set(idx, val)
        if it % 25 == 0:
            assert a._nodes.tobytes() == b._nodes.tobytes(), f"heap differs after op {it} (m = {m})"
            assert a.max_recorded_priority == b.max_recorded_priority
    assert a._nodes.tobytes() == b._nodes.tobytes()
    assert a.max_recorded_priority == b.max_recorded_priority == 12345.5
