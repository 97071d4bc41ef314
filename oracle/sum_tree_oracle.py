"""ORACLE (test infrastructure): CPU restatement of `slimdqn/sample_collection/sum_tree.py`.

Written as the *order-explicit* form the CUDA kernels implement, not as the reference's vectorised NumPy:

  set   (sum_tree.py:20-47)  leaf ids sorted ascending, first duplicate wins (np.unique(return_index)),
                             delta = value - old_leaf; for every level the nodes receive their deltas as a
                             sequential left fold in ascending-leaf order (what np.add.at does), so
                             node = ((node + d_a) + d_{a+1}) + ...   -- float64, NOT left+right.
  query (sum_tree.py:58-102) strict `<` go-left rule, `target -= left` when going right.

Pinned bit-exact against the unmodified reference by `oracle/make_golden.py` and
`tests/test_oracle_pinning.py` (golden traces in tests/golden/sumtree_trace.npz).
"""
from __future__ import annotations

import math

import numpy as np


class SumTreeOracle:
    def __init__(self, capacity: int) -> None:  # sum_tree.py:11-18
        assert capacity > 0, "Capacity to sum tree must be positive."
        self._capacity = capacity
        self._depth = int(math.ceil(math.log2(capacity))) + 1
        self._first_leaf_offset = (2 ** (self._depth - 1)) - 1
        self._nodes = np.zeros((2**self._depth) - 1, dtype=np.float64)
        self.max_recorded_priority = 1.0

    # ------------------------------------------------------------------ set
    def set(self, indices, values) -> None:
        if isinstance(indices, (int, np.integer)):
            indices = np.asarray([indices], np.int32)
        if isinstance(values, (int, float, np.floating)):
            values = np.asarray([values], np.float64)
        assert indices.shape == values.shape, "Indices and values must have the same shape."
        assert (values >= 0.0).all(), "Values must be positive."
        self.max_recorded_priority = max(self.max_recorded_priority, max(values))
        nodes = self._nodes
        off = self._first_leaf_offset
        # first occurrence of every distinct leaf, ascending leaf order
        first = {}
        for pos, leaf in enumerate(indices.tolist()):
            if leaf not in first:
                first[leaf] = pos
        leaves = sorted(first)
        vals64 = np.asarray(values, dtype=np.float64)
        deltas = [float(vals64[first[leaf]]) - float(nodes[off + leaf]) for leaf in leaves]
        ids = [off + leaf for leaf in leaves]
        for _ in range(self._depth):  # leaves, then every ancestor level up to the root
            for node, d in zip(ids, deltas):  # sequential fold, ascending-leaf order
                nodes[node] = nodes[node] + d
            ids = [(i - 1) // 2 for i in ids]

    def get(self, index):  # sum_tree.py:49-51
        return self._nodes[self._first_leaf_offset + index]

    @property
    def root(self) -> float:  # sum_tree.py:53-56
        return self._nodes[0]

    # ---------------------------------------------------------------- query
    def query(self, targets):
        if isinstance(targets, (int, float)):
            targets = np.asarray([targets], np.float64)
        targets = np.asarray(targets)
        if not ((targets >= 0) & (targets < self.root)).all():
            raise ValueError(f"Targets must be in the interval [0.0, {self.root}).")
        out = np.zeros(targets.shape, dtype=np.int32)
        flat_t = np.asarray(targets, dtype=np.float64).reshape(-1)
        flat_o = out.reshape(-1)
        nodes = self._nodes
        for j in range(flat_t.size):
            t = float(flat_t[j])
            node = 0
            while node < self._first_leaf_offset:
                assert t < nodes[node]
                left = 2 * node + 1
                ls = float(nodes[left])
                if t < ls:
                    node = left
                else:
                    node = left + 1
                    t = t - ls
            flat_o[j] = node - self._first_leaf_offset
        return out
