"""ORACLE (test infrastructure): CPU restatement of `slimdqn/sample_collection/samplers.py`.

  UniformSamplingOracle      samplers.py:13-49   dense index <-> key maps with swap-remove; `integers` draws
  PrioritizedSamplingOracle  samplers.py:52-116  + sum tree: add/update/remove/sample

Draws come from `PCG64Oracle` (pinned against numpy), tree state from `SumTreeOracle` (pinned against the
reference).  Both classes are pinned end-to-end against the unmodified reference in
`oracle/make_golden.py` / `tests/golden/replay_*.npz`.
"""
from __future__ import annotations

import numpy as np

from .pcg64_oracle import PCG64Oracle
from .sum_tree_oracle import SumTreeOracle


class UniformSamplingOracle:
    def __init__(self, seed: int) -> None:
        self.rng = PCG64Oracle.from_seed(seed)
        self.key_to_index: dict[int, int] = {}
        self.index_to_key: list[int] = []

    def add(self, key: int) -> None:  # samplers.py:22-24
        self.key_to_index[key] = len(self.index_to_key)
        self.index_to_key.append(key)

    def remove(self, key: int) -> None:  # samplers.py:26-37 (swap with last, pop)
        assert key in self.key_to_index
        index = self.key_to_index[key]
        last_key = self.index_to_key[-1]
        self.index_to_key[index] = last_key
        self.key_to_index[last_key] = index
        self.index_to_key.pop()
        del self.key_to_index[key]

    def sample_indices(self, size: int) -> np.ndarray:
        assert self.index_to_key
        return self.rng.integers(len(self.index_to_key), size)

    def sample(self, size: int) -> np.ndarray:  # samplers.py:39-49
        idx = self.sample_indices(size)
        return np.asarray([self.index_to_key[i] for i in idx], dtype=np.int32)


class PrioritizedSamplingOracle(UniformSamplingOracle):
    def __init__(self, seed: int, max_capacity: int, priority_exponent: float = 1.0) -> None:
        self.max_capacity = max_capacity
        self.priority_exponent = priority_exponent
        self.tree = SumTreeOracle(max_capacity)
        super().__init__(seed)

    def add(self, key: int, priority: float) -> None:  # samplers.py:67-74
        super().add(key)
        if priority is None:
            priority = 0.0
        self.tree.set(self.key_to_index[key], 0.0 if priority == 0.0 else priority**self.priority_exponent)

    def update(self, keys, priorities) -> None:  # samplers.py:76-88
        if not isinstance(keys, np.ndarray):
            keys = np.asarray([keys], dtype=np.int32)
        priorities = np.where(priorities == 0.0, 0.0, priorities**self.priority_exponent)
        self.tree.set(np.asarray([self.key_to_index[int(k)] for k in keys], dtype=np.int32), priorities)

    def remove(self, key: int) -> None:  # samplers.py:90-103
        index = self.key_to_index[key]
        last_index = len(self.index_to_key) - 1
        if index == last_index:
            self.tree.set(index, 0.0)
        else:
            self.tree.set(
                np.asarray([index, last_index], dtype=np.int32),
                np.asarray([self.tree.get(last_index), 0.0]),
            )
        super().remove(key)

    def sample_indices(self, size: int) -> np.ndarray:
        targets = self.rng.uniform(0.0, self.tree.root, size)
        return self.tree.query(targets)

    def sample(self, size: int) -> np.ndarray:  # samplers.py:105-116
        if self.tree.root == 0.0:
            # the reference calls `.keys` on an ndarray here (latent bug) after consuming the draws
            super().sample(size)
            raise AttributeError("'numpy.ndarray' object has no attribute 'keys'")
        idx = self.sample_indices(size)
        return np.asarray([self.index_to_key[i] for i in idx], dtype=np.int32)
