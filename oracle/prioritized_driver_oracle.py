"""ORACLE (test infrastructure): a prioritized training loop on the CPU, built from the pinned restatements.

The reference ships `PrioritizedSamplingDistribution` (samplers.py:52-116) and `ReplayBuffer.update`
(replay_buffer.py:215-220) but no agent or experiment uses them, and `ReplayBuffer.sample` (:198-213) returns the batch
without its keys — so there is no reference loop to copy.  This driver is the loop those pieces imply,
written against `ReplayOracle` / `PrioritizedSamplingOracle` (both pinned bit-exact against the unmodified reference):

  add      rb.add(transition, priority=tree.max_recorded_priority)            replay_buffer.py:185-196, samplers.py:67-74
  sample   targets = rng.uniform(0, root, B); idx = tree.query(targets)       samplers.py:105-116
           keys = index_to_key[idx]; batch = stack(memory[keys])              replay_buffer.py:206-212
           w = (N * leaf[idx] / root) ** -beta, normalised by its maximum     (importance weights; no reference line)
  update   rb.update(keys, priorities=p)                                      replay_buffer.py:215-220, samplers.py:76-88

Everything but the weights is byte/index/float64-tree work: the CUDA driver must match it bit for bit.
"""
from __future__ import annotations

import numpy as np

from .replay_oracle import ReplayOracle
from .samplers_oracle import PrioritizedSamplingOracle


class PrioritizedDriverOracle:
    def __init__(self, seed: int, capacity: int, batch: int, stack: int, horizon: int, gamma: float, exponent: float = 1.0):
        self.sampler = PrioritizedSamplingOracle(seed, capacity, exponent)
        self.rb = ReplayOracle(self.sampler, batch, capacity, stack, horizon, gamma)
        self.batch = batch

    def add(self, obs, action, reward, terminal, episode_end) -> None:
        self.rb.add(obs, action, reward, terminal, episode_end, priority=self.sampler.tree.max_recorded_priority)

    def sample(self, beta: float):
        tree = self.sampler.tree
        idx = self.sampler.sample_indices(self.batch)
        keys = np.asarray([self.sampler.index_to_key[i] for i in idx], dtype=np.int32)
        prob = np.asarray([tree.get(int(i)) for i in idx], dtype=np.float64) / tree.root
        w = (len(self.sampler.index_to_key) * prob) ** (-beta)
        return self.rb.gather(keys), keys, (w / w.max()).astype(np.float32)

    def update(self, keys, priorities) -> None:
        self.rb.update(np.asarray(keys), priorities=np.asarray(priorities, dtype=np.float64))
