"""ORACLE (test infrastructure): CPU restatement of `slimdqn/sample_collection/replay_buffer.py`.

Element-materialising formulation (every replay element owns its two stacked arrays, as the reference
does), written position-wise instead of with the reference's slice objects:

  an element is defined by its state-end position `e` in the current trajectory
      state      = trajectory[e-S+1 .. e]         (zero where the position does not exist)
      next_state = trajectory[e-S+1+n .. e+n]
      action     = trajectory[e].action
      reward     = sum_{t=e}^{e+n-1} reward_t * gamma**(t-e)   (existing t only, Python-float fold)
  (replay_buffer.py:128-149).  Which `e` are emitted per added transition is `accumulate`
  (replay_buffer.py:151-183); FIFO eviction and the sampler calls are `add` (:185-196); `sample`
  (:198-213) stacks the fields of the drawn keys.

Pinned bit-exact against the unmodified reference in `oracle/make_golden.py`; fixtures in
tests/golden/replay_*.npz.
"""
from __future__ import annotations

import collections
from typing import Any, NamedTuple

import numpy as np


class OracleElement(NamedTuple):
    state: Any
    action: Any
    reward: Any
    next_state: Any
    is_terminal: Any


class ReplayOracle:
    def __init__(
        self,
        sampling_distribution,
        batch_size: int,
        max_capacity: int,
        stack_size: int = 4,
        update_horizon: int = 1,
        gamma: float = 0.99,
        clipping=None,
    ) -> None:
        self.add_count = 0
        self.max_capacity = max_capacity
        self.memory: "collections.OrderedDict[int, OracleElement]" = collections.OrderedDict()
        self.sampler = sampling_distribution
        self.batch_size = batch_size
        self.S = stack_size
        self.n = update_horizon
        self.gamma = gamma
        self.clipping = clipping
        self.traj: collections.deque = collections.deque(maxlen=self.n + self.S)

    # -- replay_buffer.py:128-149 ----------------------------------------------------------------------
    def _element(self, e: int, is_terminal: bool) -> OracleElement:
        traj, S, n = self.traj, self.S, self.n
        obs0 = traj[0][0]
        shape = obs0.shape + (S,)
        state = np.zeros(shape, obs0.dtype)
        nxt = np.zeros(shape, obs0.dtype)
        r = 0.0
        for t in range(len(traj)):
            obs, _a, rew = traj[t][0], traj[t][1], traj[t][2]
            if e <= t <= e + n - 1:
                r += rew * (self.gamma ** (t - e))
            j = t - (e - S + 1)
            if 0 <= j < S:
                state[..., j] = obs
            j2 = t - (e - S + 1 + n)
            if 0 <= j2 < S:
                nxt[..., j2] = obs
        return OracleElement(state, traj[e][1], r, nxt, is_terminal)

    # -- replay_buffer.py:151-183 ----------------------------------------------------------------------
    def _accumulate(self, obs, action, reward, is_terminal, episode_end):
        traj, S, n = self.traj, self.S, self.n
        traj.append((obs, action, reward))
        out = []
        if is_terminal:
            L = len(traj)
            if L < S + n:
                for e in range(max(L - 1 - n, 0), L):
                    out.append(self._element(e, e + n >= L))
            else:
                out.append(self._element(L - 1 - n, False))
                traj.popleft()
                while len(traj) >= S:
                    out.append(self._element(S - 1, True))
                    traj.popleft()
            traj.clear()
        else:
            if len(traj) >= 1 + n:
                out.append(self._element(len(traj) - 1 - n, False))
            if episode_end:
                traj.clear()
        return out

    # -- replay_buffer.py:185-196 ----------------------------------------------------------------------
    def add(self, obs, action, reward, is_terminal, episode_end=False, **kwargs) -> None:
        for el in self._accumulate(obs, action, reward, is_terminal, episode_end):
            key = self.add_count
            self.memory[key] = el
            self.sampler.add(key, **kwargs)
            self.add_count += 1
            if self.add_count > self.max_capacity:
                oldest, _ = self.memory.popitem(last=False)
                self.sampler.remove(oldest)

    # -- replay_buffer.py:198-213 ----------------------------------------------------------------------
    def gather(self, keys) -> OracleElement:
        els = [self.memory[int(k)] for k in keys]
        return OracleElement(*[np.stack([el[f] for el in els]) for f in range(5)])

    def sample(self, size=None) -> OracleElement:
        assert self.add_count
        if size is None:
            size = self.batch_size
        return self.gather(self.sampler.sample(size))

    def update(self, keys, **kwargs) -> None:  # replay_buffer.py:215-220
        self.sampler.update(keys, **kwargs)
