"""n-step trajectory accumulator — host logic of `ReplayBuffer.accumulate` / `_replay_element_from_slices`
(slimdqn/sample_collection/replay_buffer.py:102-183), emitting frame REFERENCES instead of stacked copies.

An emitted record is `(refs, action, reward, is_terminal)`:
  refs[0:S]   frame ids of the state stack, refs[S:2S] of the next_state stack, -1 = zero padding
  (episode-start left padding, terminal right padding — replay_buffer.py:131-147)
  reward      sum_{i<n} gamma^i r_{t+i}, folded left to right in Python floats exactly like the reference.

Frame ids come from the `commit(frame)` callback, called the first time an emitted element references a frame
(so frames of dropped, truncated transitions are never stored).  No device code in this file: it is unit-tested on
the CPU against the oracle.
"""
from __future__ import annotations

import collections
from typing import Callable, Iterable, Tuple

import numpy as np


class Frame:
    """A trajectory entry: the observation plus, once committed, its frame id."""

    __slots__ = ("observation", "action", "reward", "frame_id")

    def __init__(self, observation, action, reward):
        self.observation = observation
        self.action = action
        self.reward = reward
        self.frame_id = -1


class NStepAccumulator:
    def __init__(self, stack_size: int, update_horizon: int, gamma: float, commit: Callable[[Frame], int]):
        self.S = stack_size
        self.n = update_horizon
        self.gamma = gamma
        self._commit = commit
        self.trajectory: "collections.deque[Frame]" = collections.deque(maxlen=self.n + self.S)

    def _emit(self, e: int, is_terminal: bool) -> Tuple[np.ndarray, object, float, bool]:
        """The replay element whose state ends at trajectory position `e`:
        state = positions e-S+1..e, next_state = e-S+1+n..e+n, zero where the position does not exist."""
        traj, S, n = self.trajectory, self.S, self.n
        refs = np.full(2 * S, -1, dtype=np.int64)
        r_t = 0.0
        for t in range(len(traj)):
            tr = traj[t]
            if e <= t <= e + n - 1:
                r_t += tr.reward * (self.gamma ** (t - e))
            j = t - (e - S + 1)
            j2 = j - n
            in_state, in_next = 0 <= j < S, 0 <= j2 < S
            if in_state or in_next:
                fid = tr.frame_id if tr.frame_id >= 0 else self._commit(tr)
                if in_state:
                    refs[j] = fid
                if in_next:
                    refs[S + j2] = fid
        return refs, traj[e].action, r_t, is_terminal

    def accumulate(self, observation, action, reward, is_terminal: bool, episode_end: bool) -> Iterable[tuple]:
        """replay_buffer.py:151-183."""
        traj, S, n = self.trajectory, self.S, self.n
        traj.append(Frame(observation, action, reward))
        if is_terminal:
            L = len(traj)
            if L < S + n:
                # terminal before stack_size + update_horizon observations: every sample not yet considered
                for e in range(max(L - 1 - n, 0), L):
                    yield self._emit(e, e + n >= L)
            else:
                # the first element is not terminal: only its next state leads to the terminal state
                yield self._emit(L - 1 - n, False)
                traj.popleft()
                while len(traj) >= S:
                    yield self._emit(S - 1, True)
                    traj.popleft()
            traj.clear()
        else:
            if len(traj) >= 1 + n:
                yield self._emit(len(traj) - 1 - n, False)
            if episode_end:  # truncation: the trajectory is dropped (replay_buffer.py:181-183)
                traj.clear()

    # ------------------------------------------------------------------------------------------ batched form
    def steady_run(self, terminals, episode_ends, start: int) -> int:
        """Length of the run of transitions start, start+1, ... that can be accumulated in closed form: the trajectory is
        full (n + S entries, all committed), no transition of the run is terminal, and a truncation may only end it."""
        if len(self.trajectory) != self.n + self.S or any(f.frame_id < 0 for f in self.trajectory):
            return 0
        N = len(terminals)
        stop = start
        while stop < N and not terminals[stop]:
            stop += 1
            if episode_ends[stop - 1]:
                break
        return stop - start

    def accumulate_run(self, first_frame_id: int, actions, rewards, truncated_at_end: bool):
        """Closed form of `accumulate` for a steady run of m transitions whose observations were committed as the
        consecutive frame ids first_frame_id ... (see steady_run): the deque slides by one per transition and every
        transition emits exactly the element whose state ends at position S - 1 (replay_buffer.py:116-126, 177-183).
        Returns (refs [m][2S] int64, actions [m], rewards [m] float64); no element of such a run is terminal."""
        S, n, D = self.S, self.n, self.n + self.S
        traj = self.trajectory
        m = len(actions)
        prev_ids = np.fromiter((f.frame_id for f in traj), dtype=np.int64, count=D)
        assert (prev_ids >= 0).all(), "steady run on a trajectory with uncommitted frames"
        ids = np.concatenate((prev_ids, first_frame_id + np.arange(m, dtype=np.int64)))
        acts = np.concatenate((np.asarray([f.action for f in traj]), np.asarray(actions)))
        rews = np.concatenate((np.asarray([f.reward for f in traj], dtype=np.float64), np.asarray(rewards, dtype=np.float64)))
        win = np.lib.stride_tricks.sliding_window_view(ids, D)[1 : m + 1]  # the deque after every append
        refs = np.concatenate((win[:, :S], win[:, n : n + S]), axis=1)
        out_actions = acts[S : S + m]
        # r_t = sum_i gamma^i r_{e+i}, folded left to right from 0.0 like the reference's loop (replay_buffer.py:139-142)
        out_rewards = np.zeros(m, dtype=np.float64)
        for i in range(n):
            out_rewards += rews[S + i : S + i + m] * (self.gamma**i)
        traj.clear()
        if not truncated_at_end:
            for j in range(m, m + D):
                fr = Frame(None, acts[j], rews[j])
                fr.frame_id = int(ids[j])
                traj.append(fr)
        return refs, out_actions, out_rewards
