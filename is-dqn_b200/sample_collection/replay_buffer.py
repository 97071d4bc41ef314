"""Replay buffer with a device-resident frame ring — drop-in for `slimdqn/sample_collection/replay_buffer.py`.

Same API (`TransitionElement`, `ReplayElement`, `ReplayBuffer(...)`, `.add`, `.sample`, `.update`, `.add_count`,
`._memory`, `._clipping`), different storage.  The reference keeps, per replay element, two stacked
`(H, W, stack)` arrays in an OrderedDict (replay_buffer.py:88,128-149).  Here every observation frame is stored ONCE
in a circular ring in HBM and an element is eight ring-slot references (zero padding = a slot that holds zeros);
`sample()` is one sampling kernel plus one gather kernel (is-dqn_b200/csrc/{sampler,gather}.cu) whose output is
byte-identical to `np.stack` over the reference's elements.

Layout in HBM (capacity C elements, ring of R frame slots, frame of F bytes padded to a 16-byte stride):
    frames        uint8  [R + 1][stride]     slot R is the all-zero frame
    elem_frames   int32  [C][2*stack]        ring slots of state frames then next_state frames
    elem_action   int64  [C]   elem_reward float64 [C]   elem_terminal uint8 [C]
Element slot = key % C (keys are `add_count`, FIFO eviction keeps the live keys contiguous, replay_buffer.py:190-196).
Frames are committed to the ring lazily, when the first element that references them is emitted, so the ring never
holds frames of dropped (truncated) transitions; R = (1 + update_horizon) * C + stack + update_horizon + 2 is a
proven upper bound of the live frames, overridable with `frame_capacity=`.

The n-step accumulator (`accumulate`, replay_buffer.py:151-183) is host code, as in the reference; it emits frame
references instead of copies.  Host -> device traffic is staged in pinned memory and flushed in bulk.
"""
from __future__ import annotations

import collections
import collections.abc
import typing
from typing import Any, Iterable, Optional

import numpy as np
import numpy.typing as npt

from .. import _lib
from . import ReplayItemID
from .accumulator import Frame as _Frame
from .accumulator import NStepAccumulator


class TransitionElement(typing.NamedTuple):  # replay_buffer.py:18-23
    observation: Optional[npt.NDArray[Any]]
    action: int
    reward: float
    is_terminal: bool
    episode_end: bool = False


class ReplayElement(typing.NamedTuple):
    """A single replay transition element, or a batch of them stacked on axis 0 (replay_buffer.py:26-33).

    The reference's `pack`/`unpack` (snappy) exist for API compatibility and are identities: frames are stored
    once in HBM, there is nothing to compress per element."""

    state: Any
    action: Any
    reward: Any
    next_state: Any
    is_terminal: Any

    def replace(self, **kw) -> "ReplayElement":
        return self._replace(**kw)

    def pack(self) -> "ReplayElement":
        return self

    def unpack(self) -> "ReplayElement":
        return self


class _MemoryView(collections.abc.Mapping):
    """`rb._memory` as the reference exposes it: an ordered key -> ReplayElement mapping (live keys only)."""

    def __init__(self, rb: "ReplayBuffer"):
        self._rb = rb

    def __len__(self):
        return self._rb.add_count - self._rb._oldest_key

    def __iter__(self):
        return iter(range(self._rb._oldest_key, self._rb.add_count))

    def __contains__(self, key):
        return isinstance(key, (int, np.integer)) and self._rb._oldest_key <= key < self._rb.add_count

    def __getitem__(self, key):
        if key not in self:
            raise KeyError(key)
        b = self._rb._gather_keys(np.asarray([key], dtype=np.int64))
        return ReplayElement(b.state[0], b.action[0].item(), b.reward[0].item(), b.next_state[0], bool(b.is_terminal[0]))


class ReplayBuffer:
    def __init__(
        self,
        sampling_distribution,
        batch_size: int,
        max_capacity: int,
        stack_size: int = 4,
        update_horizon: int = 1,
        gamma: float = 0.99,
        checkpoint_duration: int = 4,
        compress: bool = True,
        clipping: callable = None,
        frame_capacity: Optional[int] = None,
        staging_frames: int = 2048,
        pinned_ring: int = 0,
    ):
        torch = _lib.require_cuda()
        self._torch = torch
        self._lib = _lib.load()
        self._device = torch.device("cuda", torch.cuda.current_device())
        self.add_count = 0
        self._oldest_key = 0
        self._max_capacity = max_capacity
        self._compress = compress  # accepted for API compatibility; frames are stored once, uncompressed, in HBM
        self._memory = _MemoryView(self)

        self._sampling_distribution = sampling_distribution

        self._checkpoint_duration = checkpoint_duration
        self._batch_size = batch_size

        self._stack_size = stack_size
        self._update_horizon = update_horizon
        self._gamma = gamma
        self._clipping = clipping

        self._accumulator = NStepAccumulator(stack_size, update_horizon, gamma, self._commit)
        self._trajectory = self._accumulator.trajectory

        # live elements: transiently max_capacity + 1 (replay_buffer.py:190-196 adds before it evicts)
        self._slots = max_capacity + 1
        self._frame_capacity = (
            int(frame_capacity)
            if frame_capacity is not None
            else (1 + update_horizon) * (max_capacity + 1) + stack_size + update_horizon + 2
        )
        self._staging_frames = max(1, min(int(staging_frames), self._frame_capacity))
        self._allocated = False
        self._next_frame = 0        # frames committed so far (ids 0 .. _next_frame-1)
        self._flushed_frames = 0    # ... of which already on the device
        self._flushed_elems = 0     # elements whose metadata is already on the device
        self._pending_event = None
        self._action_dtype = None
        # pinned_ring = N > 0: sample() returns views of one of N pinned blocks (round-robin) in the learner's packed batch
        # layout instead of fresh arrays, so learn_on_batch can DMA the batch back without a host-side copy; such a
        # batch is valid until N - 1 further sample() calls of the same size.  0 (default): fresh arrays, as the reference.
        self._pinned_ring = int(pinned_ring)
        self._host_rings = {}       # batch size -> ring of pinned blocks handed out by sample()

    # ------------------------------------------------------------------------------------------ allocation
    def _allocate(self, observation: np.ndarray) -> None:
        t = self._torch
        obs = np.asarray(observation)
        self._obs_shape = obs.shape
        self._obs_dtype = obs.dtype
        self._elem_size = obs.dtype.itemsize
        if self._elem_size not in (1, 2, 4, 8):
            raise _lib.IsdqnNativeError(f"unsupported observation dtype {obs.dtype}")
        self._frame_elems = int(np.prod(obs.shape, dtype=np.int64)) if obs.shape else 1
        self._frame_bytes = self._frame_elems * self._elem_size
        self._frame_stride = (self._frame_bytes + 15) // 16 * 16
        R, C, S = self._frame_capacity, self._slots, self._stack_size
        self._zero_slot = R
        self._d_frames = t.zeros((R + 1, self._frame_stride), dtype=t.uint8, device=self._device)
        self._d_elem_frames = t.full((C, 2 * S), R, dtype=t.int32, device=self._device)
        self._d_action = t.zeros(C, dtype=t.int64, device=self._device)
        self._d_reward = t.zeros(C, dtype=t.float64, device=self._device)
        self._d_terminal = t.zeros(C, dtype=t.uint8, device=self._device)
        # pinned host mirrors / staging
        self._h_stage = t.zeros((self._staging_frames, self._frame_stride), dtype=t.uint8).pin_memory()
        self._h_stage_np = self._h_stage.numpy()
        self._h_elem_frames = t.full((C, 2 * S), R, dtype=t.int32).pin_memory()
        self._h_action = t.zeros(C, dtype=t.int64).pin_memory()
        self._h_reward = t.zeros(C, dtype=t.float64).pin_memory()
        self._h_terminal = t.zeros(C, dtype=t.uint8).pin_memory()
        self._hn_elem_frames = self._h_elem_frames.numpy()
        self._hn_action = self._h_action.numpy()
        self._hn_reward = self._h_reward.numpy()
        self._hn_terminal = self._h_terminal.numpy()
        self._elem_min_frame = np.zeros(C, dtype=np.int64)  # smallest frame id an element references
        self._allocated = True

    # --------------------------------------------------------------------------------- host -> device flush
    def _wait_pending(self) -> None:
        if self._pending_event is not None:
            self._pending_event.synchronize()
            self._pending_event = None

    def _copy_ring(self, dst, src, first: int, count: int, modulo: int) -> None:
        """dst[(first + i) % modulo] = src[(first + i) % modulo] for i < count, as <= 2 contiguous copies."""
        a = first % modulo
        n1 = min(count, modulo - a)
        dst[a : a + n1].copy_(src[a : a + n1], non_blocking=True)
        if count > n1:
            dst[: count - n1].copy_(src[: count - n1], non_blocking=True)

    def _flush_frames(self) -> None:
        n = self._next_frame - self._flushed_frames
        if n == 0:
            return
        R, St = self._frame_capacity, self._staging_frames
        done = 0
        while done < n:  # staging index = id % St, ring index = id % R: copy maximal runs contiguous in both
            fid = self._flushed_frames + done
            run = min(n - done, St - fid % St, R - fid % R)
            self._d_frames[fid % R : fid % R + run].copy_(self._h_stage[fid % St : fid % St + run], non_blocking=True)
            done += run
        self._flushed_frames = self._next_frame

    def _flush(self) -> None:
        """Pushes every pending frame and element record to the device (async, pinned source)."""
        if not self._allocated:
            return
        dirty = self._next_frame > self._flushed_frames or self.add_count > self._flushed_elems
        if not dirty:
            return
        self._flush_frames()
        n = self.add_count - self._flushed_elems
        if n > 0:
            C = self._slots
            first = self._flushed_elems
            if n >= C:
                first, n = 0, C
            for dst, src in (
                (self._d_elem_frames, self._h_elem_frames),
                (self._d_action, self._h_action),
                (self._d_reward, self._h_reward),
                (self._d_terminal, self._h_terminal),
            ):
                self._copy_ring(dst, src, first, n, C)
            self._flushed_elems = self.add_count
        ev = self._torch.cuda.Event()
        ev.record()
        self._pending_event = ev

    # -------------------------------------------------------------------------------------- frame commits
    def _commit(self, fr: _Frame) -> int:
        if fr.frame_id >= 0:
            return fr.frame_id
        if isinstance(fr.observation, self._torch.Tensor) and fr.observation.is_cuda:
            fr.frame_id = self._commit_many(fr.observation.unsqueeze(0), 1)
            self._check_ring()
            return fr.frame_id
        self._wait_pending()  # an async flush may still be reading the pinned staging slots
        live_from = self._elem_min_frame[self._oldest_key % self._slots] if self.add_count > self._oldest_key else self._next_frame
        if self._next_frame - live_from >= self._frame_capacity:
            raise _lib.IsdqnNativeError(
                f"frame ring overflow: {self._frame_capacity} slots cannot hold the frames of the live elements; "
                "pass a larger frame_capacity= to ReplayBuffer"
            )
        if self._next_frame - self._flushed_frames >= self._staging_frames:
            self._wait_pending()
            self._flush_frames()
            ev = self._torch.cuda.Event()
            ev.record()
            ev.synchronize()  # the staging slots are about to be overwritten
        obs = np.ascontiguousarray(fr.observation)
        if obs.shape != self._obs_shape or obs.dtype != self._obs_dtype:
            raise ValueError(f"observation {obs.shape}/{obs.dtype} differs from the first one {self._obs_shape}/{self._obs_dtype}")
        fid = self._next_frame
        self._h_stage_np[fid % self._staging_frames, : self._frame_bytes] = obs.reshape(-1).view(np.uint8)
        fr.frame_id = fid
        self._next_frame = fid + 1
        return fid

    def accumulate(self, transition: TransitionElement) -> Iterable[tuple]:
        """Add a transition to the accumulator, maybe receive valid replay elements (replay_buffer.py:151-183).
        Yields (frame refs, action, reward, is_terminal) records instead of materialised stacks."""
        if not self._allocated:
            o = transition.observation
            self._allocate(o.cpu().numpy() if isinstance(o, self._torch.Tensor) else o)
        return self._accumulator.accumulate(
            transition.observation, transition.action, transition.reward, transition.is_terminal, transition.episode_end
        )

    def add(self, transition: TransitionElement, **kwargs: Any) -> None:
        """replay_buffer.py:185-196: key = add_count; sampler.add; FIFO eviction + sampler.remove."""
        for refs, action, reward, is_terminal in self.accumulate(transition):
            self._wait_pending()
            if self._action_dtype is None:
                self._action_dtype = np.asarray(action).dtype
            key = ReplayItemID(self.add_count)
            slot = key % self._slots
            R = self._frame_capacity
            valid = refs >= 0
            self._hn_elem_frames[slot] = np.where(valid, refs % R, self._zero_slot)
            self._elem_min_frame[slot] = refs[valid].min() if valid.any() else self._next_frame
            self._hn_action[slot] = action
            self._hn_reward[slot] = reward
            self._hn_terminal[slot] = 1 if is_terminal else 0
            self._sampling_distribution.add(key, **kwargs)
            self.add_count += 1
            if self.add_count > self._max_capacity:
                oldest_key = self._oldest_key
                self._oldest_key += 1
                self._sampling_distribution.remove(oldest_key)

    # ------------------------------------------------------------------------------------------ batched add
    def _check_ring(self) -> None:
        live_from = self._elem_min_frame[self._oldest_key % self._slots] if self.add_count > self._oldest_key else self._next_frame
        if self._next_frame - live_from > self._frame_capacity:
            raise _lib.IsdqnNativeError(
                f"frame ring overflow: {self._frame_capacity} slots cannot hold the frames of the live elements; "
                "pass a larger frame_capacity= to ReplayBuffer"
            )

    def _commit_many(self, observations, m: int) -> int:
        """Commits m consecutive observations (host array or CUDA tensor, [m, *obs_shape]) to the frame ring; returns the
        id of the first.  Device observations never touch the host: one or two device-to-device copies into the ring."""
        t = self._torch
        R = self._frame_capacity
        first = self._next_frame
        if isinstance(observations, t.Tensor) and observations.is_cuda:
            self._wait_pending()
            self._flush_frames()  # staged frames have smaller ids: they reach the ring first (same stream)
            src = observations.contiguous().view(t.uint8).reshape(m, self._frame_bytes)
            done = 0
            while done < m:
                a = (first + done) % R
                run = min(m - done, R - a)
                self._d_frames[a : a + run, : self._frame_bytes].copy_(src[done : done + run], non_blocking=True)
                done += run
            self._next_frame = first + m
            self._flushed_frames = self._next_frame
            return first
        obs = np.ascontiguousarray(observations).reshape(m, -1).view(np.uint8)
        if obs.shape[1] != self._frame_bytes:
            raise ValueError("observation block does not match the buffer's observation shape / dtype")
        St = self._staging_frames
        done = 0
        while done < m:
            self._wait_pending()
            if self._next_frame - self._flushed_frames >= St:
                self._flush_frames()
                ev = t.cuda.Event()
                ev.record()
                ev.synchronize()  # the staging slots are about to be overwritten
            fid = self._next_frame
            room = St - (fid - self._flushed_frames)
            run = min(m - done, room, St - fid % St)
            self._h_stage_np[fid % St : fid % St + run, : self._frame_bytes] = obs[done : done + run]
            self._next_frame = fid + run
            done += run
        return first

    def add_batch(self, observations, actions, rewards, is_terminals, episode_ends=None, priorities=None) -> None:
        """N environment steps at once: identical in effect to
            for i in range(N): rb.add(TransitionElement(observations[i], actions[i], rewards[i], is_terminals[i], episode_ends[i]),
                                      [priority=priorities[i]])
        (replay_buffer.py:151-196), but the steady part of a trajectory — full n-step window, no terminal — is accumulated
        in closed form (accumulator.accumulate_run), its frames reach the ring with one copy (device-to-device when
        `observations` is a CUDA tensor: a device-resident environment never touches the host), the element records
        are written with array assignments and the sampler receives one run (`_add_remove_run`).  Transitions around
        episode boundaries take the per-transition path.  priorities: None, "max" (insert at max_recorded_priority,
        resolved on the device) or one value per transition (prioritized samplers only)."""
        t = self._torch
        N = len(actions)
        if episode_ends is None:
            episode_ends = is_terminals
        is_dev = isinstance(observations, t.Tensor) and observations.is_cuda
        terms = np.asarray(is_terminals, dtype=bool)
        ends = np.asarray(episode_ends, dtype=bool)
        acts = np.asarray(actions)
        rews = np.asarray(rewards)
        if not self._allocated and N:
            first = observations[0]
            self._allocate(first.cpu().numpy() if is_dev else np.asarray(first))
        per_key = priorities is not None and not isinstance(priorities, str)
        acc = self._accumulator
        cap = self._max_capacity
        i = 0
        while i < N:
            m = acc.steady_run(terms, ends, i)
            if m == 0:
                kw = {} if priorities is None else {"priority": priorities[i] if per_key else priorities}
                obs_i = observations[i]
                self.add(TransitionElement(obs_i, acts[i], rews[i], bool(terms[i]), bool(ends[i])), **kw)
                i += 1
                continue
            m = min(m, cap)  # one pass never laps the element ring
            self._wait_pending()
            fid0 = self._commit_many(observations[i : i + m], m)
            refs, e_act, e_rew = acc.accumulate_run(fid0, acts[i : i + m], rews[i : i + m], bool(ends[i + m - 1]))
            if self._action_dtype is None:
                self._action_dtype = np.asarray(acts[i]).dtype
            keys = self.add_count + np.arange(m, dtype=np.int64)
            slots = keys % self._slots
            self._hn_elem_frames[slots] = refs % self._frame_capacity  # (no padding inside a steady run)
            self._elem_min_frame[slots] = refs[:, 0]
            self._hn_action[slots] = e_act
            self._hn_reward[slots] = e_rew
            self._hn_terminal[slots] = 0
            evict_from = max(cap - self.add_count, 0)  # add j is followed by an eviction once add_count + j + 1 > cap
            n_evict = max(m - evict_from, 0)
            sd = self._sampling_distribution
            pr = None if priorities is None else (priorities if not per_key else priorities[i : i + m])
            if hasattr(sd, "_add_remove_run"):
                sd._add_remove_run(int(keys[0]), m, self._oldest_key, min(evict_from, m), pr)
            else:  # a user-supplied sampling distribution: the one-at-a-time calls
                for j in range(m):
                    kw = {} if pr is None else {"priority": pr[j] if per_key else pr}
                    sd.add(ReplayItemID(int(keys[j])), **kw)
                    if j >= evict_from:
                        sd.remove(self._oldest_key + (j - evict_from))
            self.add_count += m
            self._oldest_key += n_evict
            self._check_ring()
            i += m

    # --------------------------------------------------------------------------------------------- sampling
    def _gather_slots_device(self, d_slots, out_dtype: int = _lib.OUT_RAW, out: Optional[ReplayElement] = None):
        """Runs the gather kernel for int32 CUDA element slots; returns CUDA tensors (no synchronisation).
        `out`: preallocated CUDA tensors to gather into (e.g. the learner's persistent batch buffers)."""
        t = self._torch
        self._flush()
        n = d_slots.numel()
        S = self._stack_size
        if out is not None:
            state, action, reward, nxt, terminal = out
        elif out_dtype == _lib.OUT_RAW:
            shape = (n, self._frame_elems * S * self._elem_size)
            state = t.empty(shape, dtype=t.uint8, device=self._device)
            nxt = t.empty(shape, dtype=t.uint8, device=self._device)
        else:
            dt = t.float32 if out_dtype == _lib.OUT_F32 else t.bfloat16
            state = t.empty((n,) + tuple(self._obs_shape) + (S,), dtype=dt, device=self._device)
            nxt = t.empty_like(state)
        if out is None:
            action = t.empty(n, dtype=t.int64, device=self._device)
            reward = t.empty(n, dtype=t.float64, device=self._device)
            terminal = t.empty(n, dtype=t.uint8, device=self._device)
        _lib.check(
            self._lib.isdqn_gather_stacks(
                self._d_frames.data_ptr(), self._frame_stride, self._frame_elems, self._elem_size, S,
                self._d_elem_frames.data_ptr(), self._d_action.data_ptr(), self._d_reward.data_ptr(),
                self._d_terminal.data_ptr(), d_slots.data_ptr(), n, out_dtype, state.data_ptr(), nxt.data_ptr(),
                action.data_ptr(), reward.data_ptr(), terminal.data_ptr(), _lib.stream_ptr(),
            ),
            "isdqn_gather_stacks",
        )
        return ReplayElement(state, action, reward, nxt, terminal)

    def _to_host(self, b: ReplayElement) -> ReplayElement:
        """Device batch -> host numpy arrays: fresh arrays by default; with `pinned_ring=N` views of ONE pinned block in
        the learner's packed batch layout (_lib.batch_pack_layout), so that
        `agent.learn_on_batch(params, opt_state, rb.sample())` moves the batch back with a single DMA and no host copy."""
        if self._pinned_ring <= 0:
            n = b.action.numel()
            shape = (n,) + tuple(self._obs_shape) + (self._stack_size,)
            state = b.state.cpu().numpy().view(self._obs_dtype).reshape(shape)
            nxt = b.next_state.cpu().numpy().view(self._obs_dtype).reshape(shape)
            action = b.action.cpu().numpy().astype(self._action_dtype or np.int64, copy=False)
            return ReplayElement(state, action, b.reward.cpu().numpy(), nxt, b.is_terminal.cpu().numpy().astype(np.bool_))
        t = self._torch
        n = b.action.numel()
        shape = (n,) + tuple(self._obs_shape) + (self._stack_size,)
        row_bytes = int(np.prod(shape[1:])) * self._elem_size
        ring = self._host_rings.get(n)
        if ring is None:
            total, offs = _lib.batch_pack_layout(n, row_bytes)
            ring = self._host_rings[n] = {"i": 0, "offs": offs, "blocks": [_lib.pinned_block(total) for _ in range(self._pinned_ring)]}
        blk = ring["blocks"][ring["i"] % self._pinned_ring]
        ring["i"] += 1
        offs = ring["offs"]

        def field(name, dtype):
            o, nb = offs[name]
            return blk[o : o + nb].view(dtype)

        field("state", t.uint8).copy_(b.state.reshape(-1).view(t.uint8), non_blocking=True)
        field("next_state", t.uint8).copy_(b.next_state.reshape(-1).view(t.uint8), non_blocking=True)
        field("action", t.int64).copy_(b.action, non_blocking=True)
        field("reward", t.float64).copy_(b.reward, non_blocking=True)
        field("terminal", t.uint8).copy_(b.is_terminal, non_blocking=True)
        t.cuda.current_stream().synchronize()
        h = blk.numpy()

        def host(name, dtype):
            o, nb = offs[name]
            return h[o : o + nb].view(dtype)

        action = host("action", np.int64)
        if self._action_dtype is not None:
            action = action.astype(self._action_dtype, copy=False)
        return ReplayElement(host("state", self._obs_dtype).reshape(shape), action, host("reward", np.float64),
                             host("next_state", self._obs_dtype).reshape(shape), host("terminal", np.bool_))

    def _gather_keys(self, keys: np.ndarray) -> ReplayElement:
        slots = (np.asarray(keys, dtype=np.int64) % self._slots).astype(np.int32)
        d_slots = self._torch.from_numpy(slots).to(self._device)
        return self._to_host(self._gather_slots_device(d_slots))

    def capturable_sample(self, out: ReplayElement, size=None):
        """For a captured training step: returns (prepare, enqueue, token).  enqueue(stream_ptr) launches draw -> gather
        into `out` (persistent CUDA tensors, e.g. `agent.batch_buffers(B)`) without allocating anything, so the two
        launches can live in the same CUDA graph as the learner step; prepare() pushes whatever the host has pending
        (frames, element records, key-map patches, the live count) and must run before every launch or replay; `token`
        changes when a device table was re-allocated (a captured graph is then stale).  Uniform sampler only."""
        from .samplers import UniformSamplingDistribution

        sd = self._sampling_distribution
        if type(sd) is not UniformSamplingDistribution or not self._allocated:
            return None
        if size is None:
            size = self._batch_size
        _, _, d_slot, draw = sd.capturable_draw(size, self._slots)
        state, action, reward, nxt, terminal = out
        lib, S = self._lib, self._stack_size

        def prepare():
            self._flush()
            sd._flush_maps()

        def enqueue(stream_ptr: int) -> None:
            draw(stream_ptr)
            _lib.check(
                lib.isdqn_gather_stacks(
                    self._d_frames.data_ptr(), self._frame_stride, self._frame_elems, self._elem_size, S,
                    self._d_elem_frames.data_ptr(), self._d_action.data_ptr(), self._d_reward.data_ptr(),
                    self._d_terminal.data_ptr(), d_slot.data_ptr(), size, _lib.OUT_RAW, state.data_ptr(), nxt.data_ptr(),
                    action.data_ptr(), reward.data_ptr(), terminal.data_ptr(), stream_ptr,
                ),
                "isdqn_gather_stacks",
            )

        def token():
            return (id(self), sd._d_index_to_key.data_ptr(), self._d_frames.data_ptr(), size)

        return prepare, enqueue, token

    def capturable_prioritized_step(self, out: ReplayElement, d_weight, size=None):
        """The prioritized counterpart of `capturable_sample` for a captured training step: returns
        (prepare, enqueue, update, set_beta, token) or None.  enqueue(stream_ptr): draw (importance weights into
        `d_weight`) -> gather into `out`; update(stream_ptr, d_td_abs, rows, offset): priorities of the drawn keys <- mean
        |TD| (+ offset, ^ alpha); prepare(): pending frames / records / key maps / sum-tree ops; set_beta(beta)."""
        from .samplers import PrioritizedSamplingDistribution

        sd = self._sampling_distribution
        if not isinstance(sd, PrioritizedSamplingDistribution) or not self._allocated:
            return None
        if size is None:
            size = self._batch_size
        if size > 1024:
            return None
        _, d_slot, draw, update, set_beta = sd.capturable_train_draw(size, self._slots, d_weight)
        state, action, reward, nxt, terminal = out
        lib, S = self._lib, self._stack_size

        def prepare():
            self._flush()
            sd._flush_maps()
            sd._sum_tree.flush()

        def enqueue(stream_ptr: int) -> None:
            draw(stream_ptr)
            _lib.check(
                lib.isdqn_gather_stacks(
                    self._d_frames.data_ptr(), self._frame_stride, self._frame_elems, self._elem_size, S,
                    self._d_elem_frames.data_ptr(), self._d_action.data_ptr(), self._d_reward.data_ptr(),
                    self._d_terminal.data_ptr(), d_slot.data_ptr(), size, _lib.OUT_RAW, state.data_ptr(), nxt.data_ptr(),
                    action.data_ptr(), reward.data_ptr(), terminal.data_ptr(), stream_ptr,
                ),
                "isdqn_gather_stacks",
            )

        def token():
            return (id(self), sd._d_index_to_key.data_ptr(), self._d_frames.data_ptr(), size, "prioritized",
                    sd._d_key_to_index.data_ptr(), sd._sum_tree._d_nodes.data_ptr())

        return prepare, enqueue, update, set_beta, token

    def sample_device(self, size=None, out_dtype: int = _lib.OUT_RAW, out: Optional[ReplayElement] = None,
                      return_keys: bool = False, beta: Optional[float] = None):
        """Device-resident `sample`: draw -> key -> slot -> gather without leaving the GPU.  For uint8 stack-4
        frames the state tensors are (size, H, W, stack) uint8 (or f32/bf16 normalised with out_dtype).
        return_keys: also return the int32 CUDA keys of the batch (what `update` / `update_device` take); beta (prioritized
        samplers): also return the float32 CUDA importance weights (N P(i))^-beta / max.  Returns batch,
        (batch, keys) or (batch, keys, weights)."""
        assert self.add_count, ValueError("No samples in replay buffer!")
        if size is None:
            size = self._batch_size
        sd = self._sampling_distribution
        if beta is not None:
            _, d_key, d_slot = sd.sample_device(size, self._slots, want_prob=True)
        else:
            _, d_key, d_slot = sd.sample_device(size, self._slots)
        b = self._gather_slots_device(d_slot, out_dtype, out)
        if out is None and out_dtype == _lib.OUT_RAW and self._elem_size == 1:
            shape = (size,) + tuple(self._obs_shape) + (self._stack_size,)
            b = b._replace(state=b.state.view(shape), next_state=b.next_state.view(shape))
        if beta is not None:
            return b, d_key, sd.importance_weights(beta)
        if return_keys:
            return b, d_key
        return b

    def sample(self, size=None, return_keys: bool = False, beta: Optional[float] = None):
        """Sample a batch of elements from the replay buffer (replay_buffer.py:198-213); host numpy arrays.
        return_keys / beta: as in `sample_device`, with host arrays (the reference's `sample` returns the batch only, so a
        prioritized training loop could not name the elements it should re-prioritise)."""
        assert self.add_count, ValueError("No samples in replay buffer!")
        if size is None:
            size = self._batch_size
        if beta is not None:
            b, d_key, d_w = self.sample_device(size, beta=beta)
            host = self._to_host(b._replace(state=b.state.reshape(size, -1), next_state=b.next_state.reshape(size, -1)))
            keys, w = d_key.cpu().numpy(), d_w.cpu().numpy()
            self._sampling_distribution._pull_rng_state()
            self._sampling_distribution.check_status()
            return host, keys, w
        samples = self._sampling_distribution.sample(size)
        batch = self._gather_keys(samples)
        return (batch, samples) if return_keys else batch

    def update(self, keys, **kwargs: Any) -> None:  # replay_buffer.py:215-220
        self._sampling_distribution.update(keys, **kwargs)

    def update_device(self, d_keys, d_priorities, prio_rows: int = 0, offset: float = 0.0) -> None:
        """`update(keys, priorities=...)` for keys and priorities that live on the device (e.g. the keys of
        `sample_device(return_keys=True)` and the |TD| matrix of the step, `agent.td_abs`): no synchronisation."""
        self._sampling_distribution.update_device(d_keys, d_priorities, prio_rows, offset)
