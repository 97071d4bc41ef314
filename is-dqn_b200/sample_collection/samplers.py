"""Sampling distributions on the device — drop-in for `slimdqn/sample_collection/samplers.py`.

Same classes, methods and attributes (`_key_to_index`, `_index_to_key`, `_rng_key`, `_sum_tree`).  The key <-> dense
index bookkeeping stays on the host exactly as in the reference (it is what `add`/`remove` mutate one key at a
time); device mirrors of `index_to_key` (and, for the prioritized sampler, of `key_to_index`) are kept in sync by
patches so that the draw -> key -> element-slot chain runs in ONE kernel with no host round trip:

  UniformSamplingDistribution.sample      -> isdqn_sample_uniform      (PCG64 next32 stream + Lemire, bit-exact with
                                                                         numpy's Generator.integers; samplers.py:39-49)
  PrioritizedSamplingDistribution.sample  -> isdqn_sample_prioritized  (Generator.uniform(0, root) + sum-tree descent;
                                                                         samplers.py:105-116)
  PrioritizedSamplingDistribution.update  -> isdqn_sumtree_set (host keys) / isdqn_sumtree_set_keys (device keys and
                                                                         priorities: `update_device`; samplers.py:76-88)

`sample(size)` returns host `np.int32` keys like the reference; `sample_device(size, capacity)` returns CUDA tensors
(index, key, slot = key % capacity) without synchronising.  The numpy Generator in `_rng_key` seeds the device
stream; after every host `sample` its state is written back so `_rng_key` stays where the reference's would be.

`_add_remove_run` is the batched form of the `add(key)` / `remove(oldest)` pairs `ReplayBuffer.add` issues
(replay_buffer.py:190-196): one tight loop over the host maps, one patch list, and — prioritized — one queue of
sum-tree ops, identical in effect to the one-at-a-time calls.
"""
from __future__ import annotations

import numpy as np

from .. import _lib
from . import ReplayItemID
from . import sum_tree


class UniformSamplingDistribution:
    """A uniform sampling distribution (reference: samplers.py:13-49)."""

    def __init__(self, seed: int) -> None:
        torch = _lib.require_cuda()
        self._torch = torch
        self._lib = _lib.load()
        self._device = torch.device("cuda", torch.cuda.current_device())
        self._rng_key = np.random.default_rng(seed)
        self._d_rng = torch.zeros(6, dtype=torch.int64, device=self._device)
        self._d_uniform_ws = None
        self._push_rng_state()

        self._key_to_index = {}
        self._index_to_key = []
        # device mirror of _index_to_key (int32), grown geometrically, updated by patches
        self._d_index_to_key = torch.zeros(self._initial_table_size(), dtype=torch.int32, device=self._device)
        self._patches: dict = {}
        self._stager = None
        # number of live dense indices, on the device: a captured sampler launch reads it when it runs
        self._d_n_valid = torch.zeros(1, dtype=torch.int32, device=self._device)
        self._n_valid_dev = 0

    def _initial_table_size(self) -> int:
        return 1024

    # -- RNG state mirror ------------------------------------------------------------------------------
    def _push_rng_state(self) -> None:
        st = self._rng_key.bit_generator.state
        s, inc = st["state"]["state"], st["state"]["inc"]
        m = (1 << 64) - 1
        vals = np.array([s & m, s >> 64, inc & m, inc >> 64, st["has_uint32"], st["uinteger"]], dtype=np.uint64)
        self._d_rng.copy_(self._torch.from_numpy(vals.view(np.int64)))

    def _pull_rng_state(self) -> None:
        v = self._d_rng.cpu().numpy().view(np.uint64)
        st = self._rng_key.bit_generator.state
        st["state"]["state"] = (int(v[1]) << 64) | int(v[0])
        st["state"]["inc"] = (int(v[3]) << 64) | int(v[2])
        st["has_uint32"] = int(v[4])
        st["uinteger"] = int(v[5])
        self._rng_key.bit_generator.state = st

    # -- key maps (samplers.py:22-37) ---------------------------------------------------------------------
    def add(self, key: ReplayItemID) -> None:
        index = len(self._index_to_key)
        self._key_to_index[key] = index
        self._index_to_key.append(key)
        self._patches[index] = key

    def remove(self, key: ReplayItemID) -> None:
        assert key in self._key_to_index, ValueError(f"Key {key} not found.")
        index = self._key_to_index[key]
        # for efficient O(1) pop on the keys: the last key moves into the hole
        last_key = self._index_to_key[-1]
        self._index_to_key[index] = last_key
        self._key_to_index[last_key] = index
        self._index_to_key.pop()
        self._key_to_index.pop(key)
        if index < len(self._index_to_key):
            self._patches[index] = last_key
        self._patches.pop(len(self._index_to_key), None)

    def _add_remove_run(self, first_key: int, count: int, first_evicted: int, evict_from: int, priorities=None):
        """`count` consecutive keys first_key, first_key + 1, ... added in order; from the evict_from-th add on, every add
        is followed by the removal of the oldest key (first_evicted, first_evicted + 1, ...) — the call sequence of
        ReplayBuffer.add (replay_buffer.py:190-196) for `count` elements.  Returns (dense index every new key received
        at its add, dense index of every removed key at its removal)."""
        kti, itk, patches = self._key_to_index, self._index_to_key, self._patches
        add_idx = np.empty(count, dtype=np.int64)
        rem_idx = np.empty(max(count - evict_from, 0), dtype=np.int64)
        for j in range(count):
            key = first_key + j
            index = len(itk)
            kti[key] = index
            itk.append(key)
            patches[index] = key
            add_idx[j] = index
            if j >= evict_from:
                old = first_evicted + (j - evict_from)
                hole = kti[old]
                last_key = itk[-1]
                itk[hole] = last_key
                kti[last_key] = hole
                itk.pop()
                del kti[old]
                n = len(itk)
                if hole < n:
                    patches[hole] = last_key
                patches.pop(n, None)
                rem_idx[j - evict_from] = hole
        return add_idx, rem_idx

    def _put(self, *arrays):
        if self._stager is None:
            self._stager = _lib.PinnedStager(1 << 14)
        return self._stager.put(*arrays)

    def _grow_tables(self, n: int) -> None:
        t = self._torch
        if n > self._d_index_to_key.numel():
            grown = t.zeros(max(n, 2 * self._d_index_to_key.numel()), dtype=t.int32, device=self._device)
            grown[: self._d_index_to_key.numel()] = self._d_index_to_key
            self._d_index_to_key = grown

    def _flush_maps(self) -> None:
        n = len(self._index_to_key)
        t = self._torch
        self._grow_tables(n)
        if n != self._n_valid_dev:
            self._d_n_valid.fill_(n)
            self._n_valid_dev = n
        if not self._patches:
            return
        if len(self._patches) * 4 >= n:  # cheaper to resend the table
            self._d_index_to_key[:n] = t.from_numpy(np.asarray(self._index_to_key, dtype=np.int32)).to(self._device)
            self._resend_inverse(n)
        else:
            idx = np.fromiter(self._patches.keys(), dtype=np.int32, count=len(self._patches))
            val = np.fromiter(self._patches.values(), dtype=np.int32, count=len(self._patches))
            d_idx, d_val = self._put(idx, val)
            _lib.check(
                self._lib.isdqn_scatter_rows_i32(
                    self._d_index_to_key.data_ptr(), 1, d_idx.data_ptr(), d_val.data_ptr(), idx.size, _lib.stream_ptr()
                ),
                "isdqn_scatter_rows_i32",
            )
            self._patch_inverse(idx, val, d_idx)
        self._patches = {}

    # (the prioritized sampler mirrors key -> index as well)
    def _resend_inverse(self, n: int) -> None:
        pass

    def _patch_inverse(self, idx, keys, d_idx) -> None:
        pass

    # -- sampling -------------------------------------------------------------------------------------------
    def _draw_device(self, size: int, capacity: int):
        t = self._torch
        d_index = t.empty(size, dtype=t.int32, device=self._device)
        d_key = t.empty(size, dtype=t.int32, device=self._device)
        d_slot = t.empty(size, dtype=t.int32, device=self._device)
        if self._d_uniform_ws is None:  # scratch of the many-CTA variant (large draws)
            self._d_uniform_ws = t.zeros(int(self._lib.isdqn_sample_uniform_workspace_bytes()), dtype=t.uint8, device=self._device)
        _lib.check(
            self._lib.isdqn_sample_uniform_ws(
                self._d_rng.data_ptr(), len(self._index_to_key), size, self._d_index_to_key.data_ptr(),
                max(int(capacity), 1), d_index.data_ptr(), d_key.data_ptr(), d_slot.data_ptr(),
                self._d_uniform_ws.data_ptr(), self._d_uniform_ws.numel(), _lib.stream_ptr(),
            ),
            "isdqn_sample_uniform_ws",
        )
        return d_index, d_key, d_slot

    def sample_device(self, size: int, capacity: int):
        """(dense index, key, element slot) as int32 CUDA tensors; nothing synchronises."""
        assert self._index_to_key, ValueError("No keys to sample from.")
        self._flush_maps()
        return self._draw_device(size, capacity)

    def capturable_draw(self, size: int, capacity: int):
        """(d_index, d_key, d_slot, enqueue): persistent output tensors and a function enqueue(stream_ptr) that launches
        the draw into them with the live count read on the device — it allocates nothing and may be captured in a CUDA
        graph.  The caller runs `_flush_maps()` before every launch / replay; the table pointer is part of the returned
        token so that a re-allocated table invalidates a captured graph."""
        t = self._torch
        d_index = t.empty(size, dtype=t.int32, device=self._device)
        d_key = t.empty(size, dtype=t.int32, device=self._device)
        d_slot = t.empty(size, dtype=t.int32, device=self._device)
        lib, cap = self._lib, max(int(capacity), 1)

        def enqueue(stream_ptr: int) -> None:
            _lib.check(
                lib.isdqn_sample_uniform_dev(self._d_rng.data_ptr(), self._d_n_valid.data_ptr(), size, self._d_index_to_key.data_ptr(),
                                             cap, d_index.data_ptr(), d_key.data_ptr(), d_slot.data_ptr(), stream_ptr),
                "isdqn_sample_uniform_dev",
            )

        return d_index, d_key, d_slot, enqueue

    def sample(self, size: int):
        assert self._index_to_key, ValueError("No keys to sample from.")
        self._flush_maps()
        _, d_key, _ = UniformSamplingDistribution._draw_device(self, size, 1)
        keys = d_key.cpu().numpy()
        self._pull_rng_state()
        return keys

    def check_status(self) -> None:
        """Raises what the reference would have raised for the device-side draws since the last check (nothing to
        check for the uniform sampler)."""


class PrioritizedSamplingDistribution(UniformSamplingDistribution):
    """A prioritized sampling distribution (reference: samplers.py:52-116)."""

    def __init__(self, seed: int, max_capacity: int, priority_exponent: float = 1.0) -> None:
        self._max_capacity = max_capacity
        self._priority_exponent = priority_exponent
        self._sum_tree = sum_tree.SumTree(self._max_capacity)
        super().__init__(seed=seed)
        # device mirror of _key_to_index, addressed by key mod (max_capacity + 1) — the element slot of a ReplayBuffer of
        # the same capacity (live keys are then a contiguous window, so the slots never collide); -1 = no live key
        self._n_slots = int(max_capacity) + 1
        self._d_key_to_index = self._torch.full((self._n_slots,), -1, dtype=self._torch.int32, device=self._device)
        self._d_keys_ws = None
        self._last_prob = None

    def _initial_table_size(self) -> int:
        # every leaf of the tree can be the end of a descent: the table covers them all (an index past the live keys is
        # flagged by the kernel, never dereferenced past the table)
        return max(1024, 2 ** (self._sum_tree._depth - 1))

    def _exp(self, priority):
        return 0.0 if priority == 0.0 else priority**self._priority_exponent

    def add(self, key: ReplayItemID, priority: float) -> None:
        super().add(key)
        if priority is None:
            priority = 0.0
        if isinstance(priority, str):  # "max": insert at max_recorded_priority (resolved on the device, no read-back)
            self._sum_tree._enqueue(np.asarray([self._key_to_index[key]], dtype=np.int32), self._max_tag())
            return
        self._sum_tree.set(
            self._key_to_index[key],
            0.0 if priority == 0.0 else priority**self._priority_exponent,
        )

    def _max_tag(self) -> np.ndarray:
        if self._priority_exponent != 1.0:
            # max_recorded_priority already holds exponentiated values (sum_tree.py:32 tracks what `set` receives): the
            # device tag is exact only when the exponent is 1; otherwise read it back like a host training loop would
            return np.asarray([self._exp(self._sum_tree.max_recorded_priority)], dtype=np.float64)
        return np.asarray([_lib.SUMTREE_TAG_MAX], dtype=np.float64)

    def update(self, keys, priorities) -> None:
        if not isinstance(keys, np.ndarray):
            keys = np.asarray([keys], dtype=np.int32)
        priorities = np.where(priorities == 0.0, 0.0, priorities**self._priority_exponent)
        kti = self._key_to_index
        idx = np.fromiter((kti[key] for key in keys.tolist()), dtype=np.int32, count=keys.size)
        tree = self._sum_tree
        vals = np.ascontiguousarray(priorities, dtype=np.float64).reshape(-1)
        if 2 < idx.size <= _lib.SUMTREE_SET_MAX and vals.size == idx.size and not (vals < 0.0).any() and not np.isnan(vals).any():
            tree._set_now(idx, vals)  # (indices come from the live maps: in range by construction)
        else:
            tree.set(idx, priorities)

    def update_device(self, d_keys, d_priorities, prio_rows: int = 0, offset: float = 0.0) -> None:
        """`update` for int32 keys and priorities that live on the device (nothing synchronises): d_priorities is
        float64 [n], float32 [n] or — prio_rows = K > 0 — the float32 [K][n] per-head |TD| matrix of the step, averaged
        over the heads; `offset` is added to every priority before the exponent.  A key that is no longer live leaves the tree untouched and makes the next `check_status()` raise
        KeyError, as `self._key_to_index[key]` does in the reference (samplers.py:84)."""
        t = self._torch
        n = int(d_keys.numel())
        self._flush_maps()
        self._sum_tree.flush()
        kind = 2 if prio_rows > 0 else (0 if d_priorities.dtype == t.float64 else 1)
        if kind == 1 and d_priorities.dtype != t.float32:
            raise TypeError("priorities must be float32 or float64")
        need = int(self._lib.isdqn_sumtree_set_keys_workspace_bytes(n))
        if self._d_keys_ws is None or self._d_keys_ws.numel() < need:
            self._d_keys_ws = t.empty(need, dtype=t.uint8, device=self._device)
        tree = self._sum_tree
        _lib.check(
            self._lib.isdqn_sumtree_set_keys(
                tree._d_nodes.data_ptr(), tree._depth, d_keys.data_ptr(), d_priorities.data_ptr(), kind, int(prio_rows), n,
                float(offset), float(self._priority_exponent), self._d_key_to_index.data_ptr(), self._n_slots,
                self._d_index_to_key.data_ptr(), len(self._index_to_key), tree._d_max.data_ptr(), tree._d_status.data_ptr(),
                self._d_keys_ws.data_ptr(), self._d_keys_ws.numel(), _lib.stream_ptr(),
            ),
            "isdqn_sumtree_set_keys",
        )

    def remove(self, key: ReplayItemID) -> None:
        index = self._key_to_index[key]
        last_index = len(self._index_to_key) - 1
        if index == last_index:
            # If index and last_index are the same, simply set the priority to 0.0.
            self._sum_tree.set(index, 0.0)
        else:
            # Swap priorities with current index and last index (samplers.py:99-102).  The value moved is
            # "whatever leaf last_index holds when this op runs": resolved on the device, no read-back.
            self._sum_tree._enqueue(
                np.asarray([index, last_index], dtype=np.int32),
                np.asarray([-(1.0 + last_index), 0.0], dtype=np.float64),
            )
        super().remove(key)

    def _add_remove_run(self, first_key: int, count: int, first_evicted: int, evict_from: int, priorities=None):
        """The uniform run plus the sum-tree ops of `add` (samplers.py:67-74) and `remove` (:90-103) for every key, queued
        in call order.  priorities: None (0.0), "max", or one value per key."""
        n_before = len(self._index_to_key)
        add_idx, rem_idx = super()._add_remove_run(first_key, count, first_evicted, evict_from)
        n_rem = rem_idx.size
        # value of every add op
        if priorities is None:
            add_val = np.zeros(count, dtype=np.float64)
        elif isinstance(priorities, str):
            add_val = np.repeat(self._max_tag(), count)
        else:
            e = self._priority_exponent
            add_val = np.asarray([0.0 if p == 0.0 else p**e for p in priorities], dtype=np.float64)
            assert add_val.size == count
            if (add_val < 0.0).any():
                raise AssertionError("Values must be positive.")  # sum_tree.py:31
        # dense index of the last key when removal r runs: the table holds n_before + (adds so far) - (removals so far)
        # entries, i.e. n_before + evict_from + r + 1 - r - 1
        last_idx = np.full(n_rem, n_before + evict_from, dtype=np.int64) if n_rem else np.zeros(0, dtype=np.int64)
        same = rem_idx == last_idx
        n_entries = count + 2 * n_rem - int(same.sum())
        idx = np.empty(n_entries, dtype=np.int32)
        val = np.empty(n_entries, dtype=np.float64)
        lens = np.empty(count + n_rem, dtype=np.int32)
        # op order: add_0 .. add_{evict_from-1}, then (add_j, remove_j) pairs
        op_of_add = np.arange(count, dtype=np.int64)
        op_of_add[evict_from:] += np.arange(n_rem, dtype=np.int64)
        op_of_rem = op_of_add[evict_from:] + 1
        lens[op_of_add] = 1
        lens[op_of_rem] = np.where(same, 1, 2)
        off = np.zeros(count + n_rem + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        idx[off[op_of_add]] = add_idx
        val[off[op_of_add]] = add_val
        r0 = off[op_of_rem]
        idx[r0] = rem_idx
        val[r0] = np.where(same, 0.0, -(1.0 + last_idx))
        two = ~same
        idx[r0[two] + 1] = last_idx[two]
        val[r0[two] + 1] = 0.0
        if idx.size and int(idx.max()) + self._sum_tree._first_leaf_offset >= self._sum_tree._d_nodes.numel():
            raise IndexError(f"index out of bounds for the sum tree of capacity {self._max_capacity}")
        if evict_from > 1:
            # The adds before the first eviction set DISTINCT, ASCENDING leaves (fresh dense indices): one `set` of all of
            # them performs, node by node, the same additions in the same order as the single sets (np.add.at walks the
            # sorted leaves), so they are merged into ops of up to SUMTREE_OP_MAX entries, which the kernel spreads over
            # a whole CTA instead of one warp-op at a time.  (The tag "max" resolves to the same value in every merged
            # entry: nothing in such an op can raise max_recorded_priority above the tag's own value.)
            n_merge = evict_from
            step = _lib.SUMTREE_OP_MAX
            merged = np.full((n_merge + step - 1) // step, step, dtype=np.int32)
            merged[-1] = n_merge - step * (merged.size - 1)
            lens = np.concatenate((merged, lens[n_merge:]))
        self._sum_tree._enqueue_ops(idx, val, lens)
        return add_idx, rem_idx

    # -- device mirror of key -> index ----------------------------------------------------------------------
    def _resend_inverse(self, n: int) -> None:
        t = self._torch
        keys = np.asarray(self._index_to_key, dtype=np.int64)
        inv = np.full(self._n_slots, -1, dtype=np.int32)
        inv[keys % self._n_slots] = np.arange(n, dtype=np.int32)
        self._d_key_to_index.copy_(t.from_numpy(inv).to(self._device))

    def _patch_inverse(self, idx, keys, d_idx) -> None:
        slots = (keys.astype(np.int64) % self._n_slots).astype(np.int32)
        (d_slots,) = self._put(slots)
        _lib.check(
            self._lib.isdqn_scatter_rows_i32(
                self._d_key_to_index.data_ptr(), 1, d_slots.data_ptr(), d_idx.data_ptr(), idx.size, _lib.stream_ptr()
            ),
            "isdqn_scatter_rows_i32",
        )

    # -- sampling -------------------------------------------------------------------------------------------
    def _draw_device(self, size: int, capacity: int, want_targets: bool = False, want_prob: bool = False):
        t = self._torch
        tree = self._sum_tree
        tree.flush()
        d_index = t.empty(size, dtype=t.int32, device=self._device)
        d_key = t.empty(size, dtype=t.int32, device=self._device)
        d_slot = t.empty(size, dtype=t.int32, device=self._device)
        d_target = t.empty(size, dtype=t.float64, device=self._device) if want_targets else None
        d_prob = t.empty(size, dtype=t.float64, device=self._device) if want_prob else None
        self._last_prob = d_prob
        _lib.check(
            self._lib.isdqn_sample_prioritized(
                self._d_rng.data_ptr(), tree._d_nodes.data_ptr(), tree._depth, size, len(self._index_to_key),
                self._d_index_to_key.data_ptr(), max(int(capacity), 1), d_index.data_ptr(), d_key.data_ptr(),
                d_slot.data_ptr(), _lib.ptr(d_target), _lib.ptr(d_prob), tree._d_status.data_ptr(), _lib.stream_ptr(),
            ),
            "isdqn_sample_prioritized",
        )
        if want_targets:
            return d_index, d_key, d_slot, d_target
        return d_index, d_key, d_slot

    def sample_device(self, size: int, capacity: int, want_prob: bool = False):
        """(dense index, key, element slot) as int32 CUDA tensors; nothing synchronises — what the reference would raise
        for these draws (empty tree, a descent that ends past the live keys) is raised by the next `check_status()`.
        want_prob: the probability of every draw (leaf / root, float64) is left in `self._last_prob`."""
        assert self._index_to_key, ValueError("No keys to sample from.")
        self._flush_maps()
        return self._draw_device(size, capacity, want_prob=want_prob)

    def capturable_train_draw(self, size: int, capacity: int, d_weight):
        """For a captured prioritized training step: (d_key, d_slot, draw, update, set_beta).  draw(stream_ptr) launches the
        batch draw with the live count and beta read on the device and writes the importance weights into `d_weight`
        (float32 [size]); update(stream_ptr, d_td_abs, rows, offset) writes the mean |TD| of the drawn keys back as their
        priorities (isdqn_sumtree_set_keys).  Neither allocates nor reads host state: both can be captured in the learner's
        CUDA graph.  The caller runs `_flush_maps()` and `self._sum_tree.flush()` before every launch / replay."""
        t = self._torch
        lib, tree, cap = self._lib, self._sum_tree, max(int(capacity), 1)
        d_key = t.empty(size, dtype=t.int32, device=self._device)
        d_slot = t.empty(size, dtype=t.int32, device=self._device)
        d_prob = t.empty(size, dtype=t.float64, device=self._device)
        d_beta = t.zeros(1, dtype=t.float64, device=self._device)
        ws = t.empty(int(lib.isdqn_sumtree_set_keys_workspace_bytes(size)), dtype=t.uint8, device=self._device)
        state = {"beta": None}

        def set_beta(beta: float) -> None:
            if state["beta"] != float(beta):
                d_beta.fill_(float(beta))
                state["beta"] = float(beta)

        def draw(stream_ptr: int) -> None:
            self._last_prob = d_prob
            _lib.check(
                lib.isdqn_sample_prioritized_train(
                    self._d_rng.data_ptr(), tree._d_nodes.data_ptr(), tree._depth, size, self._d_n_valid.data_ptr(),
                    self._d_index_to_key.data_ptr(), cap, d_key.data_ptr(), d_slot.data_ptr(), d_prob.data_ptr(),
                    d_beta.data_ptr(), d_weight.data_ptr(), tree._d_status.data_ptr(), stream_ptr,
                ),
                "isdqn_sample_prioritized_train",
            )

        def update(stream_ptr: int, d_td_abs, rows: int, offset: float) -> None:
            # (n_valid: the table length — a dead key is caught through its -1 entry in the key -> index mirror)
            _lib.check(
                lib.isdqn_sumtree_set_keys(
                    tree._d_nodes.data_ptr(), tree._depth, d_key.data_ptr(), d_td_abs.data_ptr(), 2, int(rows), size, float(offset),
                    float(self._priority_exponent), self._d_key_to_index.data_ptr(), self._n_slots,
                    self._d_index_to_key.data_ptr(), int(self._d_index_to_key.numel()), tree._d_max.data_ptr(),
                    tree._d_status.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr,
                ),
                "isdqn_sumtree_set_keys",
            )

        return d_key, d_slot, draw, update, set_beta

    def importance_weights(self, beta: float):
        """Importance-sampling weights of the latest `sample_device(..., want_prob=True)` draw, float32 CUDA tensor:
        (N * P(i)) ** -beta, normalised by the largest weight of the batch (Schaul et al. 2016; new functionality — the
        reference never trains from its prioritized sampler)."""
        if self._last_prob is None:
            raise RuntimeError("importance_weights() follows sample_device(..., want_prob=True)")
        w = (self._last_prob * float(len(self._index_to_key))).pow(-float(beta))
        return (w / w.max()).to(self._torch.float32)

    def check_status(self) -> None:
        st = self._sum_tree._check_status()
        if st & _lib.ST_KEY_MISSING:
            raise KeyError("update_device: a key is not in the sampler")  # samplers.py:84
        if st & _lib.ST_EMPTY_TREE:
            raise AttributeError("'numpy.ndarray' object has no attribute 'keys'")  # samplers.py:105-108 on an empty tree
        if st & _lib.ST_DESCENT_ASSERT:
            raise AssertionError()  # sum_tree.py:82

    def sample(self, size: int):
        if self._sum_tree.root == 0.0:
            # the reference consumes the uniform draws and then trips over `.keys` on an ndarray (samplers.py:105-108)
            keys = super().sample(size).keys
            return keys
        self._flush_maps()
        _, d_key, _ = self._draw_device(size, 1)
        keys = d_key.cpu().numpy()
        self._pull_rng_state()
        st = self._sum_tree._check_status()
        if st & _lib.ST_DESCENT_ASSERT:
            raise AssertionError()  # sum_tree.py:82
        return keys
