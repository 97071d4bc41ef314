"""Sampling distributions on the device — drop-in for `slimdqn/sample_collection/samplers.py`.

Same classes, methods and attributes (`_key_to_index`, `_index_to_key`, `_rng_key`, `_sum_tree`).  The key <-> dense
index bookkeeping stays on the host exactly as in the reference (it is what `add`/`remove` mutate one key at a
time); a device mirror of `index_to_key` is kept in sync by patches so that the draw -> key -> element-slot chain
runs in ONE kernel with no host round trip:

  UniformSamplingDistribution.sample      -> isdqn_sample_uniform      (PCG64 next32 stream + Lemire, bit-exact with
                                                                         numpy's Generator.integers; samplers.py:39-49)
  PrioritizedSamplingDistribution.sample  -> isdqn_sample_prioritized  (Generator.uniform(0, root) + sum-tree descent;
                                                                         samplers.py:105-116)

`sample(size)` returns host `np.int32` keys like the reference; `sample_device(size, capacity)` returns CUDA tensors
(index, key, slot = key % capacity) without synchronising.  The numpy Generator in `_rng_key` seeds the device
stream; after every host `sample` its state is written back so `_rng_key` stays where the reference's would be.
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
        self._d_index_to_key = torch.zeros(1024, dtype=torch.int32, device=self._device)
        self._patches: dict = {}

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

    def _flush_maps(self) -> None:
        n = len(self._index_to_key)
        t = self._torch
        if n > self._d_index_to_key.numel():
            grown = t.zeros(max(n, 2 * self._d_index_to_key.numel()), dtype=t.int32, device=self._device)
            grown[: self._d_index_to_key.numel()] = self._d_index_to_key
            self._d_index_to_key = grown
        if not self._patches:
            return
        if len(self._patches) * 4 >= n:  # cheaper to resend the table
            self._d_index_to_key[:n] = t.from_numpy(np.asarray(self._index_to_key, dtype=np.int32)).to(self._device)
        else:
            idx = np.fromiter(self._patches.keys(), dtype=np.int32, count=len(self._patches))
            val = np.fromiter(self._patches.values(), dtype=np.int32, count=len(self._patches))
            d_idx, d_val = t.from_numpy(idx).to(self._device), t.from_numpy(val).to(self._device)
            _lib.check(
                self._lib.isdqn_scatter_rows_i32(
                    self._d_index_to_key.data_ptr(), 1, d_idx.data_ptr(), d_val.data_ptr(), idx.size, _lib.stream_ptr()
                ),
                "isdqn_scatter_rows_i32",
            )
        self._patches = {}

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

    def sample(self, size: int):
        assert self._index_to_key, ValueError("No keys to sample from.")
        self._flush_maps()
        _, d_key, _ = UniformSamplingDistribution._draw_device(self, size, 1)
        keys = d_key.cpu().numpy()
        self._pull_rng_state()
        return keys


class PrioritizedSamplingDistribution(UniformSamplingDistribution):
    """A prioritized sampling distribution (reference: samplers.py:52-116)."""

    def __init__(self, seed: int, max_capacity: int, priority_exponent: float = 1.0) -> None:
        self._max_capacity = max_capacity
        self._priority_exponent = priority_exponent
        self._sum_tree = sum_tree.SumTree(self._max_capacity)
        super().__init__(seed=seed)

    def add(self, key: ReplayItemID, priority: float) -> None:
        super().add(key)
        if priority is None:
            priority = 0.0
        self._sum_tree.set(
            self._key_to_index[key],
            0.0 if priority == 0.0 else priority**self._priority_exponent,
        )

    def update(self, keys, priorities) -> None:
        if not isinstance(keys, np.ndarray):
            keys = np.asarray([keys], dtype=np.int32)
        priorities = np.where(priorities == 0.0, 0.0, priorities**self._priority_exponent)
        self._sum_tree.set(
            np.fromiter((self._key_to_index[key] for key in keys), dtype=np.int32),
            priorities,
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

    def _draw_device(self, size: int, capacity: int, want_targets: bool = False):
        t = self._torch
        tree = self._sum_tree
        tree.flush()
        d_index = t.empty(size, dtype=t.int32, device=self._device)
        d_key = t.empty(size, dtype=t.int32, device=self._device)
        d_slot = t.empty(size, dtype=t.int32, device=self._device)
        d_target = t.empty(size, dtype=t.float64, device=self._device) if want_targets else None
        _lib.check(
            self._lib.isdqn_sample_prioritized(
                self._d_rng.data_ptr(), tree._d_nodes.data_ptr(), tree._depth, size, self._d_index_to_key.data_ptr(),
                max(int(capacity), 1), d_index.data_ptr(), d_key.data_ptr(), d_slot.data_ptr(), _lib.ptr(d_target),
                tree._d_status.data_ptr(), _lib.stream_ptr(),
            ),
            "isdqn_sample_prioritized",
        )
        if want_targets:
            return d_index, d_key, d_slot, d_target
        return d_index, d_key, d_slot

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
