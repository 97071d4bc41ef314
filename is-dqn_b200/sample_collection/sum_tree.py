"""Sum tree on the device — drop-in for `slimdqn/sample_collection/sum_tree.py` (same class, methods, attributes
and error behaviour), with the float64 heap living in HBM and `set` / `query` running as CUDA kernels
(is-dqn_b200/csrc/sumtree.cu) that reproduce NumPy's `np.add.at` fold order bit for bit.

Host-side differences a caller can observe: none in values.  `set` is asynchronous: calls are queued and applied
in order by ONE launch (`isdqn_sumtree_set_ops`) the next time something reads the tree (`get`, `root`, `query`,
`_nodes`, `max_recorded_priority`, a device sampler).
"""
from __future__ import annotations

import numpy as np

from .. import _lib


class SumTree:
    """A vectorized sum tree whose nodes live on the GPU (reference: sum_tree.py:8-102)."""

    _QUEUE_LIMIT = 1 << 16  # flush when this many (index, value) pairs are pending

    def __init__(self, capacity: int) -> None:  # sum_tree.py:11-18
        assert capacity > 0, "Capacity to sum tree must be positive."
        torch = _lib.require_cuda()
        self._torch = torch
        self._lib = _lib.load()
        self._capacity = capacity
        self._depth = int(np.ceil(np.log2(capacity))) + 1
        self._first_leaf_offset = (2 ** (self._depth - 1)) - 1
        self._device = torch.device("cuda", torch.cuda.current_device())
        self._d_nodes = torch.zeros((2**self._depth) - 1, dtype=torch.float64, device=self._device)
        self._d_max = torch.ones(1, dtype=torch.float64, device=self._device)  # max_recorded_priority = 1.0
        self._d_status = torch.zeros(1, dtype=torch.int32, device=self._device)
        self._q_idx: list = []
        self._q_val: list = []
        self._q_off = [0]

    # ---------------------------------------------------------------------------------------------- set
    def set(self, indices, values) -> None:
        """Set the value at a given leaf node index (sum_tree.py:20-47)."""
        if isinstance(indices, (int, np.integer)):
            indices = np.asarray([indices], np.int32)
        if isinstance(values, (int, float, np.floating)):
            values = np.asarray([values], np.float64)
        assert indices.shape == values.shape, "Indices and values must have the same shape."
        assert (values >= 0.0).all(), "Values must be positive."
        idx = np.ascontiguousarray(indices, dtype=np.int64).reshape(-1)
        if idx.size and (idx.min() < 0 or idx.max() + self._first_leaf_offset >= self._d_nodes.numel()):
            raise IndexError(f"index out of bounds for the sum tree of capacity {self._capacity}")  # NumPy's IndexError
        self._enqueue(idx.astype(np.int32), np.ascontiguousarray(values, dtype=np.float64).reshape(-1))

    def _enqueue(self, idx32: np.ndarray, val64: np.ndarray) -> None:
        """Queue one `set`.  A value -(1+j) means "the current value of leaf j" (used by the prioritized
        sampler's swap-remove, samplers.py:99-102, so that eviction needs no device read-back)."""
        m = idx32.size
        if m == 0:
            return
        if m > _lib.SUMTREE_OP_MAX:
            self.flush()
            if m > _lib.SUMTREE_SET_MAX:
                raise _lib.IsdqnNativeError(
                    f"SumTree.set with {m} indices exceeds the kernel limit {_lib.SUMTREE_SET_MAX}"
                )
            t = self._torch
            d_idx = t.from_numpy(idx32).to(self._device)
            d_val = t.from_numpy(val64).to(self._device)
            _lib.check(
                self._lib.isdqn_sumtree_set(
                    self._d_nodes.data_ptr(), self._depth, d_idx.data_ptr(), d_val.data_ptr(), m,
                    self._d_max.data_ptr(), self._d_status.data_ptr(), _lib.stream_ptr(),
                ),
                "isdqn_sumtree_set",
            )
            return
        self._q_idx.append(idx32)
        self._q_val.append(val64)
        self._q_off.append(self._q_off[-1] + m)
        if self._q_off[-1] >= self._QUEUE_LIMIT:
            self.flush()

    def flush(self) -> None:
        """Applies every queued `set`, in order, with one launch."""
        if len(self._q_off) == 1:
            return
        t = self._torch
        idx = t.from_numpy(np.concatenate(self._q_idx)).to(self._device)
        val = t.from_numpy(np.concatenate(self._q_val)).to(self._device)
        off = t.from_numpy(np.asarray(self._q_off, dtype=np.int32)).to(self._device)
        n_ops = len(self._q_off) - 1
        self._q_idx, self._q_val, self._q_off = [], [], [0]
        _lib.check(
            self._lib.isdqn_sumtree_set_ops(
                self._d_nodes.data_ptr(), self._depth, off.data_ptr(), n_ops, idx.data_ptr(), val.data_ptr(),
                self._d_max.data_ptr(), self._d_status.data_ptr(), _lib.stream_ptr(),
            ),
            "isdqn_sumtree_set_ops",
        )

    def _check_status(self) -> int:
        st = int(self._d_status.item())
        if st:
            self._d_status.zero_()
        if st & _lib.ST_NEGATIVE_VALUE:
            raise AssertionError("Values must be positive.")
        if st & (_lib.ST_INDEX_RANGE | _lib.ST_OP_TOO_LARGE):
            raise IndexError("sum tree index out of range")
        return st

    # ------------------------------------------------------------------------------------------ readers
    def get(self, index):
        """Get the value at a given leaf node index (sum_tree.py:49-51)."""
        self.flush()
        if isinstance(index, (int, np.integer)):
            return np.float64(self._d_nodes[self._first_leaf_offset + int(index)].item())
        idx = self._torch.as_tensor(np.asarray(index, dtype=np.int64) + self._first_leaf_offset, device=self._device)
        return self._d_nodes[idx].cpu().numpy()

    @property
    def root(self) -> float:
        """The root value (total sum) of the sum tree (sum_tree.py:53-56)."""
        self.flush()
        return np.float64(self._d_nodes[0].item())

    @property
    def max_recorded_priority(self) -> float:
        self.flush()
        return float(self._d_max.item())

    @property
    def _nodes(self) -> np.ndarray:
        """Host copy of the heap (the reference's tests poke `_nodes`)."""
        self.flush()
        return self._d_nodes.cpu().numpy()

    # -------------------------------------------------------------------------------------------- query
    def query(self, targets):
        """Find the smallest index where target < cumulative value up to index (sum_tree.py:58-102)."""
        if isinstance(targets, (int, float)):
            targets = np.asarray([targets], np.float64)
        targets = np.asarray(targets)
        root = self.root
        if not ((targets >= 0) & (targets < root)).all():
            raise ValueError(f"Targets must be in the interval [0.0, {root}).")
        t = self._torch
        flat = np.ascontiguousarray(targets, dtype=np.float64).reshape(-1)
        d_t = t.from_numpy(flat).to(self._device)
        d_out = self.query_device(d_t)
        out = d_out.cpu().numpy()
        st = self._check_status()
        if st & _lib.ST_DESCENT_ASSERT:
            raise AssertionError()  # sum_tree.py:82
        return out.reshape(targets.shape)

    def query_device(self, d_targets):
        """Device-resident query: float64 CUDA tensor in, int32 CUDA tensor out, no synchronisation."""
        self.flush()
        d_out = self._torch.empty(d_targets.numel(), dtype=self._torch.int32, device=self._device)
        _lib.check(
            self._lib.isdqn_sumtree_query(
                self._d_nodes.data_ptr(), self._depth, d_targets.data_ptr(), d_targets.numel(), d_out.data_ptr(),
                self._d_status.data_ptr(), _lib.stream_ptr(),
            ),
            "isdqn_sumtree_query",
        )
        return d_out
