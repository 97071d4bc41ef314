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
        self._p_nodes, self._p_max, self._p_status = self._d_nodes.data_ptr(), self._d_max.data_ptr(), self._d_status.data_ptr()
        self._q_idx: list = []   # queued sets: leaf indices / values of every op, concatenated on flush
        self._q_val: list = []
        self._q_len: list = []   # ... and the number of entries of every op (arrays)
        self._q_entries = 0
        self._stager = None

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
            self._set_now(idx32, val64)
            return
        self._q_idx.append(idx32)
        self._q_val.append(val64)
        self._q_len.append(np.asarray([m], dtype=np.int32))
        self._q_entries += m
        if self._q_entries >= self._QUEUE_LIMIT:
            self.flush()

    def _enqueue_ops(self, idx32: np.ndarray, val64: np.ndarray, lengths: np.ndarray) -> None:
        """Queue many sets at once: op j covers the next lengths[j] entries of idx32 / val64 (every op at most
        SUMTREE_OP_MAX entries; values may carry the tags of `_enqueue` and SUMTREE_TAG_MAX)."""
        if lengths.size == 0:
            return
        assert int(lengths.max()) <= _lib.SUMTREE_OP_MAX and int(lengths.sum()) == idx32.size == val64.size
        self._q_idx.append(np.ascontiguousarray(idx32, dtype=np.int32))
        self._q_val.append(np.ascontiguousarray(val64, dtype=np.float64))
        self._q_len.append(np.ascontiguousarray(lengths, dtype=np.int32))
        self._q_entries += idx32.size
        if self._q_entries >= self._QUEUE_LIMIT:
            self.flush()

    def _set_now(self, idx32: np.ndarray, val64: np.ndarray) -> None:
        """One `set` applied right away (after whatever is queued): one staged copy + one launch, no queue bookkeeping —
        the path of a batch-sized priority update (PrioritizedSamplingDistribution.update)."""
        self.flush()
        if self._stager is None:
            self._stager = _lib.PinnedStager(1 << 14)
        p_idx, p_val = self._stager.put_ptrs(idx32, val64)
        _lib.check(
            self._lib.isdqn_sumtree_set(self._p_nodes, self._depth, p_idx, p_val, idx32.size, self._p_max, self._p_status,
                                        _lib.stream_ptr()),
            "isdqn_sumtree_set",
        )

    _FLUSH_ENTRIES = 1 << 18  # entries per launch (bounds the pinned staging block)

    def flush(self) -> None:
        """Applies every queued `set`, in order (one launch per _FLUSH_ENTRIES entries)."""
        if not self._q_len:
            return
        idx = np.concatenate(self._q_idx) if len(self._q_idx) > 1 else self._q_idx[0]
        val = np.concatenate(self._q_val) if len(self._q_val) > 1 else self._q_val[0]
        lens = np.concatenate(self._q_len) if len(self._q_len) > 1 else self._q_len[0]
        self._q_idx, self._q_val, self._q_len, self._q_entries = [], [], [], 0
        if self._stager is None:
            self._stager = _lib.PinnedStager(1 << 14)
        off = np.zeros(lens.size + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        op0 = 0
        while op0 < lens.size:
            # ops [op0, op1): as many as fit _FLUSH_ENTRIES entries (at least one)
            op1 = int(np.searchsorted(off, off[op0] + self._FLUSH_ENTRIES, side="right")) - 1
            op1 = min(max(op1, op0 + 1), lens.size)
            e0, e1 = int(off[op0]), int(off[op1])
            p_idx, p_val, p_off = self._stager.put_ptrs(idx[e0:e1], val[e0:e1], (off[op0 : op1 + 1] - e0).astype(np.int32))
            _lib.check(
                self._lib.isdqn_sumtree_set_ops(
                    self._p_nodes, self._depth, p_off, op1 - op0, p_idx, p_val, self._p_max, self._p_status, _lib.stream_ptr(),
                ),
                "isdqn_sumtree_set_ops",
            )
            op0 = op1

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
