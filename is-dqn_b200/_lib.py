"""ctypes binding of libisdqn_b200.so (the C-ABI declared in include/isdqn_b200.h).

There is no CPU fallback: if the library is missing or CUDA is unavailable the product raises
`IsdqnNativeError` — loudly, at the first use of a device object.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (ISDQN_LIB: another build of the same library, e.g. an experiment variant next to the default one)
LIB_PATH = os.environ.get("ISDQN_LIB") or os.path.join(_HERE, "lib", "libisdqn_b200.so")

ABI_VERSION = 3
MAX_FEATURES = 8
MAX_LEAVES = 64
SUMTREE_SET_MAX = 8192
SUMTREE_OP_MAX = 1024

ST_NEGATIVE_VALUE = 1
ST_TARGET_RANGE = 2
ST_DESCENT_ASSERT = 4
ST_EMPTY_TREE = 8
ST_INDEX_RANGE = 16
ST_OP_TOO_LARGE = 32
ST_KEY_MISSING = 64
SUMTREE_TAG_MAX = -0.5

OUT_RAW, OUT_F32, OUT_BF16 = 0, 1, 2
ARCH_CNN, ARCH_FC, ARCH_IMPALA = 0, 1, 2
COMPUTE_F32, COMPUTE_BF16 = 0, 1


class IsdqnNativeError(RuntimeError):
    pass


class Net(C.Structure):
    _fields_ = [
        ("arch", C.c_int32),
        ("layer_norm", C.c_int32),
        ("obs_h", C.c_int32),
        ("obs_w", C.c_int32),
        ("obs_c", C.c_int32),
        ("n_features", C.c_int32),
        ("features", C.c_int32 * MAX_FEATURES),
        ("n_heads", C.c_int32),
        ("n_actions", C.c_int32),
    ]


class Layout(C.Structure):
    _fields_ = [
        ("n_leaves", C.c_int32),
        ("offset", C.c_int64 * MAX_LEAVES),
        ("size", C.c_int64 * MAX_LEAVES),
        ("total", C.c_int64),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("d_state", C.c_void_p),
        ("d_next_state", C.c_void_p),
        ("d_action", C.c_void_p),
        ("d_reward", C.c_void_p),
        ("d_terminal", C.c_void_p),
    ]


class Train(C.Structure):
    _fields_ = [
        ("gamma_n", C.c_float),
        ("lr", C.c_float),
        ("b1", C.c_float),
        ("b2", C.c_float),
        ("eps", C.c_float),
        ("batch", C.c_int32),
        ("batch_global", C.c_int32),
        ("d_params", C.c_void_p),
        ("d_grads", C.c_void_p),
        ("d_mu", C.c_void_p),
        ("d_nu", C.c_void_p),
        ("d_count", C.c_void_p),
        ("d_losses", C.c_void_p),
        ("d_workspace", C.c_void_p),
        ("workspace_bytes", C.c_int64),
        ("nccl_comm", C.c_void_p),
        ("compute_dtype", C.c_int32),
        ("d_workspace_tc", C.c_void_p),
        ("workspace_tc_bytes", C.c_int64),
        ("d_params_bf16", C.c_void_p),
        ("refresh_shadow", C.c_int32),
        ("d_cumulated", C.c_void_p),
        ("d_is_weights", C.c_void_p),
        ("d_td_abs", C.c_void_p),
    ]


_P = C.c_void_p
_I32 = C.c_int32
_I64 = C.c_int64
_F = C.c_float

# name -> (restype, argtypes); every symbol include/isdqn_b200.h declares
PROTOTYPES = {
    "isdqn_abi_version": (C.c_int, []),
    "isdqn_strerror": (C.c_char_p, [C.c_int]),
    "isdqn_last_cuda_error": (C.c_char_p, []),
    "isdqn_sumtree_query": (C.c_int, [_P, C.c_int, _P, _I64, _P, _P, _P]),
    "isdqn_sumtree_set": (C.c_int, [_P, C.c_int, _P, _P, _I32, _P, _P, _P]),
    "isdqn_sumtree_set_ops": (C.c_int, [_P, C.c_int, _P, _I32, _P, _P, _P, _P, _P]),
    "isdqn_sample_uniform": (C.c_int, [_P, _I32, _I32, _P, _I32, _P, _P, _P, _P]),
    "isdqn_sample_prioritized": (C.c_int, [_P, _P, C.c_int, _I32, _I32, _P, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "isdqn_sample_prioritized_train": (C.c_int, [_P, _P, C.c_int, _I32, _P, _P, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "isdqn_sumtree_set_keys_workspace_bytes": (_I64, [_I32]),
    "isdqn_sumtree_set_keys": (C.c_int, [_P, C.c_int, _P, _P, _I32, _I32, _I32, C.c_double, C.c_double, _P, _I32, _P, _I32, _P, _P, _P, _I64, _P]),
    "isdqn_scatter_rows_i32": (C.c_int, [_P, _I32, _P, _P, _I32, _P]),
    "isdqn_scatter_rows_f64": (C.c_int, [_P, _P, _P, _I32, _P]),
    "isdqn_gather_stacks": (
        C.c_int,
        [_P, _I64, _I32, _I32, _I32, _P, _P, _P, _P, _P, _I32, _I32, _P, _P, _P, _P, _P, _P],
    ),
    "isdqn_net_layout": (C.c_int, [C.POINTER(Net), C.POINTER(Layout)]),
    "isdqn_forward_workspace_bytes": (_I64, [C.POINTER(Net), _I32]),
    "isdqn_learn_workspace_bytes": (_I64, [C.POINTER(Net), _I32]),
    "isdqn_learn_workspace_tc_bytes": (_I64, [C.POINTER(Net), _I32]),
    "isdqn_cast_f32_to_bf16": (C.c_int, [_P, _P, _I64, _P]),
    "isdqn_tc_gemm_bf16": (C.c_int, [_P, _I64, _I32, _P, _I64, _I32, _P, _I32, _I32, _I32, _I32, _P]),
    "isdqn_forward": (C.c_int, [C.POINTER(Net), _P, _P, _I32, _I32, _P, _P, _I64, _P]),
    "isdqn_heads_td_loss": (C.c_int, [_P, _P, _P, _P, _F, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "isdqn_heads_td_loss_weighted": (C.c_int, [_P, _P, _P, _P, _F, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "isdqn_adam_step": (C.c_int, [_P, _P, _P, _P, _P, _F, _F, _F, _F, _I64, _P]),
    "isdqn_adam_step_nocount": (C.c_int, [_P, _P, _P, _P, _P, _F, _F, _F, _F, _I64, _P]),
    "isdqn_shift_heads": (C.c_int, [_P, _P, _I32, _I32, _I32, _P]),
    "isdqn_loss_on_batch": (C.c_int, [C.POINTER(Net), C.POINTER(Train), C.POINTER(Batch), _P, _P]),
    "isdqn_learn_on_batch": (C.c_int, [C.POINTER(Net), C.POINTER(Train), C.POINTER(Batch), _P]),
    "isdqn_grad_on_batch": (C.c_int, [C.POINTER(Net), C.POINTER(Train), C.POINTER(Batch), _P]),
    "isdqn_best_action": (C.c_int, [C.POINTER(Net), _P, _P, _I32, _I32, _P, _P, _I64, _P]),
    "isdqn_graph_begin": (C.c_int, [_P]),
    "isdqn_graph_end": (C.c_int, [_P, C.POINTER(C.c_void_p)]),
    "isdqn_graph_launch": (C.c_int, [_P, _P]),
    "isdqn_graph_destroy": (C.c_int, [_P]),
    "isdqn_profile_begin": (C.c_int, []),
    "isdqn_profile_end": (C.c_int, [_P, _I32, C.c_char_p, _I32, C.POINTER(C.c_float)]),
    "isdqn_spin": (C.c_int, [_P, _I32]),
    "isdqn_trace_set": (C.c_int, [_P]),
    "isdqn_sample_uniform_dev": (C.c_int, [_P, _P, _I32, _P, _I32, _P, _P, _P, _P]),
    "isdqn_sample_uniform_workspace_bytes": (C.c_int64, []),
    "isdqn_sample_uniform_ws": (C.c_int, [_P, _I32, _I32, _P, _I32, _P, _P, _P, _P, _I64, _P]),
    "isdqn_event_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "isdqn_event_destroy": (C.c_int, [_P]),
    "isdqn_event_record": (C.c_int, [_P, _P]),
    "isdqn_event_synchronize": (C.c_int, [_P]),
    "isdqn_stage_batch": (C.c_int, [_P, _P, _P, _I64, _P, _P, _P, _P]),
    "isdqn_read_async": (C.c_int, [_P, _P, _I64, _P, _P]),
    "isdqn_write_async": (C.c_int, [_P, _P, _I64, _P]),
    "isdqn_argmax_heads": (C.c_int, [_P, _I32, _I32, _P, _P]),
    "isdqn_act_workspace_bytes": (_I64, [C.POINTER(Net)]),
    "isdqn_act": (C.c_int, [C.POINTER(Net), _P, _P, _P, _P, _P, _I64, _P]),
    "isdqn_act_host": (C.c_int, [C.POINTER(Net), _P, _P, _P, _I64, _P, _P, _P, _P, _I64, _P, _P]),
    "isdqn_act_mapped": (C.c_int, [C.POINTER(Net), _P, _P, _P, _P, _P, _I32, _P, _I64, _P, _I64]),
    "isdqn_act_wait": (C.c_int, [_P, _I32, _I64]),
    "isdqn_threefry2x32": (None, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _P]),
    "isdqn_threefry_randint": (_I32, [C.c_uint32, C.c_uint32, _I32, _I32]),
    "isdqn_threefry_split": (None, [C.c_uint32, C.c_uint32, _I32, _P]),
    "isdqn_threefry_uniform": (C.c_float, [C.c_uint32, C.c_uint32]),
    "isdqn_dp_unique_id": (C.c_int, [_P]),
    "isdqn_dp_init": (C.c_int, [_P, _I32, _I32, C.POINTER(C.c_void_p)]),
    "isdqn_dp_allreduce_f32": (C.c_int, [_P, _P, _I64, _P]),
    "isdqn_dp_destroy": (C.c_int, [_P]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Loads the shared library (no GPU needed for this) and binds every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IsdqnNativeError(
            f"{LIB_PATH} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "isdqn_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover
            raise IsdqnNativeError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.isdqn_abi_version() != ABI_VERSION:
        raise IsdqnNativeError(f"ABI mismatch: library {lib.isdqn_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    lib = load()
    msg = lib.isdqn_strerror(rc).decode()
    detail = lib.isdqn_last_cuda_error().decode() if rc in (-3, -5) else ""
    raise IsdqnNativeError(f"{what or 'isdqn call'} failed: {msg} ({rc}) {detail}".strip())


def profile(fn, spin_us: int = 300):
    """Runs fn() with the library's per-launch event marks on; returns [(kernel name, ms)] in launch order."""
    lib = load()
    stream = stream_ptr()
    check(lib.isdqn_spin(stream, spin_us), "isdqn_spin")
    check(lib.isdqn_profile_begin(), "isdqn_profile_begin")
    try:
        fn()
    finally:
        names = C.create_string_buffer(512 * 48)
        ms = (C.c_float * 512)()
        n = lib.isdqn_profile_end(stream, 512, names, 48, ms)
    if n < 0:
        check(n, "isdqn_profile_end")
    return [(names.raw[i * 48 : (i + 1) * 48].split(b"\0", 1)[0].decode(), float(ms[i])) for i in range(n)]


def require_cuda():
    """Returns torch after making sure a CUDA device is usable; raises loudly otherwise."""
    import torch

    if not torch.cuda.is_available():
        raise IsdqnNativeError("isdqn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    load()
    return torch


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------- packed host batches
BATCH_FIELDS = ("state", "next_state", "action", "reward", "terminal")


def batch_pack_layout(batch: int, state_bytes_per_row: int):
    """One batch as ONE block: state | next_state | action int64 | reward float64 | terminal uint8, every field on a
    16-byte boundary.  Returns (total bytes, {field: (offset, nbytes)}).  The learner's device batch buffers, its staging
    slots and the pinned blocks `ReplayBuffer.sample()` returns all use this layout, so a sampled batch reaches the
    learner with a single host -> device copy."""
    offs, o = {}, 0
    for name, nb in zip(BATCH_FIELDS, (batch * state_bytes_per_row, batch * state_bytes_per_row, 8 * batch, 8 * batch, batch)):
        offs[name] = (o, nb)
        o = (o + nb + 15) // 16 * 16
    return o, offs


_PINNED_BLOCKS = {}  # base address -> (tensor, nbytes): pinned blocks handed out by pinned_block()


def pinned_block(nbytes: int):
    """A zeroed pinned uint8 tensor, remembered so that arrays living inside it can be recognised later."""
    torch = require_cuda()
    t = torch.zeros(max(int(nbytes), 16), dtype=torch.uint8).pin_memory()
    _PINNED_BLOCKS[t.data_ptr()] = (t, t.numel())
    return t


def pinned_pack_base(arrays, offs) -> Optional[int]:
    """Base address of the registered pinned block that holds the five numpy `arrays` (BATCH_FIELDS order) exactly at
    the offsets of `offs`, or None."""
    try:
        a0 = arrays[0]
        base = a0.ctypes.data - offs[BATCH_FIELDS[0]][0]
        blk = _PINNED_BLOCKS.get(base)
        if blk is None:
            return None
        for name, a in zip(BATCH_FIELDS, arrays):
            off, nb = offs[name]
            if a.ctypes.data != base + off or a.nbytes != nb or not a.flags.c_contiguous:
                return None
        return base if off + nb <= blk[1] else None
    except AttributeError:
        return None


class PinnedStager:
    """Small host arrays -> device with ONE asynchronous copy from pinned memory (instead of one synchronous pageable copy
    per array): a ring of `slots` (pinned block, device block) pairs.  The device addresses / views a call returns stay
    valid until the same slot comes round again (`slots` - 1 further calls); consumers must be enqueued on the stream the
    call was made on.  `put_ptrs` is the lean form (a handful of ctypes calls, no tensor objects) for per-step paths."""

    def __init__(self, nbytes: int = 1 << 16, slots: int = 4):
        self._torch = require_cuda()
        self._lib = load()
        self._n = 0
        self._slots = slots
        self._i = 0
        self._host, self._dev, self._ev = [], [], []
        self._grow(nbytes)

    def _grow(self, nbytes: int) -> None:
        t = self._torch
        for ev in self._ev:
            check(self._lib.isdqn_event_synchronize(ev), "isdqn_event_synchronize")
        self._n = max(int(nbytes), 2 * self._n)
        self._host = [t.empty(self._n, dtype=t.uint8).pin_memory() for _ in range(self._slots)]
        self._host_np = [h.numpy() for h in self._host]
        self._host_ptr = [h.data_ptr() for h in self._host]
        self._dev = [t.empty(self._n, dtype=t.uint8, device="cuda") for _ in range(self._slots)]
        self._dev_ptr = [d.data_ptr() for d in self._dev]
        if not self._ev:
            for _ in range(self._slots):
                ev = C.c_void_p()
                check(self._lib.isdqn_event_create(ev), "isdqn_event_create")
                self._ev.append(ev)

    def _stage(self, arrays):
        offs, o = [], 0
        for a in arrays:
            offs.append(o)
            o = (o + a.nbytes + 15) // 16 * 16
        if o > self._n:
            self._grow(o)
        s = self._i
        self._i = (s + 1) % self._slots
        lib = self._lib
        lib.isdqn_event_synchronize(self._ev[s])  # the previous copy out of this pinned block has completed
        h = self._host_np[s]
        for a, off in zip(arrays, offs):
            h[off : off + a.nbytes] = a.reshape(-1).view(np.uint8) if a.flags.c_contiguous else np.ascontiguousarray(a).reshape(-1).view(np.uint8)
        stream = stream_ptr()
        check(lib.isdqn_write_async(self._dev_ptr[s], self._host_ptr[s], o, stream), "isdqn_write_async")
        check(lib.isdqn_event_record(self._ev[s], stream), "isdqn_event_record")
        return s, offs

    def put_ptrs(self, *arrays):
        """Device addresses (ints) of the staged copies of `arrays`."""
        s, offs = self._stage(arrays)
        base = self._dev_ptr[s]
        return [base + off for off in offs]

    def put(self, *arrays):
        s, offs = self._stage(arrays)
        d = self._dev[s]
        return [d[off : off + a.nbytes].view(_TORCH_DTYPES[a.dtype.str]()) for a, off in zip(arrays, offs)]


def _torch_dtype(name):
    def get():
        import torch

        return getattr(torch, name)

    return get


_TORCH_DTYPES = {
    np.dtype(np.int32).str: _torch_dtype("int32"), np.dtype(np.int64).str: _torch_dtype("int64"),
    np.dtype(np.float64).str: _torch_dtype("float64"), np.dtype(np.float32).str: _torch_dtype("float32"),
    np.dtype(np.uint8).str: _torch_dtype("uint8"),
}
