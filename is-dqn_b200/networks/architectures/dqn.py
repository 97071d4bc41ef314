"""`DQNNet` on the device — mirrors `slimdqn/networks/architectures/dqn.py:7-103` (architecture_type `cnn`, `fc`
and `impala`, optional LayerNorm).  `batch_norm` is out of scope
and raises.

Parameters live in ONE flat float32 CUDA vector (leaves packed in execution order, 16-byte aligned; layout from
`isdqn_net_layout`); the flax-shaped pytree `{"params": {"Conv_0": {"kernel", "bias"}, "LayerNorm_0": {"scale",
"bias"}, ..., "Dense_1": {...}}}` the reference exposes is a tree of VIEWS into that vector, so
`params["params"]["Dense_1"]["kernel"]` reads and writes the memory the kernels use.  impala nests one level more
(`params["params"]["Stack_0"]["Conv_1"]["kernel"]`, flax's naming of the `Stack` sub-module); in the leaf list such a
module is written as the path "Stack_0/Conv_1".
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

from ... import _lib

CONV_GEOMETRY = ((8, 4), (4, 2), (3, 1))  # (kernel, stride) of Conv_0..2 — dqn.py:55,62,69


def _same_out(size: int, s: int) -> int:
    return -(-size // s)


def leaf_specs(arch: str, obs_dim: Sequence[int], features: Sequence[int], n_out: int, layer_norm: bool) -> List[Tuple[str, str, Tuple[int, ...]]]:
    """(module, leaf, shape) in execution order with flax's auto-naming; must agree with isdqn_net_layout."""
    out: List[Tuple[str, str, Tuple[int, ...]]] = []
    ln = 0
    if arch == "cnn":
        h, w, c = obs_dim
        for i, (k, s) in enumerate(CONV_GEOMETRY):
            out += [(f"Conv_{i}", "kernel", (k, k, c, int(features[i]))), (f"Conv_{i}", "bias", (int(features[i]),))]
            if layer_norm:
                out += [(f"LayerNorm_{ln}", "scale", (int(features[i]),)), (f"LayerNorm_{ln}", "bias", (int(features[i]),))]
                ln += 1
            h, w, c = _same_out(h, s), _same_out(w, s), int(features[i])
        fan_in, start = h * w * c, 3
    elif arch == "impala":
        # Stack (dqn.py:7-36): Conv_0, max_pool, 2 x [LayerNorm_j, relu, Conv_{1+2j}, relu, Conv_{2+2j}, + skip]; the
        # DQNNet's own LayerNorm_0 follows the three stacks (dqn.py:84-86)
        h, w, c = obs_dim
        for s in range(3):
            f = int(features[s])
            out += [(f"Stack_{s}/Conv_0", "kernel", (3, 3, c, f)), (f"Stack_{s}/Conv_0", "bias", (f,))]
            for j in range(2):
                if layer_norm:
                    out += [(f"Stack_{s}/LayerNorm_{j}", "scale", (f,)), (f"Stack_{s}/LayerNorm_{j}", "bias", (f,))]
                for q in (1 + 2 * j, 2 + 2 * j):
                    out += [(f"Stack_{s}/Conv_{q}", "kernel", (3, 3, f, f)), (f"Stack_{s}/Conv_{q}", "bias", (f,))]
            h, w, c = _same_out(h, 2), _same_out(w, 2), f
        if layer_norm:
            out += [("LayerNorm_0", "scale", (c,)), ("LayerNorm_0", "bias", (c,))]
            ln = 1
        fan_in, start = h * w * c, 3
    elif arch == "fc":
        fan_in, start = int(np.prod(obs_dim)), 0
    else:
        raise NotImplementedError(f"architecture_type {arch!r} is not a DQNNet architecture (cnn, impala, fc)")
    d = 0
    for i in range(start, len(features)):
        f = int(features[i])
        out += [(f"Dense_{d}", "kernel", (fan_in, f)), (f"Dense_{d}", "bias", (f,))]
        if layer_norm:
            out += [(f"LayerNorm_{ln}", "scale", (f,)), (f"LayerNorm_{ln}", "bias", (f,))]
            ln += 1
        fan_in, d = f, d + 1
    out += [(f"Dense_{d}", "kernel", (fan_in, n_out)), (f"Dense_{d}", "bias", (n_out,))]
    return out


def truncated_standard_normal(g: np.random.Generator, shape) -> np.ndarray:
    """jax.random.truncated_normal(key, -2, 2): draws of N(0, 1) conditioned on |x| <= 2 (rejection sampling — clipping
    would pile 4.6 % of the mass onto the two bounds)."""
    n = int(np.prod(shape))
    out = np.empty(n)
    have = 0
    while have < n:
        v = g.standard_normal(max(int((n - have) * 1.1) + 16, 16))
        v = v[np.abs(v) <= 2.0][: n - have]
        out[have : have + v.size] = v
        have += v.size
    return out.reshape(shape)


class _LeafDict(dict):
    """Module dict whose item assignment writes INTO the flat-vector view instead of rebinding it, so
    `params["params"]["Dense_1"]["bias"] = new_value` (tests/test_isdqn.py:102) reaches the kernels."""

    owner = None  # the ParamTree whose flat vector the views alias

    def __setitem__(self, key, value):
        if key in self:
            import torch

            view = dict.__getitem__(self, key)
            view.copy_(torch.as_tensor(np.asarray(value.cpu() if hasattr(value, "cpu") else value), dtype=view.dtype).reshape(view.shape))
            if self.owner is not None:
                self.owner.shadow_dirty = True
        else:
            dict.__setitem__(self, key, value)


class ParamTree(dict):
    """`{"params": {...}}` pytree of views + the flat vector they alias (`.flat`)."""

    flat = None
    specs = None
    # bf16 copy of `flat` for the tensor-core path (allocated on first use).  Adam keeps it current inside the captured
    # step; anything that writes parameters outside of it (item assignment, shift_params, or in-place torch ops on the
    # views — call mark_dirty() after those) makes the next step rebuild it.
    shadow = None
    shadow_dirty = True

    def mark_dirty(self) -> None:
        self.shadow_dirty = True

    def leaves(self):
        for mod, leaf, _ in self.specs:
            yield mod, leaf, module_at(self["params"], mod)[leaf]


def module_at(inner: dict, mod: str) -> dict:
    """The module dict at the path `mod` ("Dense_0", or "Stack_1/Conv_2" for a nested flax sub-module)."""
    for part in mod.split("/"):
        inner = inner[part]
    return inner


def build_tree(flat, specs, offsets) -> ParamTree:
    tree = ParamTree()
    inner = {}
    for (mod, leaf, shape), off in zip(specs, offsets):
        n = int(np.prod(shape))
        *parents, name = mod.split("/")
        node = inner
        for part in parents:
            node = node.setdefault(part, {})
        d = node.setdefault(name, _LeafDict())
        d.owner = tree
        dict.__setitem__(d, leaf, flat[off : off + n].view(shape))
    tree["params"] = inner
    tree.flat = flat
    tree.specs = specs
    return tree


class DQNNet:
    def __init__(self, features: Sequence[int], architecture_type: str, final_feature: int, layer_norm: bool = False, batch_norm: bool = False):
        if batch_norm:
            raise NotImplementedError("batch_norm=True is out of scope for isdqn_b200 (SURVEY.md §2: needs cross-replica statistics)")
        if architecture_type not in ("cnn", "fc", "impala"):
            raise NotImplementedError(f"architecture_type {architecture_type!r} is not a DQNNet architecture (cnn, impala, fc)")
        self.features = [int(f) for f in features]
        self.architecture_type = architecture_type
        self.final_feature = int(final_feature)
        self.layer_norm = bool(layer_norm)
        self.batch_norm = False
        self._net = None
        self._layout = None
        self._specs = None
        self._ws = {}

    @property
    def image_input(self) -> bool:
        """uint8 (H, W, C) frames, normalised by 255 inside the network (dqn.py:51, 80)"""
        return self.architecture_type in ("cnn", "impala")

    # ------------------------------------------------------------------------------------------------ layout
    def configure(self, observation_dim, n_heads: int, n_actions: int) -> None:
        """Fixes the input shape (flax does this lazily in `init`) and queries the native parameter layout."""
        lib = _lib.load()
        obs = tuple(int(x) for x in observation_dim)
        net = _lib.Net()
        net.arch = {"cnn": _lib.ARCH_CNN, "fc": _lib.ARCH_FC, "impala": _lib.ARCH_IMPALA}[self.architecture_type]
        net.layer_norm = 1 if self.layer_norm else 0
        if self.image_input:
            if len(obs) != 3:
                raise ValueError(f"{self.architecture_type} expects (H, W, C) observations, got {obs}")
            net.obs_h, net.obs_w, net.obs_c = obs
        else:
            net.obs_h, net.obs_w, net.obs_c = 1, 1, int(np.prod(obs))
        if len(self.features) > _lib.MAX_FEATURES:
            raise ValueError("too many feature layers")
        net.n_features = len(self.features)
        for i, f in enumerate(self.features):
            net.features[i] = f
        net.n_heads, net.n_actions = int(n_heads), int(n_actions)
        assert (1 + n_heads) * n_actions == self.final_feature
        layout = _lib.Layout()
        _lib.check(lib.isdqn_net_layout(net, layout), "isdqn_net_layout")
        specs = leaf_specs(self.architecture_type, obs, self.features, self.final_feature, self.layer_norm)
        assert layout.n_leaves == len(specs), (layout.n_leaves, len(specs))
        for i, (_, _, shape) in enumerate(specs):
            assert layout.size[i] == int(np.prod(shape)), (i, specs[i], layout.size[i])
        self.observation_dim = obs
        self._net, self._layout, self._specs = net, layout, specs
        self._offsets = [int(layout.offset[i]) for i in range(layout.n_leaves)]
        self.n_params_padded = int(layout.total)
        self.n_params = sum(int(np.prod(s)) for _, _, s in specs)

    def new_tree(self, fill: float = 0.0) -> ParamTree:
        torch = _lib.require_cuda()
        flat = torch.full((self.n_params_padded,), fill, dtype=torch.float32, device="cuda")
        return build_tree(flat, self._specs, self._offsets)

    def init(self, key, x=None) -> ParamTree:
        """Flax-equivalent initialisers (xavier_uniform for cnn / impala incl. their Dense tails, lecun_normal — a
        normal truncated at two standard deviations — for fc and for the convolutions inside an impala residual block,
        which keep flax's default kernel_init, dqn.py:32-33; zero biases, unit LayerNorm scales — dqn.py:16,49,78,90).
        The random stream is NumPy's, not JAX's threefry: initial VALUES differ from the reference's for the same key
        (outside the parity contract)."""
        torch = _lib.require_cuda()
        seed = np.random.SeedSequence(np.frombuffer(np.asarray(key).tobytes(), dtype=np.uint8).tolist() or [0])
        g = np.random.default_rng(seed)
        tree = self.new_tree()
        host = np.zeros(self.n_params_padded, dtype=np.float32)
        for (mod, leaf, shape), off in zip(self._specs, self._offsets):
            n = int(np.prod(shape))
            if leaf == "kernel":
                rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
                fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
                block_conv = "/" in mod and mod.rsplit("/", 1)[1] != "Conv_0" and mod.rsplit("/", 1)[1].startswith("Conv_")
                if self.architecture_type != "fc" and not block_conv:
                    lim = math.sqrt(6.0 / (fan_in + fan_out))
                    v = g.uniform(-lim, lim, shape)
                else:
                    std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
                    v = truncated_standard_normal(g, shape) * std
            elif leaf == "scale":
                v = np.ones(shape)
            else:
                v = np.zeros(shape)
            host[off : off + n] = np.asarray(v, dtype=np.float32).reshape(-1)
        tree.flat.copy_(torch.from_numpy(host))
        return tree

    # ------------------------------------------------------------------------------------------------- apply
    def _workspace(self, rows: int):
        torch = _lib.require_cuda()
        ws = self._ws.get(rows)
        if ws is None:
            nbytes = _lib.load().isdqn_forward_workspace_bytes(self._net, rows)
            if nbytes < 0:
                raise _lib.IsdqnNativeError("isdqn_forward_workspace_bytes failed")
            ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device="cuda")
            if len(self._ws) > 8:
                self._ws.clear()
            self._ws[rows] = ws
        return ws

    def prepare_input(self, x):
        """-> (contiguous CUDA tensor, rows, input_is_float).  Accepts numpy / torch, uint8 or float, with or
        without the batch axis (`jnp.array(x, ndmin=4)`, dqn.py:51)."""
        torch = _lib.require_cuda()
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
        nd = len(self.observation_dim)
        if t.dim() == nd:
            t = t.unsqueeze(0)
        if tuple(t.shape[1:]) != tuple(self.observation_dim):
            raise ValueError(f"input shape {tuple(t.shape)} does not match observation_dim {self.observation_dim}")
        if self.image_input and t.dtype == torch.uint8:
            is_float = 0
        else:
            t = t.to(torch.float32)
            is_float = 1
        return t.to("cuda").contiguous(), int(t.shape[0]), is_float

    def apply(self, params: ParamTree, x, use_running_average: bool = False, mutable=None):
        """Forward pass; returns float32 CUDA `(N, final_feature)` (squeezed like `jnp.squeeze` for N == 1)."""
        torch = _lib.require_cuda()
        t, rows, is_float = self.prepare_input(x)
        q = torch.empty((rows, self.final_feature), dtype=torch.float32, device="cuda")
        ws = self._workspace(rows)
        _lib.check(
            _lib.load().isdqn_forward(
                self._net, params.flat.data_ptr(), t.data_ptr(), is_float, rows, q.data_ptr(), ws.data_ptr(), ws.numel(),
                _lib.stream_ptr(),
            ),
            "isdqn_forward",
        )
        out = q[0] if rows == 1 else q
        if mutable is not None:
            return out, {}
        return out
