"""ORACLE (test infrastructure): PyTorch-CPU restatement of the reference learner.

PARITY UNPINNED for the numerics: the arithmetic of this part lives in jax==0.4.30 / jaxlib==0.4.30 /
flax==0.10.2 / optax==0.2.4 (reference setup.cfg:15-24), none installable here (no network), and the
reference's own tests (tests/test_isdqn.py:51-116) are self-consistency checks with no golden numbers.
The restatement follows the published semantics of those pinned versions:

  forward      slimdqn/networks/architectures/dqn.py:47-103  (cnn: x/255, Conv SAME + bias, LayerNorm over
               the channel axis eps=1e-6 with the fast variance E[x^2]-E[x]^2 clamped at 0, ReLU, HWC
               flatten, Dense+LN+ReLU, Dense; fc: Dense(+LN)+ReLU ..., Dense)
               dqn.py:7-36, 77-86  (impala: 3 x Stack = Conv3x3, max_pool 3x3/2 SAME (-inf padding), 2 x residual
               block [LN, ReLU, Conv3x3, ReLU, Conv3x3, + skip]; LN, ReLU, HWC flatten, Dense tail.  Modules of a
               Stack are keyed by their path, "Stack_0/Conv_1")
  apply_fn     slimdqn/networks/isdqn.py:39-41   reshape to (N, 1+K, A)
  loss         isdqn.py:92-103   sum_k mean_b (Q_k(s,a) - stopgrad(r + (1-d) gamma^n max_a' Q_{k-1}(s',a')))^2
  target       isdqn.py:105-109  evaluation order r + (((1-d) * gamma^n) * max)
  step         isdqn.py:82-90    grad -> optax.adam(lr, eps) -> apply_updates
  shift        isdqn.py:111-125  kernel[:, :-A] = kernel[:, A:]; bias[:-A] = bias[A:]
  best_action  isdqn.py:127-135  argmax of head 1+idx

`dtype=torch.float64` is the truth the CUDA fp32 path is compared with at 1e-5 (relative to the largest
magnitude of the compared tensor); `torch.float32` is the CPU baseline that `bench.py --impl reference`
times (labelled "port").
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, Dict[str, torch.Tensor]]

CONV_GEOMETRY = ((8, 4), (4, 2), (3, 1))  # (kernel, stride) of Conv_0..2, architectures/dqn.py:55,62,69


def same_padding(size: int, k: int, s: int) -> Tuple[int, int, int]:
    """flax `padding='SAME'`: out=ceil(in/s); total=max((out-1)s+k-in,0); lo=total//2, hi=total-lo."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    lo = total // 2
    return out, lo, total - lo


def param_shapes(arch: str, obs_dim: Sequence[int], features: Sequence[int], n_out: int, layer_norm: bool) -> List[Tuple[str, str, Tuple[int, ...]]]:
    """(module, leaf, shape) in execution order with flax's auto-naming (Conv_i, LayerNorm_i, Dense_i)."""
    out: List[Tuple[str, str, Tuple[int, ...]]] = []
    ln = 0
    if arch == "cnn":
        h, w, c = obs_dim
        for i, (k, s) in enumerate(CONV_GEOMETRY):
            out += [(f"Conv_{i}", "kernel", (k, k, c, features[i])), (f"Conv_{i}", "bias", (features[i],))]
            if layer_norm:
                out += [(f"LayerNorm_{ln}", "scale", (features[i],)), (f"LayerNorm_{ln}", "bias", (features[i],))]
                ln += 1
            h, w, c = same_padding(h, k, s)[0], same_padding(w, k, s)[0], features[i]
        fan_in, start = h * w * c, 3
    elif arch == "impala":
        h, w, c = obs_dim
        for st in range(3):
            f = features[st]
            out += [(f"Stack_{st}/Conv_0", "kernel", (3, 3, c, f)), (f"Stack_{st}/Conv_0", "bias", (f,))]
            for j in range(2):
                if layer_norm:
                    out += [(f"Stack_{st}/LayerNorm_{j}", "scale", (f,)), (f"Stack_{st}/LayerNorm_{j}", "bias", (f,))]
                for q in (1 + 2 * j, 2 + 2 * j):
                    out += [(f"Stack_{st}/Conv_{q}", "kernel", (3, 3, f, f)), (f"Stack_{st}/Conv_{q}", "bias", (f,))]
            h, w, c = same_padding(h, 3, 2)[0], same_padding(w, 3, 2)[0], f
        if layer_norm:
            out += [("LayerNorm_0", "scale", (c,)), ("LayerNorm_0", "bias", (c,))]
            ln = 1
        fan_in, start = h * w * c, 3
    elif arch == "fc":
        fan_in, start = int(np.prod(obs_dim)), 0
    else:
        raise NotImplementedError(arch)
    d = 0
    for i in range(start, len(features)):
        out += [(f"Dense_{d}", "kernel", (fan_in, features[i])), (f"Dense_{d}", "bias", (features[i],))]
        if layer_norm:
            out += [(f"LayerNorm_{ln}", "scale", (features[i],)), (f"LayerNorm_{ln}", "bias", (features[i],))]
            ln += 1
        fan_in, d = features[i], d + 1
    out += [(f"Dense_{d}", "kernel", (fan_in, n_out)), (f"Dense_{d}", "bias", (n_out,))]
    return out


def init_params(seed: int, arch: str, obs_dim, features, n_out: int, layer_norm: bool, dtype=torch.float64) -> Params:
    """xavier_uniform (cnn) / lecun_normal (fc) kernels, zero biases, unit LN scales.  NOT bit-equal to
    JAX's threefry initialisation: initial values are outside the parity contract."""
    g = np.random.default_rng(seed)
    params: Params = {}
    for mod, leaf, shape in param_shapes(arch, obs_dim, features, n_out, layer_norm):
        if leaf == "kernel":
            rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
            fi, fo = shape[-2] * rf, shape[-1] * rf
            block_conv = "/" in mod and not mod.endswith("/Conv_0")  # flax's default lecun_normal (dqn.py:32-33)
            if arch != "fc" and not block_conv:
                lim = math.sqrt(6.0 / (fi + fo))
                v = g.uniform(-lim, lim, shape)
            else:
                std = math.sqrt(1.0 / fi) / 0.87962566103423978
                v = np.clip(g.standard_normal(shape), -2, 2) * std
        elif leaf == "scale":
            v = np.ones(shape)
        else:
            v = np.zeros(shape)
        params.setdefault(mod, {})[leaf] = torch.tensor(v, dtype=dtype)
    return params


def randomize_small_leaves(params: Params, seed: int) -> None:
    """Biases / LN scales are 0 / 1 at init, which hides indexing bugs: perturb them for parity tests."""
    g = torch.Generator().manual_seed(seed)
    for mod in params.values():
        for leaf, t in mod.items():
            if leaf != "kernel":
                t.add_(0.2 * torch.randn(t.shape, generator=g, dtype=torch.float64).to(t.dtype))


def layer_norm_lastdim(x: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    mean = x.mean(-1, keepdim=True)
    mean2 = (x * x).mean(-1, keepdim=True)
    var = torch.clamp(mean2 - mean * mean, min=0.0)
    return (x - mean) * (torch.rsqrt(var + eps) * scale) + bias


class _RoundGradBf16(torch.autograd.Function):
    """identity forward; rounds the incoming gradient to bfloat16 (the CUDA bf16 path feeds bf16 dz to its GEMMs)"""

    @staticmethod
    def forward(ctx, x):
        return x

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _bf16_ste(x: torch.Tensor) -> torch.Tensor:
    """round to bfloat16 in the forward pass, identity in the backward pass"""
    return x + (x.to(torch.bfloat16).to(x.dtype) - x).detach()


def forward(params: Params, x: torch.Tensor, arch: str, layer_norm: bool, n_heads_total: int, n_actions: int, taps=None,
            emulate_bf16: bool = False) -> torch.Tensor:
    """DQNNet.__call__ + iSDQN.apply reshape.  x: (N,H,W,C) uint8/float for cnn, (N,D) for fc.
    `taps`: optional list that receives every pre-ReLU tensor (used by `relu_margin`).
    `emulate_bf16`: place bfloat16 roundings exactly where the CUDA tensor-core path has them (GEMM operands:
    normalised input, conv / hidden-Dense weights, post-ReLU activations, dz) while computing in this dtype —
    the tight checker for the bf16 kernels; the 2e-2 bar itself is measured against the unrounded oracle."""
    dtype = params["Dense_0"]["kernel"].dtype
    rnd = _bf16_ste if emulate_bf16 else (lambda t: t)
    ln = 0
    d = 0
    if arch == "cnn":
        # (the CUDA bf16 path feeds the exact integer pixel values to the tensor cores and applies 1/255 to the fp32
        #  accumulator, so the normalised input carries no bf16 rounding)
        x = x.to(dtype) / 255.0
        for i, (k, s) in enumerate(CONV_GEOMETRY):
            w = rnd(params[f"Conv_{i}"]["kernel"])  # HWIO
            _, plo_h, phi_h = same_padding(x.shape[1], k, s)
            _, plo_w, phi_w = same_padding(x.shape[2], k, s)
            xc = F.pad(x.permute(0, 3, 1, 2), (plo_w, phi_w, plo_h, phi_h))
            y = F.conv2d(xc, w.permute(3, 2, 0, 1), stride=s)
            if emulate_bf16:
                y = _RoundGradBf16.apply(y)
            y = y + params[f"Conv_{i}"]["bias"].view(1, -1, 1, 1)
            x = y.permute(0, 2, 3, 1)
            if layer_norm:
                x = layer_norm_lastdim(x, params[f"LayerNorm_{ln}"]["scale"], params[f"LayerNorm_{ln}"]["bias"])
                ln += 1
            if taps is not None:
                taps.append(x)
            x = rnd(torch.relu(x))
        x = x.reshape(x.shape[0], -1)
    elif arch == "impala":
        def conv3(x, mod):  # nn.Conv(features, (3, 3)): stride 1, padding SAME = 1 on every side
            # emulate_bf16: the convolutions run on the tensor cores with bf16 input, kernel and output gradient, and their
            # epilogues emit bf16 (the frames of the very first one are exact integers, the 1/255 meets the fp32 accumulator)
            on_tc = emulate_bf16
            w = params[mod]["kernel"]
            if on_tc:
                w = rnd(w)
                if mod != "Stack_0/Conv_0":
                    x = rnd(x)
            y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=1)
            if on_tc:
                y = _RoundGradBf16.apply(y)
            y = (y + params[mod]["bias"].view(1, -1, 1, 1)).permute(0, 2, 3, 1)
            return rnd(y) if on_tc else y

        x = x.to(dtype) / 255.0
        for st in range(3):
            x = conv3(x, f"Stack_{st}/Conv_0")
            _, plo_h, phi_h = same_padding(x.shape[1], 3, 2)
            _, plo_w, phi_w = same_padding(x.shape[2], 3, 2)
            xp = F.pad(x.permute(0, 3, 1, 2), (plo_w, phi_w, plo_h, phi_h), value=float("-inf"))
            x = F.max_pool2d(xp, 3, 2).permute(0, 2, 3, 1)
            for j in range(2):
                block_input = x
                if layer_norm:
                    m = params[f"Stack_{st}/LayerNorm_{j}"]
                    x = layer_norm_lastdim(x, m["scale"], m["bias"])
                if taps is not None:
                    taps.append(x)
                x = torch.relu(x)
                x = conv3(x, f"Stack_{st}/Conv_{1 + 2 * j}")
                if taps is not None:
                    taps.append(x)
                x = torch.relu(x)
                x = conv3(x, f"Stack_{st}/Conv_{2 + 2 * j}") + block_input
        if layer_norm:
            x = layer_norm_lastdim(x, params["LayerNorm_0"]["scale"], params["LayerNorm_0"]["bias"])
            ln = 1
        if taps is not None:
            taps.append(x)
        x = rnd(torch.relu(x)).reshape(x.shape[0], -1)  # (the hidden Dense layers read a bf16 copy)
    else:
        x = x.to(dtype)
    n_dense = sum(1 for m in params if m.startswith("Dense_"))
    for d in range(n_dense - 1):
        y = x @ rnd(params[f"Dense_{d}"]["kernel"])
        if emulate_bf16:
            y = _RoundGradBf16.apply(y)
        x = y + params[f"Dense_{d}"]["bias"]
        if layer_norm:
            x = layer_norm_lastdim(x, params[f"LayerNorm_{ln}"]["scale"], params[f"LayerNorm_{ln}"]["bias"])
            ln += 1
        if taps is not None:
            taps.append(x)
        x = torch.relu(x)
    last = params[f"Dense_{n_dense - 1}"]
    q = x @ last["kernel"] + last["bias"]
    return q.reshape(-1, n_heads_total, n_actions)


def compute_targets(reward, is_terminal, next_q, gamma: float, horizon: int) -> torch.Tensor:
    """next_q: (B, K, A) = all_q[B:, :-1].  r + (((1-d) * gamma**n) * max_a)."""
    mx = next_q.max(dim=-1).values
    coef = (1 - is_terminal.to(torch.int32)).to(next_q.dtype) * (gamma**horizon)
    return reward.to(next_q.dtype).unsqueeze(-1) + coef.unsqueeze(-1) * mx


def loss_on_batch(params: Params, batch, arch: str, layer_norm: bool, K: int, A: int, gamma: float, horizon: int,
                  emulate_bf16: bool = False):
    """Returns (scalar loss, per-head losses (K,), all_q (2B,1+K,A), targets (B,K))."""
    state, action, reward, next_state, terminal = batch
    B = state.shape[0]
    all_q = forward(params, torch.cat((state, next_state)), arch, layer_norm, 1 + K, A, emulate_bf16=emulate_bf16)
    q = all_q[:B, 1:, :].gather(-1, action.long().view(B, 1, 1).expand(B, K, 1)).squeeze(-1)
    targets = compute_targets(reward, terminal, all_q[B:, :-1], gamma, horizon).detach()
    td = (q - targets) ** 2
    losses = td.mean(dim=0)
    return losses.sum(), losses, all_q, targets


def adam_step(params: Params, grads: Params, mu: Params, nu: Params, count: int, lr: float, eps: float, b1=0.9, b2=0.999) -> int:
    """optax 0.2.4 adam, in place; returns the new count."""
    count += 1
    c1 = 1.0 - b1**count
    c2 = 1.0 - b2**count
    for mod in params:
        for leaf in params[mod]:
            g = grads[mod][leaf]
            m = mu[mod][leaf]
            v = nu[mod][leaf]
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).add_(g * g, alpha=1 - b2)
            params[mod][leaf].add_(-lr * (m / c1) / (torch.sqrt(v / c2) + eps))
    return count


def zeros_like_params(params: Params) -> Params:
    return {m: {k: torch.zeros_like(v) for k, v in leaves.items()} for m, leaves in params.items()}


def clone_params(params: Params, dtype=None) -> Params:
    return {m: {k: v.detach().clone().to(dtype or v.dtype) for k, v in leaves.items()} for m, leaves in params.items()}


def learn_on_batch(params: Params, mu: Params, nu: Params, count: int, batch, arch, layer_norm, K, A, gamma, horizon, lr, eps,
                   emulate_bf16: bool = False):
    """One reference learner step, in place.  Returns (count, losses(K), grads, all_q, targets)."""
    leaves = [v for mod in params.values() for v in mod.values()]
    for v in leaves:
        v.requires_grad_(True)
        v.grad = None
    loss, losses, all_q, targets = loss_on_batch(params, batch, arch, layer_norm, K, A, gamma, horizon, emulate_bf16)
    loss.backward()
    grads = {m: {k: v.grad.detach().clone() for k, v in lv.items()} for m, lv in params.items()}
    for v in leaves:
        v.requires_grad_(False)
        v.grad = None
    with torch.no_grad():
        count = adam_step(params, grads, mu, nu, count, lr, eps)
    return count, losses.detach(), grads, all_q.detach(), targets


def shift_params(params: Params, last_idx_mlp: int, A: int) -> None:
    k = params[f"Dense_{last_idx_mlp}"]["kernel"]
    b = params[f"Dense_{last_idx_mlp}"]["bias"]
    k[:, :-A] = k[:, A:].clone()
    b[:-A] = b[A:].clone()


def best_action(params: Params, state, arch, layer_norm, K, A, idx_network: int) -> int:
    q = forward(params, state.unsqueeze(0), arch, layer_norm, 1 + K, A)[0]
    return int(torch.argmax(q[1 + idx_network]))


def relu_margin(params: Params, state: torch.Tensor, arch: str, layer_norm: bool, n_heads_total: int, n_actions: int) -> float:
    """min |pre-ReLU value| over every unit of the forward pass on `state`.  A unit closer to zero than the fp32
    rounding of its pre-activation (~1e-6) can take the other ReLU branch in fp32 than in float64, which changes
    the GRADIENT discontinuously (one flipped unit moved Conv_0.kernel's gradient by 7e-3 in a probe, identically
    for torch-CPU fp32 and the CUDA path).  Parity tests therefore use batches whose margin is comfortably
    larger than that rounding; Q-values/targets/losses are continuous and need no such care."""
    taps: list = []
    with torch.no_grad():
        forward(params, state, arch, layer_norm, n_heads_total, n_actions, taps)
    return min(float(t.abs().min()) for t in taps) if taps else float("inf")


def make_batch(seed: int, B: int, obs_dim, A: int, arch: str, p_terminal: float = 0.2):
    """Synthetic batch in the dtypes `rb.sample()` returns (u8 stacks, i64 action, f64 reward, bool terminal)."""
    g = np.random.default_rng(seed)
    if arch in ("cnn", "impala"):
        s = g.integers(0, 256, (B,) + tuple(obs_dim), dtype=np.uint8)
        s2 = g.integers(0, 256, (B,) + tuple(obs_dim), dtype=np.uint8)
    else:
        s = g.uniform(-1, 1, (B,) + tuple(obs_dim)).astype(np.float32)
        s2 = g.uniform(-1, 1, (B,) + tuple(obs_dim)).astype(np.float32)
    a = g.integers(0, A, B).astype(np.int64)
    r = g.integers(-1, 2, B).astype(np.float64)
    d = g.random(B) < p_terminal
    return tuple(torch.from_numpy(x) for x in (s, a, r, s2, d))
