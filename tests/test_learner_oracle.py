"""CPU: the learner restatement re-asserts the four algebraic properties the reference pins
(reference tests/test_isdqn.py:51-116) plus checks that need no JAX: analytic gradients against finite differences
in float64, optax-0.2.4 Adam against a hand-rolled scalar, flax SAME-padding geometry."""
import numpy as np
import pytest
import torch

from oracle import learner_oracle as L


def small_cfg(seed):
    g = np.random.default_rng(seed)
    A, K = int(g.integers(2, 10)), int(g.integers(1, 10))
    feats = [int(g.integers(5, 20)) for _ in range(4)]
    return A, K, feats


def test_same_padding_geometry():  # 84 -> 21 -> 11 -> 11, conv1 pads (1, 2)
    assert L.same_padding(84, 8, 4) == (21, 2, 2)
    assert L.same_padding(21, 4, 2) == (11, 1, 2)
    assert L.same_padding(11, 3, 1) == (11, 1, 1)
    shapes = L.param_shapes("cnn", (84, 84, 4), [32, 64, 64, 512], 90, True)
    assert sum(int(np.prod(s)) for _, _, s in shapes) == 4_090_938
    assert dict(((m, l), s) for m, l, s in shapes)[("Dense_0", "kernel")] == (7744, 512)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_target_loss_shift_best_action_properties(seed):
    A, K, feats = small_cfg(seed)
    gamma = 0.94
    p = L.init_params(seed, "cnn", (84, 84, 4), feats, (1 + K) * A, True)
    L.randomize_small_leaves(p, seed)
    g = np.random.default_rng(seed)
    # tests/utils.py Generator: float states in [0, 1)
    B = 10
    s = torch.from_numpy(g.random((B, 84, 84, 4)))
    s2 = torch.from_numpy(g.random((B, 84, 84, 4)))
    a = torch.from_numpy(g.integers(0, A, B))
    r = torch.from_numpy(g.random(B))
    d = torch.from_numpy(g.integers(0, 2, B).astype(bool))
    # compute_target (test_isdqn.py:51-63)
    all_q = L.forward(p, s2[:1], "cnn", True, 1 + K, A)[0]
    k = int(g.integers(K))
    t = L.compute_targets(r[:1], d[:1], all_q[None, k : k + 1], gamma, 1)[0, 0]
    assert t == r[0] + (1 - int(d[0])) * gamma * all_q[k].max()
    # loss == its own definition (test_isdqn.py:65-82)
    loss, losses, q2, targets = L.loss_on_batch(p, (s, a, r, s2, d), "cnn", True, K, A, gamma, 1)
    q = torch.stack([q2[b, 1:, a[b]] for b in range(B)])
    want = ((q - targets) ** 2).mean(0).sum()
    assert torch.allclose(loss, want, rtol=1e-12)
    assert losses.shape == (K,)
    # head-0 columns receive exactly zero gradient
    _, _, grads, _, _ = L.learn_on_batch(L.clone_params(p), L.zeros_like_params(p), L.zeros_like_params(p), 0,
                                         (s, a, r, s2, d), "cnn", True, K, A, gamma, 1, 1e-3, 1e-8)
    assert grads["Dense_1"]["kernel"][:, :A].abs().max() == 0
    # shift (test_isdqn.py:99-116)
    p["Dense_1"]["bias"] = torch.arange((1 + K) * A, dtype=torch.float64) / 100
    q_before = L.forward(p, s[:1], "cnn", True, 1 + K, A)[0]
    L.shift_params(p, 1, A)
    q_after = L.forward(p, s[:1], "cnn", True, 1 + K, A)[0]
    assert torch.linalg.norm(q_after[:-1] - q_before[1:]) == 0
    # best_action (test_isdqn.py:84-97)
    assert L.best_action(p, s[0], "cnn", True, K, A, k) == int(torch.argmax(q_after[1 + k]))


def test_gradients_against_finite_differences():
    A, K = 3, 2
    feats = [4, 5, 6, 7]
    p = L.init_params(5, "cnn", (20, 20, 2), feats, (1 + K) * A, True)
    L.randomize_small_leaves(p, 5)
    batch = L.make_batch(5, 3, (20, 20, 2), A, "cnn")
    _, _, grads, _, _ = L.learn_on_batch(L.clone_params(p), L.zeros_like_params(p), L.zeros_like_params(p), 0, batch,
                                         "cnn", True, K, A, 0.9, 1, 0.0, 1e-8)
    g = np.random.default_rng(0)
    for mod, leaf in [("Conv_0", "kernel"), ("LayerNorm_1", "scale"), ("Conv_2", "bias"), ("Dense_0", "kernel"), ("Dense_1", "kernel")]:
        t = p[mod][leaf]
        idx = tuple(int(g.integers(s)) for s in t.shape)
        if mod == "Dense_1":
            idx = (idx[0], A + idx[1] % (K * A))  # an online head column
        h = 1e-6
        old = t[idx].item()
        t[idx] = old + h
        lp = L.loss_on_batch(p, batch, "cnn", True, K, A, 0.9, 1)[0].item()
        t[idx] = old - h
        lm = L.loss_on_batch(p, batch, "cnn", True, K, A, 0.9, 1)[0].item()
        t[idx] = old
        fd = (lp - lm) / (2 * h)
        # stop_gradient on the targets: finite differences see the target path too unless the leaf only feeds
        # Q(s, a); compare where that holds exactly (last layer) and loosely elsewhere is not meaningful -> skip
        if mod == "Dense_1":
            # column of head k>=1 also feeds the target of head k+1 through s': remove that path analytically
            continue
        del fd
    # exact check on a leaf without a target path: freeze targets by hand
    s, a, r, s2, d = batch

    def frozen_loss(pp, targets):
        all_q = L.forward(pp, torch.cat((s, s2)), "cnn", True, 1 + K, A)
        q = all_q[:3, 1:, :].gather(-1, a.view(3, 1, 1).expand(3, K, 1)).squeeze(-1)
        return ((q - targets) ** 2).mean(0).sum()

    targets = L.loss_on_batch(p, batch, "cnn", True, K, A, 0.9, 1)[3]
    for mod, leaf in [("Conv_0", "kernel"), ("Conv_1", "kernel"), ("LayerNorm_1", "scale"), ("Conv_2", "bias"), ("Dense_0", "kernel"), ("LayerNorm_3", "bias"), ("Dense_1", "kernel")]:
        t = p[mod][leaf]
        for _ in range(3):
            idx = tuple(int(g.integers(sh)) for sh in t.shape)
            h = 1e-6
            old = t[idx].item()
            t[idx] = old + h
            lp = frozen_loss(p, targets).item()
            t[idx] = old - h
            lm = frozen_loss(p, targets).item()
            t[idx] = old
            fd = (lp - lm) / (2 * h)
            an = grads[mod][leaf][idx].item()
            assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)) + 1e-7, (mod, leaf, idx, fd, an)


def test_adam_matches_optax_formula():
    p = {"m": {"w": torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)}}
    mu, nu = L.zeros_like_params(p), L.zeros_like_params(p)
    lr, eps, b1, b2 = 6.25e-5, 1.5e-4, 0.9, 0.999
    w = p["m"]["w"].clone()
    m = torch.zeros(3, dtype=torch.float64)
    v = torch.zeros(3, dtype=torch.float64)
    count = 0
    for step in range(1, 6):
        g = torch.tensor([0.1 * step, -0.3, 0.0], dtype=torch.float64)
        count = L.adam_step(p, {"m": {"w": g}}, mu, nu, count, lr, eps)
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        w = w - lr * (m / (1 - b1**step)) / (torch.sqrt(v / (1 - b2**step)) + eps)
        assert torch.allclose(p["m"]["w"], w, rtol=1e-14, atol=0)
    assert count == 5
    assert p["m"]["w"][2] == 0.5  # zero gradient => 0 / (0 + eps) = 0: untouched (head 0 between shifts)


def test_fc_architecture_config1_shapes():
    K, A = 3, 4
    p = L.init_params(0, "fc", (8,), [100, 100], (1 + K) * A, False)
    assert set(p) == {"Dense_0", "Dense_1", "Dense_2"}
    batch = L.make_batch(0, 32, (8,), A, "fc")
    loss, losses, q, _ = L.loss_on_batch(p, batch, "fc", False, K, A, 0.99, 1)
    assert q.shape == (64, 1 + K, A) and losses.shape == (K,) and torch.isfinite(loss)


# ---- second opinions (CPU): the restatement against INDEPENDENT implementations of the same operators.  The learner
# oracle cannot be pinned on the reference itself (jax / flax / optax are not installable here); these
# tests pin each building block on PyTorch's own kernels instead, which share no code with the restatement.
def test_adam_matches_torch_optim_adam():
    """optax.adam(lr, eps) and torch.optim.Adam use the same update (eps outside the square root, bias corrections on
    both moments): five steps on random gradients, including exact zeros."""
    g0 = torch.Generator().manual_seed(5)
    w0 = torch.randn(64, generator=g0, dtype=torch.float64)
    p = {"m": {"w": w0.clone()}}
    mu, nu = L.zeros_like_params(p), L.zeros_like_params(p)
    tw = torch.nn.Parameter(w0.clone())
    opt = torch.optim.Adam([tw], lr=6.25e-5, betas=(0.9, 0.999), eps=1.5e-4)
    count = 0
    for step in range(5):
        g = torch.randn(64, generator=g0, dtype=torch.float64) * 10.0 ** (step - 3)
        g[::7] = 0.0
        count = L.adam_step(p, {"m": {"w": g}}, mu, nu, count, 6.25e-5, 1.5e-4)
        tw.grad = g.clone()
        opt.step()
        assert torch.allclose(p["m"]["w"], tw.detach(), rtol=1e-12, atol=1e-15), step


def test_layer_norm_matches_torch_layer_norm():
    """flax's fast variance E[x^2] - E[x]^2 against F.layer_norm's two-pass statistics, float64."""
    import torch.nn.functional as F

    g0 = torch.Generator().manual_seed(6)
    for shape in ((5, 21, 21, 32), (3, 512), (2, 11, 11, 64)):
        x = torch.randn(shape, generator=g0, dtype=torch.float64) * 3 + 0.7
        scale = torch.randn(shape[-1], generator=g0, dtype=torch.float64)
        bias = torch.randn(shape[-1], generator=g0, dtype=torch.float64)
        got = L.layer_norm_lastdim(x, scale, bias)
        want = F.layer_norm(x, (shape[-1],), scale, bias, eps=1e-6)
        assert torch.allclose(got, want, rtol=1e-10, atol=1e-12)


def test_same_padded_conv_matches_an_unfold_matmul():
    """The oracle's SAME-padded convolution (F.conv2d on an explicitly padded image, HWIO kernel) against a convolution
    written as im2col + matmul with its own index arithmetic (the form the CUDA kernels implement)."""
    g0 = torch.Generator().manual_seed(7)
    x = torch.randint(0, 256, (2, 84, 84, 4), generator=g0, dtype=torch.uint8)
    p = L.init_params(3, "cnn", (84, 84, 4), [8, 16, 16, 32], 6, False)
    taps = []
    L.forward(p, x, "cnn", False, 2, 3, taps=taps)  # taps[i] = pre-ReLU output of conv i (NHWC)
    act = x.double() / 255.0
    for i, (k, s) in enumerate(L.CONV_GEOMETRY):
        w = p[f"Conv_{i}"]["kernel"]  # [kh][kw][cin][cout]
        n, H, W, C = act.shape
        OH, OW = -(-H // s), -(-W // s)
        pad_h = max((OH - 1) * s + k - H, 0)
        pad_w = max((OW - 1) * s + k - W, 0)
        lo_h, lo_w = pad_h // 2, pad_w // 2
        cols = torch.zeros(n, OH, OW, k, k, C, dtype=torch.float64)
        for ky in range(k):
            for kx in range(k):
                for oy in range(OH):
                    iy = oy * s - lo_h + ky
                    if not 0 <= iy < H:
                        continue
                    ix = torch.arange(OW) * s - lo_w + kx
                    ok = (ix >= 0) & (ix < W)
                    cols[:, oy, ok, ky, kx, :] = act[:, iy, ix[ok], :]
        z = cols.reshape(n, OH, OW, k * k * C) @ w.reshape(k * k * C, -1) + p[f"Conv_{i}"]["bias"]
        assert z.shape == taps[i].shape
        assert torch.allclose(z, taps[i], rtol=1e-10, atol=1e-12), i
        act = torch.relu(z)
