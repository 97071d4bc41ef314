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


# ------------------------------------------------------------------------------------------------------ impala
def _impala_cfg():
    return dict(obs_dim=(21, 18, 3), features=[5, 7, 6, 11], K=2, A=3)


def test_impala_oracle_max_pool_is_same_padded_with_minus_infinity():
    """the pooling of the impala oracle against a brute-force window maximum written with its own index arithmetic
    (flax nn.max_pool padding='SAME': out = ceil(in / 2), pad_lo = total // 2, padded cells never win)"""
    import torch.nn.functional as F

    g = np.random.default_rng(3)
    for H, W in ((84, 84), (42, 42), (21, 21), (37, 50), (5, 4)):
        x = torch.from_numpy(g.standard_normal((2, H, W, 3))) - 5.0  # all negative: a zero-padded pooling would show
        _, lo_h, hi_h = L.same_padding(H, 3, 2)
        _, lo_w, hi_w = L.same_padding(W, 3, 2)
        xp = F.pad(x.permute(0, 3, 1, 2), (lo_w, hi_w, lo_h, hi_h), value=float("-inf"))
        got = F.max_pool2d(xp, 3, 2).permute(0, 2, 3, 1).numpy()
        OH, OW = -(-H // 2), -(-W // 2)
        assert got.shape == (2, OH, OW, 3)
        ph = max((OH - 1) * 2 + 3 - H, 0) // 2
        pw = max((OW - 1) * 2 + 3 - W, 0) // 2
        xn = x.numpy()
        for oy in range(OH):
            for ox in range(OW):
                ys = [y for y in range(oy * 2 - ph, oy * 2 - ph + 3) if 0 <= y < H]
                xs = [c for c in range(ox * 2 - pw, ox * 2 - pw + 3) if 0 <= c < W]
                want = xn[:, ys][:, :, xs].max(axis=(1, 2))
                np.testing.assert_array_equal(got[:, oy, ox], want)


def test_impala_oracle_structure_and_gradient():
    c = _impala_cfg()
    shapes = L.param_shapes("impala", c["obs_dim"], c["features"], (1 + c["K"]) * c["A"], True)
    names = [m for m, _, _ in shapes]
    # flax auto-names: five convolutions and two LayerNorms per Stack, the DQNNet's LayerNorm_0, then the Dense tail
    assert names.count("Stack_1/Conv_4") == 2 and names.count("Stack_2/LayerNorm_1") == 2 and "LayerNorm_0" in names
    assert [s for m, l, s in shapes if m == "Dense_0" and l == "kernel"][0] == (3 * 3 * 6, 11)  # 21x18 -> 11x9 -> 6x5 -> 3x3
    p = L.init_params(5, "impala", c["obs_dim"], c["features"], (1 + c["K"]) * c["A"], True)
    L.randomize_small_leaves(p, 6)
    g = np.random.default_rng(5)
    x = torch.from_numpy(g.integers(0, 256, (2,) + c["obs_dim"], dtype=np.uint8))
    q = L.forward(p, x, "impala", True, 1 + c["K"], c["A"])
    assert q.shape == (2, 1 + c["K"], c["A"]) and torch.isfinite(q).all()
    # a residual block whose second convolution is zero is the identity on its input: zeroing Conv_2 / Conv_4 of every
    # Stack must give the same Q-values as a network whose blocks are skipped altogether
    pz = L.clone_params(p)
    for st in range(3):
        for q_ in (2, 4):
            pz[f"Stack_{st}/Conv_{q_}"]["kernel"].zero_()
            pz[f"Stack_{st}/Conv_{q_}"]["bias"].zero_()
    qz = L.forward(pz, x, "impala", True, 1 + c["K"], c["A"])
    pw = L.clone_params(pz)
    for st in range(3):  # garbage in the (now unused) first convolutions of the blocks must not matter either
        pw[f"Stack_{st}/Conv_1"]["kernel"].add_(3.0)
        pw[f"Stack_{st}/LayerNorm_0"]["scale"].mul_(-2.0)
    assert torch.equal(L.forward(pw, x, "impala", True, 1 + c["K"], c["A"]), qz)
    # autograd through pooling / skips against central differences on a few coordinates
    batch = L.make_batch(9, 2, c["obs_dim"], c["A"], "impala")
    pp = L.clone_params(p)
    _, _, grads, _, _ = L.learn_on_batch(pp, L.zeros_like_params(p), L.zeros_like_params(p), 0, batch, "impala", True, c["K"],
                                         c["A"], 0.99, 1, 0.0, 1.0)
    s_, a_, _, s2_, _ = batch
    targets = L.loss_on_batch(p, batch, "impala", True, c["K"], c["A"], 0.99, 1)[3]

    def frozen_loss(pp):  # the targets carry a stop_gradient (isdqn.py:100): hold them fixed
        all_q = L.forward(pp, torch.cat((s_, s2_)), "impala", True, 1 + c["K"], c["A"])
        q_ = all_q[:2, 1:, :].gather(-1, a_.view(2, 1, 1).expand(2, c["K"], 1)).squeeze(-1)
        return ((q_ - targets) ** 2).mean(0).sum()

    for mod, leaf, idx in (("Stack_0/Conv_0", "kernel", (1, 2, 0, 3)), ("Stack_1/Conv_3", "kernel", (0, 1, 2, 4)),
                           ("Stack_2/LayerNorm_1", "scale", (2,)), ("Stack_0/Conv_2", "bias", (1,)), ("LayerNorm_0", "bias", (4,)),
                           ("Stack_2/Conv_0", "bias", (0,)), ("Stack_1/LayerNorm_0", "bias", (3,))):
        h = 1e-6
        vals = []
        for sgn in (+1, -1):
            q_ = L.clone_params(p)
            q_[mod][leaf][idx] += sgn * h
            with torch.no_grad():
                vals.append(float(frozen_loss(q_)))
        fd = (vals[0] - vals[1]) / (2 * h)
        an = float(grads[mod][leaf][idx])
        assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)) + 1e-7, (mod, leaf, fd, an)
