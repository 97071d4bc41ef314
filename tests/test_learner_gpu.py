"""GPU: the fp32 CUDA learner against the float64 oracle.  Tolerances (north_star): Q-values, targets, losses within
1e-5 relative (max-norm, tests/learner_utils.rel_err); gradients within 1e-4; parameters after Adam steps within
1e-5.  PARITY UNPINNED w.r.t. the real JAX reference (no jax in this image): the oracle is a restatement."""
import numpy as np
import pytest
import torch

from oracle import learner_oracle as L
from tests.learner_utils import batch_as_element, make_agent, oracle_params_for, push_params, rel_err, tree_to_numpy

pytestmark = pytest.mark.gpu

TOL_Q = 1e-5
TOL_LOSS = 1e-5
TOL_GRAD = 1e-4
# Adam divides by (sqrt(v)+eps): an absolute gradient error e moves the update by up to lr/eps * e (0.42 for the
# Atari recipe), so parameters are compared at 1e-4 of their largest magnitude, not 1e-5.
TOL_PARAM = 1e-4
# fp32 rounding of a pre-activation is ~1e-6: batches used for gradient parity keep every ReLU input of the s half
# at least this far from zero in the float64 oracle (see oracle.learner_oracle.relu_margin).
RELU_MARGIN = 4e-6


def safe_batch(p, seed0, B, cfg, max_tries=200):
    """First batch (seed0, seed0+1000, ...) whose ReLU margin on the float64 oracle is >= RELU_MARGIN."""
    for t in range(max_tries):
        batch = L.make_batch(seed0 + 1000 * t, B, cfg["obs_dim"], cfg["A"], cfg["arch"])
        m = L.relu_margin(p, batch[0], cfg["arch"], cfg["layer_norm"], 1 + cfg["K"], cfg["A"])
        if m >= RELU_MARGIN:
            return batch
    raise AssertionError("no margin-safe batch found")

ATARI = dict(obs_dim=(84, 84, 4), A=9, K=9, features=[32, 64, 64, 512], layer_norm=True, arch="cnn")


def check_step(cfg, B, seed, n_steps=1, check_grads=True, **agent_kw):
    agent = make_agent(seed, **cfg, **agent_kw)
    p = oracle_params_for(agent, seed)
    push_params(agent, p)
    arch, ln, K, A = cfg["arch"], cfg["layer_norm"], cfg["K"], cfg["A"]
    mu, nu, count = L.zeros_like_params(p), L.zeros_like_params(p), 0
    report = {}
    for step in range(n_steps):
        batch = safe_batch(p, seed * 100 + step, B, cfg)
        el = batch_as_element(batch)
        # forward + loss (no update)
        loss, (losses, _) = agent.loss_on_batch(agent.params, el)
        o_loss, o_losses, o_q, o_targets = L.loss_on_batch(p, batch, arch, ln, K, A, agent.gamma, agent.update_horizon)
        e_q = rel_err(agent.last_all_q_values, o_q)
        e_l = rel_err(losses, o_losses)
        assert e_q <= TOL_Q, f"step {step}: Q-values rel err {e_q:.3e}"
        assert e_l <= TOL_LOSS, f"step {step}: losses rel err {e_l:.3e}"
        assert rel_err(loss.reshape(1), o_loss.detach().reshape(1)) <= TOL_LOSS
        # targets recomputed from the product's Q-values must match the oracle's targets
        q_prod = agent.last_all_q_values
        t_prod = agent.compute_target(el, q_prod[B:, :-1].transpose(0, 1)).transpose(0, 1)
        e_t = rel_err(t_prod, o_targets)
        assert e_t <= TOL_Q, f"step {step}: targets rel err {e_t:.3e}"
        if check_grads:
            grads, g_losses = agent.grad_on_batch(agent.params, el)
            pp = L.clone_params(p)
            _, _, o_grads, _, _ = L.learn_on_batch(pp, L.zeros_like_params(p), L.zeros_like_params(p), 0, batch, arch, ln, K, A,
                                                   agent.gamma, agent.update_horizon, 0.0, 1.0)
            gn = tree_to_numpy(grads)
            worst = ("", 0.0)
            for mod in o_grads:
                for leaf in o_grads[mod]:
                    e = rel_err(gn[mod][leaf], o_grads[mod][leaf])
                    if e > worst[1]:
                        worst = (f"{mod}.{leaf}", e)
                    assert e <= TOL_GRAD, f"step {step}: grad {mod}.{leaf} rel err {e:.3e}"
            last = f"Dense_{agent.last_idx_mlp}"
            assert np.abs(gn[last]["kernel"][:, :A]).max() == 0.0 and np.abs(gn[last]["bias"][:A]).max() == 0.0
            report["worst_grad"] = worst
        # the update
        agent.params, agent.optimizer_state, s_losses = agent.learn_on_batch(agent.params, agent.optimizer_state, el)
        count, _, _, _, _ = L.learn_on_batch(p, mu, nu, count, batch, arch, ln, K, A, agent.gamma, agent.update_horizon,
                                             agent.learning_rate, agent.adam_eps)
        assert rel_err(s_losses, o_losses) <= TOL_LOSS
        pn = tree_to_numpy(agent.params)
        for mod in p:
            for leaf in p[mod]:
                e = rel_err(pn[mod][leaf], p[mod][leaf])
                assert e <= TOL_PARAM, f"step {step}: param {mod}.{leaf} rel err {e:.3e}"
        assert int(agent.optimizer_state["count"].item()) == count
        report["q"], report["loss"], report["target"] = e_q, e_l, e_t
    mn = tree_to_numpy(agent.optimizer_state["mu"])
    vn = tree_to_numpy(agent.optimizer_state["nu"])
    for mod in p:
        for leaf in p[mod]:
            assert rel_err(mn[mod][leaf], mu[mod][leaf]) <= 1e-4, f"mu {mod}.{leaf}"
            assert rel_err(vn[mod][leaf], nu[mod][leaf]) <= 2e-4, f"nu {mod}.{leaf}"
    print("parity report", cfg["arch"], cfg["features"], "B", B, report)
    return agent


def test_atari_k9_batch32_three_steps():
    """BASELINE config 2: K=9 Nature-CNN + LayerNorm, 84x84x4 uint8, batch 32 (steps 2+ run from the CUDA graph)."""
    check_step(ATARI, 32, seed=1, n_steps=3)


def test_graph_and_direct_launch_agree_bitwise():
    outs = []
    for use_graph in (True, False):
        agent = make_agent(3, **ATARI, use_cuda_graph=use_graph)
        push_params(agent, oracle_params_for(agent, 3))
        for step in range(4):
            el = batch_as_element(L.make_batch(500 + step, 32, ATARI["obs_dim"], 9, "cnn"))
            agent.learn_on_batch(agent.params, agent.optimizer_state, el)
        outs.append(agent.params.flat.cpu().numpy())
    assert outs[0].tobytes() == outs[1].tobytes()


def test_run_to_run_determinism():
    outs = []
    for _ in range(2):
        agent = make_agent(4, **ATARI)
        push_params(agent, oracle_params_for(agent, 4))
        el = batch_as_element(L.make_batch(7, 32, ATARI["obs_dim"], 9, "cnn"))
        grads, losses = agent.grad_on_batch(agent.params, el)
        outs.append((grads.flat.cpu().numpy().tobytes(), losses.cpu().numpy().tobytes()))
    assert outs[0] == outs[1]


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_reference_style_random_small_networks(seed):
    """reference tests/test_isdqn.py: features 5-19, K 1-9, A 2-9, LayerNorm on, float states in [0,1)."""
    g = np.random.default_rng(seed)
    A, K = int(g.integers(2, 10)), int(g.integers(1, 10))
    feats = [int(g.integers(5, 20)) for _ in range(4)]
    cfg = dict(obs_dim=(84, 84, 4), A=A, K=K, features=feats, layer_norm=True, arch="cnn")
    agent = check_step(cfg, 10, seed=seed, n_steps=2, lr=0.001, gamma=0.94, eps=1e-3)
    from isdqn_b200.sample_collection.replay_buffer import ReplayElement

    B = 10
    s = g.random((B, 84, 84, 4)).astype(np.float32)
    s2 = g.random((B, 84, 84, 4)).astype(np.float32)
    a = g.integers(0, A, B).astype(np.int8)
    r = g.random(B).astype(np.float32)
    d = g.integers(0, 2, B)
    samples = ReplayElement(s, a, r, s2, d)
    # test_loss: loss == its own definition through apply_fn
    computed = agent.loss_on_batch(agent.params, samples)[0]
    all_q, _ = agent.network.apply_fn(agent.params, np.concatenate((s, s2)))
    q = torch.stack([all_q[b, 1:, int(a[b])] for b in range(B)])
    targets = torch.stack([agent.compute_target(ReplayElement(None, None, r[b], None, d[b]), all_q[B + b, :-1]) for b in range(B)])
    want = ((q - targets) ** 2).mean(0).sum()
    assert rel_err(computed.reshape(1), want.reshape(1)) <= 1e-6
    # test_compute_target
    k = int(g.integers(K))
    nq = all_q[B, k]
    t = agent.compute_target(ReplayElement(None, None, r[0], None, d[0]), nq)
    assert float(t) == float(np.float32(r[0]) + np.float32((1 - d[0]) * np.float32(agent.gamma)) * nq.max().item()) or \
        abs(float(t) - (r[0] + (1 - d[0]) * agent.gamma * nq.max().item())) < 1e-6
    # test_best_action
    state = s[0]
    for head in range(K):
        qv = agent.network.apply(agent.params, state).reshape(1 + K, A)[head + 1]
        assert int(agent.best_action_of_head(agent.params, state, head).item()) == int(torch.argmax(qv))
    ba = int(agent.best_action(agent.params, state, seed).item())
    assert 0 <= ba < A
    # test_shift_params: shifted[:-1] == q[1:] exactly
    agent.params["params"][f"Dense_{agent.last_idx_mlp}"]["bias"] = np.arange((1 + K) * A) / 100
    qv = agent.network.apply(agent.params, state).reshape(1 + K, A).clone()
    agent.params = agent.shift_params(agent.params)
    shifted = agent.network.apply(agent.params, state).reshape(1 + K, A)
    assert torch.linalg.norm(shifted[:-1] - qv[1:]).item() == 0


def test_config1_lunar_lander_fc_k3():
    """BASELINE config 1 shapes: fc [100, 100], obs (8,), A=4, K=3, batch 32 (with and without LayerNorm)."""
    for ln in (False, True):
        cfg = dict(obs_dim=(8,), A=4, K=3, features=[100, 100], layer_norm=ln, arch="fc")
        check_step(cfg, 32, seed=20 + ln, n_steps=3, lr=3e-4, eps=1e-3)


def test_cnn_without_layer_norm():
    cfg = dict(ATARI, layer_norm=False)
    check_step(cfg, 8, seed=30, n_steps=2)


def test_wider_cnn_config5_shapes_small_batch():
    """BASELINE config 5 widths (x2 and x4) at a small batch: covers the 128/256-channel conv tiles and the
    wide-row LayerNorm backward."""
    check_step(dict(ATARI, features=[64, 128, 128, 1024]), 8, seed=40, n_steps=1)
    check_step(dict(ATARI, features=[128, 256, 256, 2048]), 4, seed=41, n_steps=1)


def test_odd_batch_and_non_square_observation():
    cfg = dict(obs_dim=(50, 37, 3), A=5, K=4, features=[7, 9, 11, 33], layer_norm=True, arch="cnn")
    check_step(cfg, 5, seed=50, n_steps=2, eps=1e-3)


def test_heads_td_loss_kernel_standalone():
    from isdqn_b200 import _lib

    lib = _lib.load()
    g = np.random.default_rng(0)
    for B, K, A in ((32, 9, 9), (300, 3, 4), (4096, 9, 18)):
        q = torch.from_numpy(g.standard_normal((2 * B, 1 + K, A))).float()
        a = torch.from_numpy(g.integers(0, A, B))
        r = torch.from_numpy(g.standard_normal(B))
        d = torch.from_numpy(g.random(B) < 0.3)
        qd, ad, rd, dd = q.cuda().contiguous(), a.cuda(), r.cuda(), d.cuda().to(torch.uint8)  # keep alive
        losses = torch.empty(K, device="cuda")
        dq = torch.empty((B, (1 + K) * A), device="cuda")
        _lib.check(lib.isdqn_heads_td_loss(qd.data_ptr(), ad.data_ptr(), rd.data_ptr(), dd.data_ptr(), 0.99 ** 3, B, B, K, A,
                                           losses.data_ptr(), dq.data_ptr(), _lib.stream_ptr()))
        torch.cuda.synchronize()
        q64 = q.double().requires_grad_(True)
        qs = q64[:B, 1:, :].gather(-1, a.view(B, 1, 1).expand(B, K, 1)).squeeze(-1)
        t = L.compute_targets(r, d, q64[B:, :-1], 0.99, 3).detach()
        per = ((qs - t) ** 2).mean(0)
        per.sum().backward()
        assert rel_err(losses, per.detach()) <= 1e-6
        assert rel_err(dq.reshape(B, 1 + K, A), q64.grad[:B]) <= 1e-6


def test_adam_kernel_standalone():
    from isdqn_b200 import _lib

    lib = _lib.load()
    n = 4 * 1000
    g = torch.Generator().manual_seed(0)
    p = torch.randn(n, generator=g)
    po = {"m": {"w": p.double().clone()}}
    mu, nu = L.zeros_like_params(po), L.zeros_like_params(po)
    pd, md, vd = p.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    count = 0
    for step in range(5):
        gr = torch.randn(n, generator=g) * (10.0 ** float(step - 3))
        gr[::7] = 0.0
        gd = gr.cuda()  # keep alive until the kernel ran
        _lib.check(lib.isdqn_adam_step(pd.data_ptr(), gd.data_ptr(), md.data_ptr(), vd.data_ptr(), cnt.data_ptr(),
                                       6.25e-5, 0.9, 0.999, 1.5e-4, n, _lib.stream_ptr()))
        torch.cuda.synchronize()
        count = L.adam_step(po, {"m": {"w": gr.double()}}, mu, nu, count, 6.25e-5, 1.5e-4)
        assert rel_err(pd, po["m"]["w"]) <= 1e-6
    assert int(cnt.item()) == 5
    assert torch.equal(pd.cpu()[::7], p[::7])  # zero gradient: parameter untouched (head 0 between shifts)


def test_pipelined_host_batches_match_device_batches():
    """Host numpy batches through the double-buffered staging (H2D on the copy stream, losses read one step late)
    give bit-identical parameters and losses to the same batches passed as CUDA tensors with a sync per step."""
    import torch

    cfg = dict(obs_dim=(84, 84, 4), A=6, K=3, features=[32, 64, 64, 512], layer_norm=True, arch="cnn")
    a_host, a_dev = make_agent(11, **cfg), make_agent(11, **cfg)
    p = oracle_params_for(a_host, 11)
    push_params(a_host, p)
    push_params(a_dev, p)
    batches = [L.make_batch(700 + i, 16, cfg["obs_dim"], cfg["A"], "cnn") for i in range(7)]
    handles, dev_losses = [], []
    with torch.cuda.stream(torch.cuda.Stream()):
        for b in batches:
            a_host.learn_on_batch(a_host.params, a_host.optimizer_state, batch_as_element(b))
            handles.append(a_host.losses_to_host_async())
        host_losses = [h.get() for h in handles[-4:]]
    for b in batches:
        el = batch_as_element(b)
        el = type(el)(*[torch.as_tensor(np.asarray(f)).cuda() for f in el])
        _, _, losses = a_dev.learn_on_batch(a_dev.params, a_dev.optimizer_state, el)
        dev_losses.append(losses.cpu().numpy().copy())
    torch.cuda.synchronize()
    assert a_host.params.flat.cpu().numpy().tobytes() == a_dev.params.flat.cpu().numpy().tobytes()
    for got, want in zip(host_losses, dev_losses[-4:]):
        assert got.tobytes() == want.tobytes()


def test_update_online_params_accumulates_losses_on_device():
    """isdqn.py:55-62: `cumulated_losses += losses` — done by the loss kernel of the captured step; learn_on_batch alone
    must not touch the accumulator."""

    class OneBatchReplay:
        def __init__(self, batches):
            self.batches, self.i = batches, 0

        def sample(self):
            b = self.batches[self.i % len(self.batches)]
            self.i += 1
            return batch_as_element(b)

    cfg = dict(obs_dim=(84, 84, 4), A=6, K=3, features=[32, 64, 64, 512], layer_norm=True, arch="cnn")
    a, twin = make_agent(12, **cfg), make_agent(12, **cfg)
    p = oracle_params_for(a, 12)
    push_params(a, p)
    push_params(twin, p)
    batches = [L.make_batch(800 + i, 8, cfg["obs_dim"], cfg["A"], "cnn") for i in range(5)]
    rb = OneBatchReplay(batches)
    want = np.zeros(cfg["K"], dtype=np.float64)
    for i in range(5):
        a.update_online_params(i + 1, rb)
        _, _, losses = twin.learn_on_batch(twin.params, twin.optimizer_state, batch_as_element(batches[i]))
        want += losses.cpu().numpy().astype(np.float64)
    assert np.array_equal(a.cumulated_losses, want)
    assert np.array_equal(twin.cumulated_losses, np.zeros(cfg["K"]))
    a.cumulated_losses = np.zeros(cfg["K"])
    assert np.array_equal(a.cumulated_losses, np.zeros(cfg["K"]))


def test_pinned_ring_batches_take_the_zero_copy_path():
    """`ReplayBuffer(pinned_ring=N).sample()` returns views of a pinned block in the learner's packed layout: learn_on_batch
    recognises it (no host copy) and computes exactly what it computes from ordinary (pageable) copies of the arrays."""
    from isdqn_b200 import _lib
    from isdqn_b200.sample_collection.replay_buffer import ReplayBuffer, ReplayElement, TransitionElement
    from isdqn_b200.sample_collection.samplers import UniformSamplingDistribution

    rb = ReplayBuffer(UniformSamplingDistribution(3), 8, 200, stack_size=4, update_horizon=1, gamma=0.99, pinned_ring=4)
    rng = np.random.default_rng(5)
    for i in range(120):
        rb.add(TransitionElement(rng.integers(0, 256, (84, 84), dtype=np.uint8), int(rng.integers(6)), float(rng.integers(-1, 2)),
                                 bool(i % 37 == 36), False))
    cfg = dict(obs_dim=(84, 84, 4), A=6, K=3, features=[32, 64, 64, 512], layer_norm=True, arch="cnn")
    a, twin = make_agent(13, **cfg), make_agent(13, **cfg)
    p = oracle_params_for(a, 13)
    push_params(a, p)
    push_params(twin, p)
    for _ in range(6):  # more steps than ring slots: blocks are reused
        batch = rb.sample()
        offs = a._context(8)["pack_offs"]
        arrays = [np.asarray(x) for x in (batch.state, batch.next_state, batch.action, batch.reward, batch.is_terminal)]
        assert _lib.pinned_pack_base(arrays, offs) is not None
        copy = ReplayElement(*[np.array(x) for x in batch])
        assert _lib.pinned_pack_base([np.asarray(x) for x in (copy.state, copy.next_state, copy.action, copy.reward, copy.is_terminal)], offs) is None
        _, _, l1 = a.learn_on_batch(a.params, a.optimizer_state, batch)
        _, _, l2 = twin.learn_on_batch(twin.params, twin.optimizer_state, copy)
        assert l1.cpu().numpy().tobytes() == l2.cpu().numpy().tobytes()
    torch.cuda.synchronize()
    assert a.params.flat.cpu().numpy().tobytes() == twin.params.flat.cpu().numpy().tobytes()


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_best_action_graph_path(dtype):
    """Host uint8 observations take the graph-replayed acting path (pinned staging, batch-of-one forward, argmax of every
    head): it must follow the parameters through learner steps and shift_params, and agree with the float forward."""
    cfg = dict(obs_dim=(84, 84, 4), A=9, K=9, features=[32, 64, 64, 512], layer_norm=True, arch="cnn")
    agent = make_agent(21, **cfg, compute_dtype=dtype)
    p = oracle_params_for(agent, 21)
    push_params(agent, p)
    g = np.random.default_rng(21)
    K, A = cfg["K"], cfg["A"]

    def check(tag):
        for trial in range(3):
            obs = g.integers(0, 256, (84, 84, 4), dtype=np.uint8)
            q = agent.network.apply(agent.params, obs).reshape(1 + K, A).cpu().numpy()  # fp32 forward of the same params
            for head in (0, K // 2, K - 1):
                act = agent.best_action_of_head(agent.params, obs, head)
                assert isinstance(act.item(), int) and 0 <= int(act) < A
                row = q[1 + head]
                if dtype == "float32":
                    assert int(act) == int(np.argmax(row)), (tag, trial, head)
                else:  # bf16 forward: the chosen action is a maximiser up to the 2e-2 parity bar
                    assert row[int(act)] >= row.max() - 2e-2 * max(np.abs(q).max(), 1e-6), (tag, trial, head, row, int(act))
            assert 0 <= int(agent.best_action(agent.params, obs, trial).item()) < A

    check("initial")
    for step in range(3):
        el = batch_as_element(L.make_batch(300 + step, 16, cfg["obs_dim"], A, "cnn"))
        agent.params, agent.optimizer_state, _ = agent.learn_on_batch(agent.params, agent.optimizer_state, el)
    check("after learner steps")
    agent.params = agent.shift_params(agent.params)
    check("after shift_params")
