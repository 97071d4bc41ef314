"""GPU: architecture_type "impala" (slimdqn/networks/architectures/dqn.py:7-36, 77-86) on the fp32 path
against the float64 oracle, with the bars of tests/test_learner_gpu.py (Q-values / targets / losses 1e-5, gradients
1e-4, parameters after Adam 1e-4).  PARITY UNPINNED w.r.t. the real JAX reference like the rest of the learner."""
import numpy as np
import pytest
import torch

from oracle import learner_oracle as L
from tests.learner_utils import batch_as_element, make_agent, oracle_params_for, push_params, rel_err, tree_to_numpy
from tests.test_learner_gpu import TOL_LOSS, TOL_Q, check_step

pytestmark = pytest.mark.gpu

# gradient parity needs a batch whose every ReLU input keeps a margin from zero in the oracle (test_learner_gpu.py):
# with ~6 M ReLU units per batch at 84 x 84 x batch 32 no such batch exists, so the strict step check runs on 42 x 42
# frames and the full-size check bounds the gradients in L2 instead
IMPALA_42 = dict(obs_dim=(42, 42, 4), A=6, K=4, features=[16, 32, 32, 256], layer_norm=True, arch="impala")
IMPALA_84 = dict(obs_dim=(84, 84, 4), A=9, K=9, features=[16, 32, 32, 512], layer_norm=True, arch="impala")


def test_impala_layer_norm_three_steps():
    check_step(IMPALA_42, 4, seed=61, n_steps=3)


def test_impala_without_layer_norm():
    check_step(dict(IMPALA_42, layer_norm=False), 4, seed=62, n_steps=2)


def test_impala_odd_shapes_and_channel_counts():
    """non-square frames, odd sizes through the three poolings (37 -> 19 -> 10 -> 5, 50 -> 25 -> 13 -> 7), channel counts
    that are not multiples of 4 or 32, two hidden Dense layers"""
    cfg = dict(obs_dim=(37, 50, 3), A=5, K=3, features=[7, 12, 33, 40, 24], layer_norm=True, arch="impala")
    check_step(cfg, 3, seed=63, n_steps=2, eps=1e-3)


def test_impala_atari_shape_batch32():
    """84 x 84 x 4, batch 32, K = 9: Q-values / targets / losses at 1e-5; per-leaf gradients within 2e-3 in L2 (single
    ReLU units within fp32 rounding of zero take the other branch than in float64, see the module docstring); the
    learner step itself runs (graph replay from step 2) and is deterministic"""
    cfg = IMPALA_84
    agent = make_agent(71, **cfg)
    p = oracle_params_for(agent, 71)
    push_params(agent, p)
    B = 32
    batch = L.make_batch(7100, B, cfg["obs_dim"], cfg["A"], "impala")
    el = batch_as_element(batch)
    loss, (losses, _) = agent.loss_on_batch(agent.params, el)
    _, o_losses, o_q, _ = L.loss_on_batch(p, batch, "impala", True, cfg["K"], cfg["A"], agent.gamma, agent.update_horizon)
    assert rel_err(agent.last_all_q_values, o_q) <= TOL_Q
    assert rel_err(losses, o_losses) <= TOL_LOSS
    grads, _ = agent.grad_on_batch(agent.params, el)
    raw = grads.flat.cpu().numpy().tobytes()
    pp = L.clone_params(p)
    _, _, o_grads, _, _ = L.learn_on_batch(pp, L.zeros_like_params(p), L.zeros_like_params(p), 0, batch, "impala", True, cfg["K"],
                                           cfg["A"], agent.gamma, agent.update_horizon, 0.0, 1.0)
    gn = tree_to_numpy(grads)
    for mod in o_grads:
        for leaf in o_grads[mod]:
            w = o_grads[mod][leaf].numpy()
            e = np.linalg.norm(gn[mod][leaf] - w) / max(np.linalg.norm(w), 1e-30)
            assert e <= 2e-3, f"grad {mod}.{leaf}: L2 rel err {e:.3e}"
    grads2, _ = agent.grad_on_batch(agent.params, el)
    assert grads2.flat.cpu().numpy().tobytes() == raw  # run-to-run bit-identical
    before = agent.params.flat.clone()
    for _ in range(3):
        agent.params, agent.optimizer_state, s_losses = agent.learn_on_batch(agent.params, agent.optimizer_state, el)
    assert int(agent.optimizer_state["count"].item()) == 3
    assert torch.isfinite(s_losses).all() and not torch.equal(before, agent.params.flat)


def test_impala_acting_checkpoint_and_shift(tmp_path):
    cfg = dict(IMPALA_42, K=5)
    agent = make_agent(81, **cfg)
    p = oracle_params_for(agent, 81)
    push_params(agent, p)
    K, A = cfg["K"], cfg["A"]
    g = np.random.default_rng(81)
    state = g.integers(0, 256, cfg["obs_dim"], dtype=np.uint8)
    q = agent.network.apply(agent.params, state).reshape(1 + K, A)
    o_q = L.forward(p, torch.from_numpy(state).unsqueeze(0), "impala", True, 1 + K, A)[0]
    assert rel_err(q, o_q) <= TOL_Q
    for head in range(K):
        assert int(agent.best_action_of_head(agent.params, state, head).item()) == int(torch.argmax(o_q[head + 1]))
    assert 0 <= int(agent.best_action(agent.params, state, 3).item()) < A
    # model file round trip keeps flax's nesting: params/params/Stack_1/Conv_3/kernel
    from isdqn_b200 import checkpoint

    path = tmp_path / "model"
    checkpoint.save_model(agent, path)
    model = checkpoint.load_model_pickle(path)
    assert model["params"]["params"]["Stack_1"]["Conv_3"]["kernel"].shape == (3, 3, 32, 32)
    other = make_agent(82, **cfg)
    other.load_model(str(path))
    assert torch.equal(other.params.flat, agent.params.flat)
    # resume archive
    st = tmp_path / "state.npz"
    batch = L.make_batch(8100, 4, cfg["obs_dim"], A, "impala")
    el = batch_as_element(batch)
    agent.params, agent.optimizer_state, _ = agent.learn_on_batch(agent.params, agent.optimizer_state, el)
    checkpoint.save_agent_state(agent, st)
    checkpoint.load_agent_state(other, st)
    assert torch.equal(other.params.flat, agent.params.flat)
    assert torch.equal(other.optimizer_state["mu"].flat, agent.optimizer_state["mu"].flat)
    # shift_params: head k takes head k+1's place exactly
    qv = agent.network.apply(agent.params, state).reshape(1 + K, A).clone()
    agent.params = agent.shift_params(agent.params)
    shifted = agent.network.apply(agent.params, state).reshape(1 + K, A)
    assert torch.linalg.norm(shifted[:-1] - qv[1:]).item() == 0


def _check_impala_bf16(cfg, B, seed):
    """Q-values / targets / losses: north_star's 2e-2 against the float64 oracle.  Gradients: this network puts 15
    convolutions and 13 ReLU layers between the first kernel and the loss, and bf16 rounding of activations and output
    gradients flips ReLU masks all the way: the oracle with the SAME roundings emulated (emulate_bf16) already deviates from
    the unrounded one by 20-30 % (L2) in the first stack, 5 % in the last, 0.5 % at the head.  A leaf passes if the product
    is as close to the unrounded oracle, and to the emulation, as the emulation is to the unrounded oracle (x 2; 20 % at least: single leaves of a
    few dozen elements scatter).  The layers next to the loss are held to 3e-2."""
    agent = make_agent(seed, **cfg, compute_dtype="bfloat16")
    p = oracle_params_for(agent, seed)
    push_params(agent, p)
    ln, K, A = cfg["layer_norm"], cfg["K"], cfg["A"]
    batch = L.make_batch(seed * 100, B, cfg["obs_dim"], A, "impala")
    el = batch_as_element(batch)
    loss, (losses, _) = agent.loss_on_batch(agent.params, el)
    args = (batch, "impala", ln, K, A, agent.gamma, agent.update_horizon)
    _, o_losses, o_q, o_targets = L.loss_on_batch(p, *args)
    assert rel_err(agent.last_all_q_values, o_q) <= 2e-2
    assert rel_err(losses, o_losses) <= 2e-2
    t_prod = agent.compute_target(el, agent.last_all_q_values[B:, :-1].transpose(0, 1)).transpose(0, 1)
    assert rel_err(t_prod, o_targets) <= 2e-2
    grads, _ = agent.grad_on_batch(agent.params, el)
    gn = tree_to_numpy(grads)
    z = lambda: L.zeros_like_params(p)
    _, _, o_grads, _, _ = L.learn_on_batch(L.clone_params(p), z(), z(), 0, *args, 0.0, 1.0)
    _, _, e_grads, e_q, _ = L.learn_on_batch(L.clone_params(p), z(), z(), 0, *args, 0.0, 1.0, emulate_bf16=True)
    assert rel_err(agent.last_all_q_values, e_q) <= 2e-2

    def l2(g, w):
        w = w.numpy()
        return float(np.linalg.norm(g - w) / max(np.linalg.norm(w), 1e-30))

    report = {}
    for mod in o_grads:
        for leaf in o_grads[mod]:
            vs_f64, vs_emu = l2(gn[mod][leaf], o_grads[mod][leaf]), l2(gn[mod][leaf], e_grads[mod][leaf])
            floor = l2(e_grads[mod][leaf].numpy(), o_grads[mod][leaf])
            report[f"{mod}.{leaf}"] = (round(vs_f64, 4), round(vs_emu, 4), round(floor, 4))
            assert vs_f64 <= max(0.2, 2.0 * floor), f"grad {mod}.{leaf}: {vs_f64:.3e} vs the unrounded oracle (emulation: {floor:.3e})"
            assert vs_emu <= max(0.2, 2.0 * floor), f"grad {mod}.{leaf}: {vs_emu:.3e} vs the bf16-emulating oracle"
    last = f"Dense_{agent.last_idx_mlp}"
    assert report[f"{last}.kernel"][0] <= 3e-2 and report[f"{last}.bias"][0] <= 3e-2
    print("impala bf16 gradient report (vs float64, vs emulation, emulation vs float64):", report)
    # the step itself: losses of the update equal the losses of the forward, parameters move, the shadow follows
    agent.params, agent.optimizer_state, s_losses = agent.learn_on_batch(agent.params, agent.optimizer_state, el)
    assert rel_err(s_losses, o_losses) <= 2e-2
    q_after = agent.network.apply(agent.params, batch[0][:2].numpy())  # fp32 forward on the updated master weights
    loss2, _ = agent.loss_on_batch(agent.params, el)  # bf16 forward on the updated shadow
    assert torch.isfinite(q_after).all() and torch.isfinite(loss2)
    return agent


def test_impala_bf16_tensor_core_convolutions():
    """compute_dtype="bfloat16": every convolution but Stack_0/Conv_0 on the tcgen05 tile engine (bf16 operands, fp32
    accumulation; TMA-fed problems where the shape allows, gathers elsewhere), fp32 residual stream / LayerNorm / Dense
    tail."""
    _check_impala_bf16(dict(IMPALA_84, features=[32, 64, 64, 512]), 32, seed=92)
    _check_impala_bf16(dict(IMPALA_42, features=[32, 32, 64, 128], layer_norm=False), 5, seed=93)
    _check_impala_bf16(dict(IMPALA_42, features=[64, 128, 256, 256]), 3, seed=95)
    # two hidden Dense layers on the engine; 12 input channels: the first convolution stays on the CUDA cores (no padding
    # to one 16-byte chunk) inside the tensor-core mode
    _check_impala_bf16(dict(IMPALA_42, features=[32, 32, 64, 128, 64]), 4, seed=97)
    agent = _check_impala_bf16(dict(IMPALA_42, obs_dim=(42, 42, 12), features=[32, 64, 64, 128]), 4, seed=98)
    # acting goes through the same kernels on one observation (2 rows): it runs and returns a valid action
    state = np.random.default_rng(98).integers(0, 256, (42, 42, 12), dtype=np.uint8)
    for k in range(3):
        assert 0 <= int(agent.best_action(agent.params, state, k).item()) < IMPALA_42["A"]


def test_impala_bf16_gather_and_tma_problems_agree(monkeypatch):
    """ISDQN_IMPALA_TMA=0 keeps every convolution on the gather-fed problems: same numbers up to accumulation order"""
    import subprocess, sys, os

    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np, torch\n"
        "from oracle import learner_oracle as L\n"
        "from tests.learner_utils import batch_as_element, make_agent\n"
        "cfg = dict(obs_dim=(84, 84, 4), A=9, K=9, features=[32, 64, 64, 512], layer_norm=True, arch='impala')\n"
        "agent = make_agent(96, **cfg, compute_dtype='bfloat16')\n"
        "el = batch_as_element(L.make_batch(9600, 16, cfg['obs_dim'], 9, 'impala'))\n"
        "g, losses = agent.grad_on_batch(agent.params, el)\n"
        "np.save(sys.argv[1], np.concatenate((g.flat.cpu().numpy(), losses.cpu().numpy())))\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for tma in ("1", "0"):
        path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"impala_tma_{tma}_{os.getpid()}.npy")
        env = dict(os.environ, ISDQN_IMPALA_TMA=tma)
        r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        outs.append(np.load(path))
        os.remove(path)
    a, b = outs
    assert np.isfinite(a).all() and np.isfinite(b).all()
    assert np.linalg.norm(a - b) / np.linalg.norm(a) <= 2e-2  # (bf16 noise through different accumulation orders)


def test_impala_bf16_learns_from_replay_batches_deterministically():
    cfg = dict(IMPALA_84, features=[32, 64, 64, 512])
    outs = []
    for _ in range(2):
        agent = make_agent(94, **cfg, compute_dtype="bfloat16")
        batch = L.make_batch(9400, 32, cfg["obs_dim"], cfg["A"], "impala")
        el = batch_as_element(batch)
        for _ in range(4):
            agent.params, agent.optimizer_state, losses = agent.learn_on_batch(agent.params, agent.optimizer_state, el)
        assert torch.isfinite(losses).all()
        outs.append(agent.params.flat.cpu().numpy().tobytes())
        q = agent.network.apply(agent.params, batch[0][0].numpy())
        assert torch.isfinite(q).all()
    assert outs[0] == outs[1]


def test_impala_ineligible_widths_refuse_the_tensor_core_dtype():
    from isdqn_b200 import _lib

    agent = make_agent(91, **IMPALA_42, compute_dtype="bfloat16")  # 16 channels: not a tile-engine width
    batch = L.make_batch(9100, 4, IMPALA_42["obs_dim"], IMPALA_42["A"], "impala")
    with pytest.raises(_lib.IsdqnNativeError):
        agent.learn_on_batch(agent.params, agent.optimizer_state, batch_as_element(batch))
