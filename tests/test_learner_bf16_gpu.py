"""GPU: the bf16 tensor-core (tcgen05) learner path against the float64 oracle.  Tolerance (north_star): Q-values,
targets and losses within 2e-2 relative (max-norm).  Gradients are checked per leaf in the relative L2 norm at 3e-2
(bf16 operand rounding is ~4e-3 per element and largely averages out over the reduction)."""
import numpy as np
import pytest
import torch

from oracle import learner_oracle as L
from tests.learner_utils import batch_as_element, make_agent, oracle_params_for, push_params, rel_err, tree_to_numpy

pytestmark = pytest.mark.gpu

TOL = 2e-2
# Against the UNROUNDED float64 oracle the gradients of the earliest layers carry the bf16 noise of four layers of
# ReLU-mask flips (5-9 % in L2 was measured): that comparison is reported, and bounded loosely.  The kernels are
# checked tightly against the oracle with bf16 roundings emulated at the same places (emulate_bf16=True).
TOL_GRAD_L2 = 0.15
TOL_EMU_Q = 2e-3
TOL_EMU_GRAD_L2 = 3e-2
ATARI = dict(obs_dim=(84, 84, 4), A=9, K=9, features=[32, 64, 64, 512], layer_norm=True, arch="cnn")


def l2_rel(got, want):
    g = np.asarray(got, dtype=np.float64)
    w = np.asarray(want.detach().cpu() if isinstance(want, torch.Tensor) else want, dtype=np.float64)
    return float(np.linalg.norm(g - w) / max(np.linalg.norm(w), 1e-30))


def check_bf16(cfg, B, seed, n_steps=2):
    agent = make_agent(seed, **cfg, compute_dtype="bfloat16")
    p = oracle_params_for(agent, seed)
    push_params(agent, p)
    arch, ln, K, A = cfg["arch"], cfg["layer_norm"], cfg["K"], cfg["A"]
    mu, nu, count = L.zeros_like_params(p), L.zeros_like_params(p), 0
    rep = {}
    for step in range(n_steps):
        batch = L.make_batch(seed * 100 + step, B, cfg["obs_dim"], A, arch)
        el = batch_as_element(batch)
        loss, (losses, _) = agent.loss_on_batch(agent.params, el)
        o_loss, o_losses, o_q, o_targets = L.loss_on_batch(p, batch, arch, ln, K, A, agent.gamma, agent.update_horizon)
        rep["q"] = rel_err(agent.last_all_q_values, o_q)
        rep["loss"] = rel_err(losses, o_losses)
        t_prod = agent.compute_target(el, agent.last_all_q_values[B:, :-1].transpose(0, 1)).transpose(0, 1)
        rep["target"] = rel_err(t_prod, o_targets)
        assert rep["q"] <= TOL, f"step {step}: Q-values rel err {rep['q']:.3e}"
        assert rep["target"] <= TOL, f"step {step}: targets rel err {rep['target']:.3e}"
        assert rep["loss"] <= TOL, f"step {step}: losses rel err {rep['loss']:.3e}"
        grads, _ = agent.grad_on_batch(agent.params, el)
        _, _, o_grads, _, _ = L.learn_on_batch(L.clone_params(p), L.zeros_like_params(p), L.zeros_like_params(p), 0, batch, arch, ln,
                                               K, A, agent.gamma, agent.update_horizon, 0.0, 1.0)
        gn = tree_to_numpy(grads)
        worst = ("", 0.0)
        for mod in o_grads:
            for leaf in o_grads[mod]:
                e = l2_rel(gn[mod][leaf], o_grads[mod][leaf])
                if e > worst[1]:
                    worst = (f"{mod}.{leaf}", e)
        rep["worst_grad_l2"] = worst
        assert worst[1] <= TOL_GRAD_L2, f"step {step}: grad {worst[0]} rel L2 err {worst[1]:.3e}"
        # tight check: same roundings emulated in float64
        _, e_losses, e_grads, e_q, _ = L.learn_on_batch(L.clone_params(p), L.zeros_like_params(p), L.zeros_like_params(p), 0, batch,
                                                        arch, ln, K, A, agent.gamma, agent.update_horizon, 0.0, 1.0, emulate_bf16=True)
        rep["emu_q"] = rel_err(agent.last_all_q_values, e_q)
        assert rep["emu_q"] <= TOL_EMU_Q, f"step {step}: Q-values vs bf16-emulating oracle {rep['emu_q']:.3e}"
        emu = {f"{m}.{k}": l2_rel(gn[m][k], e_grads[m][k]) for m in e_grads for k in e_grads[m]}
        rep["emu_grad_l2"] = {k: float(f"{v:.2e}") for k, v in emu.items()}
        bad = {k: v for k, v in emu.items() if v > TOL_EMU_GRAD_L2}
        assert not bad, f"step {step}: grads vs bf16-emulating oracle (rel L2): {rep['emu_grad_l2']}"
        agent.params, agent.optimizer_state, s_losses = agent.learn_on_batch(agent.params, agent.optimizer_state, el)
        count, _, _, _, _ = L.learn_on_batch(p, mu, nu, count, batch, arch, ln, K, A, agent.gamma, agent.update_horizon,
                                             agent.learning_rate, agent.adam_eps)
        assert rel_err(s_losses, o_losses) <= TOL
        # keep both sides on the same trajectory: the comparison is per step, not of accumulated bf16 drift
        push_params(agent, p)
        for tree, src in ((agent.optimizer_state["mu"], mu), (agent.optimizer_state["nu"], nu)):
            push_params(agent, src, tree)
    print("bf16 parity report", cfg["features"], "B", B, rep)
    return agent


def test_atari_k9_batch32_bf16():
    check_bf16(ATARI, 32, seed=1, n_steps=3)


def test_wider_cnn_bf16():
    check_bf16(dict(ATARI, features=[64, 128, 128, 1024]), 8, seed=2, n_steps=1)
    check_bf16(dict(ATARI, features=[128, 256, 256, 2048]), 4, seed=3, n_steps=1)


def test_batch_not_multiple_of_tile_bf16():
    check_bf16(ATARI, 5, seed=4, n_steps=1)
    check_bf16(dict(ATARI, layer_norm=False), 16, seed=5, n_steps=1)


def test_large_batch_bf16():
    # batch 300: split batch axis in the head weight gradient (2 splits, ragged), many-row LayerNorm backward with a
    # ragged tail, several M tiles in every tensor-core problem
    check_bf16(ATARI, 300, seed=7, n_steps=1)


def test_ineligible_network_is_refused():
    from isdqn_b200 import _lib

    agent = make_agent(0, obs_dim=(84, 84, 4), A=4, K=2, features=[7, 9, 11, 33], layer_norm=True, arch="cnn", compute_dtype="bfloat16")
    with pytest.raises(_lib.IsdqnNativeError):
        agent.grad_on_batch(agent.params, batch_as_element(L.make_batch(0, 4, (84, 84, 4), 4, "cnn")))


def test_bf16_graph_replay_is_deterministic():
    outs = []
    for _ in range(2):
        agent = make_agent(6, **ATARI, compute_dtype="bfloat16")
        push_params(agent, oracle_params_for(agent, 6))
        for step in range(4):
            el = batch_as_element(L.make_batch(900 + step, 32, ATARI["obs_dim"], 9, "cnn"))
            agent.learn_on_batch(agent.params, agent.optimizer_state, el)
        outs.append(agent.params.flat.cpu().numpy().tobytes())
    assert outs[0] == outs[1]


@pytest.mark.parametrize("B", [32, 5, 64, 300])
def test_update_is_adam_of_the_reported_gradient_bf16(B):
    """learn_on_batch from a zero optimiser state must leave mu = (1-b1) g, nu = (1-b2) g^2 and the optax step of exactly
    the gradient grad_on_batch reports — for every leaf.  At batch <= 64 the hidden Dense kernel takes the fused
    rank-B gradient + Adam kernel (its gradient never reaches memory), above that the separate launches: both must agree
    with the tensor-core gradient (fp32 accumulation order is the only difference)."""
    agent = make_agent(11, **ATARI, compute_dtype="bfloat16")
    push_params(agent, oracle_params_for(agent, 11))
    el = batch_as_element(L.make_batch(1234 + B, B, ATARI["obs_dim"], 9, "cnn"))
    for _ in range(2):  # two steps: the second is the graph-captured launch on a non-zero state
        mu0 = agent.optimizer_state["mu"].flat.detach().clone().double()
        nu0 = agent.optimizer_state["nu"].flat.detach().clone().double()
        t = int(agent.optimizer_state["count"].item()) + 1
        p0 = agent.params.flat.detach().clone().double()
        grads, _ = agent.grad_on_batch(agent.params, el)
        g = grads.flat.detach().clone().double()
        agent.params, agent.optimizer_state, _ = agent.learn_on_batch(agent.params, agent.optimizer_state, el)
        torch.cuda.synchronize()
        b1, b2 = agent.adam_b1, agent.adam_b2
        mu = b1 * mu0 + (1 - b1) * g
        nu = b2 * nu0 + (1 - b2) * g * g
        step = agent.learning_rate * (mu / (1 - b1**t)) / (torch.sqrt(nu / (1 - b2**t)) + agent.adam_eps)
        got_mu = agent.optimizer_state["mu"].flat.double()
        got_nu = agent.optimizer_state["nu"].flat.double()
        got_step = p0 - agent.params.flat.double()
        # (the step is read back as a difference of fp32 parameters: their rounding is ~1e-4 of a 6e-5 step)
        for name, got, want, tol in (("mu", got_mu, mu, 2e-5), ("nu", got_nu, nu, 4e-5), ("step", got_step, step, 2e-3)):
            e = float((got - want).norm() / want.norm().clamp_min(1e-30))
            assert e <= tol, f"B={B} t={t}: {name} differs from Adam(grad_on_batch) by {e:.3e} (rel L2)"
        # the bf16 shadow the next forward reads is the rounding of the new parameters, everywhere
        assert torch.equal(agent.params.shadow.float(), agent.params.flat.to(torch.bfloat16).float())


@pytest.mark.parametrize("env", [{"ISDQN_MID": "1"}, {"ISDQN_PAIR": "0", "ISDQN_LNFUSE": "0", "ISDQN_TD_FOLD": "1"},
                                 {"ISDQN_PAIR_D1": "0", "ISDQN_TMA_STORE": "0", "ISDQN_ADAM_L2": "0"}],
                         ids=["head_mid", "unfused_backward_td_fold", "pair_proportional_staged_stores"])
def test_alternative_launch_plans_keep_parity(env):
    """The opt-in / fallback launch plans of the batch-32 chain (one-launch head step; separate weight-gradient,
    input-gradient and LayerNorm-backward launches) against the same oracle bars, in a fresh process (the switches are
    read once per process)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_learner_bf16_gpu.py"), "-q", "-m", "gpu",
                        "-p", "no:cacheprovider", "-k", "atari_k9_batch32 or update_is_adam or graph_replay"],
                       cwd=root, env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
