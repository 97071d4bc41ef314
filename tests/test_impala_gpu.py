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


def test_impala_bf16_tensor_core_convolutions():
    """compute_dtype="bfloat16": every convolution but Stack_0/Conv_0 on the tcgen05 tile engine (bf16 operands, fp32
    accumulation, fp32 residual stream / LayerNorm / Dense tail).  Bars of tests/test_learner_bf16_gpu.py: Q-values,
    targets, losses 2e-2 against the float64 oracle, gradients 15 % in L2; tight against the bf16-emulating oracle."""
    from tests.test_learner_bf16_gpu import check_bf16

    check_bf16(dict(IMPALA_84, features=[32, 64, 64, 512]), 32, seed=92, n_steps=2)
    check_bf16(dict(IMPALA_42, features=[32, 32, 64, 128], layer_norm=False), 5, seed=93, n_steps=1)


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
