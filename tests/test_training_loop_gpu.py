"""GPU: the reference's training loop (experiments/base/dqn.py:13-85) on the drop-in classes — `collect_single_sample` /
`select_action` (sample_collection/utils.py), `update_online_params`, `update_target_params` — with a synthetic
environment, uniform and prioritized replay; the replay contents are checked against the oracle fed the same transitions."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


@pytest.mark.parametrize("prioritized,arch", [(False, "cnn"), (True, "cnn"), (False, "impala"), (True, "impala")])
def test_reference_training_loop_runs_on_the_drop_in_classes(prioritized, arch):
    import torch
    from train_synthetic import SyntheticAtari, train

    from isdqn_b200.networks.isdqn import iSDQN
    from isdqn_b200.sample_collection.replay_buffer import ReplayBuffer
    from isdqn_b200.sample_collection.samplers import PrioritizedSamplingDistribution, UniformSamplingDistribution
    from oracle.replay_oracle import ReplayOracle
    from oracle.samplers_oracle import UniformSamplingOracle

    cap, steps = 300, (700 if arch == "cnn" else 520)
    env = SyntheticAtari(3, p_terminal=0.02)
    sampler = PrioritizedSamplingDistribution(0, cap) if prioritized else UniformSamplingDistribution(0)
    rb = ReplayBuffer(sampler, 32, cap, stack_size=4, update_horizon=1, gamma=0.99, clipping=lambda x: np.clip(x, -1, 1))
    mirror = ReplayOracle(UniformSamplingOracle(0), 32, cap, 4, 1, 0.99)
    add = rb.add

    def add_both(t, **kw):
        mirror.add(t.observation, t.action, t.reward, t.is_terminal, t.episode_end)
        add(t, **({"priority": "max"} if prioritized else {}))

    rb.add = add_both
    feats = [32, 64, 64, 512] if arch == "cnn" else [16, 32, 32, 256]
    agent = iSDQN(0, (84, 84, 4), env.n_actions, 3, feats, True, False, arch, 1e-4, 0.99, 1, 4, 200,
                  adam_eps=1.5e-4, compute_dtype="bfloat16" if arch == "cnn" else "float32")
    if prioritized:
        agent.prioritized_beta = 0.4
    p = {"epsilon_end": 0.05, "epsilon_duration": 300, "n_epochs": 1, "n_training_steps_per_epoch": steps,
         "n_initial_samples": 100, "horizon": 150}
    before = agent.params.flat.clone()
    logs = train(5, p, agent, env, rb, log=lambda x: None)
    torch.cuda.synchronize()
    rb._sampling_distribution.check_status()
    # target updates every 200 steps after the first 100: at least 3 log records with the reference's keys, finite losses
    assert len(logs) >= (3 if arch == "cnn" else 2)
    for rec in logs:
        assert {"loss", "networks/0_loss", "networks/2_loss", "n_training_steps"} <= set(rec)
        assert np.isfinite(rec["loss"]) and rec["loss"] >= 0
    assert not torch.equal(before, agent.params.flat)
    # what the buffer holds is what the reference would hold for these transitions
    assert rb.add_count == mirror.add_count and list(rb._memory.keys()) == list(mirror.memory.keys())
    keys = np.asarray(list(mirror.memory.keys())[-40:])
    got, want = rb._gather_keys(keys), mirror.gather(keys)
    for g, w in zip(got, want):
        assert np.asarray(g).astype(np.float64).tolist() == np.asarray(w).astype(np.float64).tolist()
