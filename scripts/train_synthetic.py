"""The reference's training loop (`experiments/base/dqn.py:13-85` + `experiments/atari/isdqn.py:15-48`) on the drop-in
classes, with a synthetic Atari-shaped environment in place of ALE (not in this image): acting through
`collect_single_sample` / `select_action` (sample_collection/utils.py), `agent.update_online_params(step, rb)` and
`agent.update_target_params(step)` exactly where the reference calls them.  `--envs N` steps N environments per
iteration with one batched forward (`select_actions` + `rb.add_batch` per environment-major order is NOT what the
reference does; it is the throughput form) — the default 1 is the reference's loop.

    python scripts/train_synthetic.py --steps 4000 --prioritized
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


class SyntheticAtari:
    """The observation contract of `slimdqn/environments/atari.py`: `.observation` is the latest (84, 84) uint8 frame,
    `.state` the (84, 84, 4) stack, `.step(action) -> (reward, absorbing)`, `.reset()`, `.n_steps`, `.n_actions`."""

    def __init__(self, seed: int, n_actions: int = 9, p_terminal: float = 2e-3):
        self.rng = np.random.default_rng(seed)
        self.n_actions, self.p_terminal = n_actions, p_terminal
        self.state_ = np.zeros((84, 84, 4), dtype=np.uint8)
        self.reset()

    def _frame(self):
        f = self.rng.integers(0, 256, (84, 84), dtype=np.uint8)
        f *= self.rng.random((84, 84), dtype=np.float32) >= 0.9
        return f

    @property
    def observation(self):
        return np.copy(self.state_[:, :, -1])

    @property
    def state(self):
        return self.state_

    def reset(self):
        self.state_ = np.zeros((84, 84, 4), dtype=np.uint8)
        self.state_[:, :, -1] = self._frame()
        self.n_steps = 0

    def step(self, action):
        self.state_ = np.roll(self.state_, -1, axis=-1)
        self.state_[:, :, -1] = self._frame()
        self.n_steps += 1
        reward = float(self.rng.integers(-1, 2)) * (1.0 + 0.1 * (action % 3))
        return reward, bool(self.rng.random() < self.p_terminal)


def train(key, p, agent, env, rb, log=print):
    """experiments/base/dqn.py:13-85 without tqdm / wandb / the analysis branch; returns the list of log dicts."""
    from isdqn_b200.sample_collection.utils import collect_single_sample, linear_schedule, split

    epsilon_schedule = linear_schedule(1.0, p["epsilon_end"], p["epsilon_duration"])
    n_training_steps = 0
    env.reset()
    episode_returns_per_epoch = [[0]]
    episode_lengths_per_epoch = [[0]]
    logs_out = []
    for idx_epoch in range(p["n_epochs"]):
        n_training_steps_epoch = 0
        has_reset = False
        while n_training_steps_epoch < p["n_training_steps_per_epoch"] or not has_reset:
            key, exploration_key = split(key, 2)
            reward, has_reset = collect_single_sample(exploration_key, env, agent, rb, p, epsilon_schedule, n_training_steps)
            n_training_steps_epoch += 1
            n_training_steps += 1
            episode_returns_per_epoch[idx_epoch][-1] += reward
            episode_lengths_per_epoch[idx_epoch][-1] += 1
            if has_reset and n_training_steps_epoch < p["n_training_steps_per_epoch"]:
                episode_returns_per_epoch[idx_epoch].append(0)
                episode_lengths_per_epoch[idx_epoch].append(0)
            if n_training_steps > p["n_initial_samples"]:
                agent.update_online_params(n_training_steps, rb)
                target_updated, logs = agent.update_target_params(n_training_steps)
                if target_updated:
                    logs_out.append({"n_training_steps": n_training_steps, **logs})
                    log(logs_out[-1])
        log({"epoch": idx_epoch, "n_training_steps": n_training_steps,
             "avg_return": float(np.mean(episode_returns_per_epoch[idx_epoch])),
             "avg_length_episode": float(np.mean(episode_lengths_per_epoch[idx_epoch]))})
        if idx_epoch < p["n_epochs"] - 1:
            episode_returns_per_epoch.append([0])
            episode_lengths_per_epoch.append([0])
    return logs_out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--capacity", type=int, default=50_000)
    ap.add_argument("--prioritized", action="store_true")
    ap.add_argument("--dtype", default="bfloat16")
    ap.add_argument("--architecture", default="cnn", choices=["cnn", "impala"])
    args = ap.parse_args()
    if args.architecture == "impala":
        args.dtype = "float32"  # the tensor-core path covers cnn
    from isdqn_b200.networks.isdqn import iSDQN
    from isdqn_b200.sample_collection.replay_buffer import ReplayBuffer
    from isdqn_b200.sample_collection.samplers import PrioritizedSamplingDistribution, UniformSamplingDistribution

    env = SyntheticAtari(0)
    sampler = PrioritizedSamplingDistribution(0, args.capacity) if args.prioritized else UniformSamplingDistribution(0)
    rb = ReplayBuffer(sampler, 32, args.capacity, stack_size=4, update_horizon=1, gamma=0.99, clipping=lambda x: np.clip(x, -1, 1))
    if args.prioritized:  # new transitions enter at the maximum recorded priority
        add = rb.add
        rb.add = lambda t, **kw: add(t, priority="max")
    agent = iSDQN(0, (84, 84, 4), env.n_actions, 9, [32, 64, 64, 512], True, False, args.architecture, 6.25e-5, 0.99, 1, 4, 1000,
                  adam_eps=1.5e-4, compute_dtype=args.dtype)
    if args.prioritized:
        agent.prioritized_beta = 0.4
    p = {"epsilon_end": 0.01, "epsilon_duration": 2000, "n_epochs": 1, "n_training_steps_per_epoch": args.steps,
         "n_initial_samples": 500, "horizon": 27_000}
    t0 = time.perf_counter()
    train(1, p, agent, env, rb)
    dt = time.perf_counter() - t0
    print(f"{args.steps} environment steps, {max(0, (args.steps - 500) // 4)} updates in {dt:.1f} s "
          f"({args.steps / dt:.0f} env steps/s incl. the synthetic environment)")


if __name__ == "__main__":
    main()
