"""ORACLE (test infrastructure): golden-fixture generator.  Runs ONLY in the build container, where
`/root/reference` is mounted.

    python oracle/make_golden.py            # regenerate tests/golden/*.npz and pin the restatements
    python oracle/make_golden.py --big      # only the 1 M-capacity scenarios (minutes of reference time each)

For every scenario in `tests/scenarios.py` it
  1. imports the UNMODIFIED reference modules (`slimdqn.sample_collection.*`) under `oracle/refshim.py`,
  2. drives them through the seeded scenario and writes the outputs to `tests/golden/<name>.npz`,
  3. drives the oracle restatements through the same scenario and asserts bit-equality (the pin).
The fixtures travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from oracle.replay_oracle import ReplayOracle  # noqa: E402
from oracle.samplers_oracle import PrioritizedSamplingOracle, UniformSamplingOracle  # noqa: E402
from oracle.sum_tree_oracle import SumTreeOracle  # noqa: E402
from tests import scenarios as S  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


class ReferenceAdapter(S.Adapter):
    """The reference's own ReplayBuffer + samplers (compress=False: the identity-snappy shim makes
    compression a no-op on results anyway, replay_buffer.py:58-68)."""

    def __init__(self, sc: S.Scenario):
        from slimdqn.sample_collection import replay_buffer, samplers

        self._te = replay_buffer.TransitionElement
        if sc.sampler == "uniform":
            sampler = samplers.UniformSamplingDistribution(seed=sc.seed)
        else:
            sampler = samplers.PrioritizedSamplingDistribution(sc.seed, sc.capacity, sc.priority_exponent)
        self.rb = replay_buffer.ReplayBuffer(
            sampler, sc.batch, sc.capacity, stack_size=sc.stack, update_horizon=sc.horizon, gamma=sc.gamma, compress=False
        )
        self.sc = sc

    def add(self, obs, action, reward, terminal, episode_end, priority):
        t = self._te(obs, action, reward, terminal, episode_end)
        if priority is None:
            self.rb.add(t)
        else:
            self.rb.add(t, priority=priority)

    def add_count(self):
        return self.rb.add_count

    def sample(self):
        b = self.rb.sample()
        return (b.state, b.action, b.reward, b.next_state, b.is_terminal)

    def sample_keys(self, size):
        return self.rb._sampling_distribution.sample(size)

    def update(self, keys, priorities):
        self.rb.update(keys, priorities=priorities)

    def memory_keys(self):
        return list(self.rb._memory.keys())

    def index_to_key(self):
        return list(self.rb._sampling_distribution._index_to_key)

    def tree_nodes(self):
        sd = self.rb._sampling_distribution
        return sd._sum_tree._nodes.copy() if hasattr(sd, "_sum_tree") else None


class OracleAdapter(S.Adapter):
    def __init__(self, sc: S.Scenario):
        if sc.sampler == "uniform":
            sampler = UniformSamplingOracle(sc.seed)
        else:
            sampler = PrioritizedSamplingOracle(sc.seed, sc.capacity, sc.priority_exponent)
        self.rb = ReplayOracle(sampler, sc.batch, sc.capacity, sc.stack, sc.horizon, sc.gamma)

    def add(self, obs, action, reward, terminal, episode_end, priority):
        if priority is None:
            self.rb.add(obs, action, reward, terminal, episode_end)
        else:
            self.rb.add(obs, action, reward, terminal, episode_end, priority=priority)

    def add_count(self):
        return self.rb.add_count

    def sample(self):
        return tuple(self.rb.sample())

    def sample_keys(self, size):
        return self.rb.sampler.sample(size)

    def update(self, keys, priorities):
        self.rb.update(keys, priorities=priorities)

    def memory_keys(self):
        return list(self.rb.memory.keys())

    def index_to_key(self):
        return list(self.rb.sampler.index_to_key)

    def tree_nodes(self):
        return self.rb.sampler.tree._nodes.copy() if hasattr(self.rb.sampler, "tree") else None


def main() -> int:
    if not refshim.install():
        print("reference not mounted; nothing to do", file=sys.stderr)
        return 1
    from slimdqn.sample_collection import sum_tree as ref_sum_tree

    os.makedirs(GOLDEN, exist_ok=True)
    big = "--big" in sys.argv
    for sc in (S.BIG_SCENARIOS if big else S.SCENARIOS):
        want = S.run_scenario(sc, ReferenceAdapter(sc))
        if not big or "--with-oracle" in sys.argv:  # (the restatement is pinned at the small sizes; at 1 M it doubles the run)
            got = S.run_scenario(sc, OracleAdapter(sc))
            S.compare_results(got, want, where=f"oracle-vs-reference:{sc.name}")
        path = os.path.join(GOLDEN, f"replay_{sc.name}.npz")
        np.savez_compressed(path, **want)
        print(f"[golden] {sc.name:24s} adds={int(want['add_count']):5d} batches={len(want['digests']):3d} "
              f"-> {os.path.relpath(path, ROOT)} ({os.path.getsize(path) / 1024:.1f} KiB)  oracle==reference")
    for tt in ([] if big else S.TREE_TRACES):
        want = S.run_tree_trace(tt, ref_sum_tree.SumTree(tt.capacity))
        got = S.run_tree_trace(tt, SumTreeOracle(tt.capacity))
        S.compare_results(got, want, where=f"oracle-vs-reference:{tt.name}")
        path = os.path.join(GOLDEN, f"sumtree_{tt.name}.npz")
        np.savez_compressed(path, **want)
        print(f"[golden] sumtree {tt.name:16s} root={want['roots'][-1]:.17g} -> {os.path.relpath(path, ROOT)} "
              f"({os.path.getsize(path) / 1024:.1f} KiB)  oracle==reference")
    return 0


if __name__ == "__main__":
    sys.exit(main())
