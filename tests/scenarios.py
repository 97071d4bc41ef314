"""Seeded replay scenarios shared by `oracle/make_golden.py` (runs them on the UNMODIFIED reference and
commits the results as fixtures), the oracle pin tests and the GPU parity tests (run them on the CUDA
path).  A scenario drives an implementation only through a tiny adapter (see `Adapter`), so the very
same call sequence hits reference, oracle and product.

Inputs are regenerated from the seed with `np.random.default_rng` (stable stream), outputs are compared
in full for the first `full_batches` sampled batches and by SHA-256 digest for the rest.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import Optional

import numpy as np


@dataclass(frozen=True)
class Scenario:
    name: str
    seed: int
    capacity: int
    batch: int
    stack: int
    horizon: int  # update_horizon n
    gamma: float
    steps: int
    obs_shape: tuple
    obs_dtype: str
    p_terminal: float
    p_truncate: float
    sampler: str  # "uniform" | "prioritized"
    sample_every: int
    sparse: float = 0.9  # fraction of zero pixels (Atari-like)
    priority_exponent: float = 1.0
    full_batches: int = 2
    int_rewards: bool = False  # the reference's tests feed Python ints as rewards/actions
    digest_only: bool = False  # big scenarios: maps / tree nodes are stored as digests, not in full


SCENARIOS = [
    # Atari-shaped uint8 frames, FIFO wrap-around, uniform sampler (BASELINE config 2 in miniature)
    Scenario("atari_u8_uniform", 0, 64, 8, 4, 1, 0.99, 400, (84, 84), "uint8", 0.02, 0.01, "uniform", 37, full_batches=1),
    # n-step returns, gamma != 1, terminal tails with right zero padding
    Scenario("nstep3_u8_uniform", 1, 50, 16, 4, 3, 0.5, 500, (12, 10), "uint8", 0.05, 0.03, "uniform", 23),
    # very short episodes: exercises the `trajectory_len < stack + n` branch and truncation drops
    Scenario("short_episodes", 2, 40, 16, 4, 2, 0.9, 400, (6, 5), "uint8", 0.35, 0.2, "uniform", 11),
    # non-uint8 observations (the reference's own tests use int64 frames; config 1 uses float32 vectors)
    Scenario("int64_frames", 3, 30, 8, 3, 1, 1.0, 200, (5, 4), "int64", 0.03, 0.02, "uniform", 17, int_rewards=True),
    Scenario("lunar_f32_vectors", 4, 100, 32, 1, 1, 0.99, 400, (8,), "float32", 0.02, 0.01, "uniform", 29),
    # stack 1 (reference testKeyMappingsForSampling shape) and capacity 10
    Scenario("stack1_cap10", 5, 10, 32, 1, 1, 0.99, 60, (4, 4), "uint8", 0.0, 0.0, "uniform", 7),
    # prioritized: non-power-of-two capacity, eviction wrap, priority updates with duplicate keys
    Scenario("prioritized_cap37", 6, 37, 16, 4, 1, 0.99, 300, (8, 8), "uint8", 0.04, 0.02, "prioritized", 13),
    Scenario("prioritized_alpha06", 7, 100, 32, 2, 2, 0.97, 500, (6, 6), "uint8", 0.03, 0.02, "prioritized", 19, priority_exponent=0.6),
]


# BASELINE.json configs[2] at its real size: a 1 M-transition prioritized buffer (sum tree of depth 21, leaf index =
# capacity touched by every add-then-evict), filled and then driven through 130 k evictions with priority updates in
# between.  Small frames keep the reference run that produced the fixture (oracle/make_golden.py --big) to minutes.
BIG_SCENARIOS = [
    Scenario("prioritized_cap1M", 8, 1_000_000, 32, 4, 1, 0.99, 1_131_000, (4, 4), "uint8", 0.001, 0.0005, "prioritized",
             50_000, sparse=0.5, digest_only=True),
]


def scenario_by_name(name: str) -> Scenario:
    for s in SCENARIOS + BIG_SCENARIOS:
        if s.name == name:
            return s
    raise KeyError(name)


def transition_stream(sc: Scenario):
    """Yields (obs, action, reward, is_terminal, episode_end, priority)."""
    rng = np.random.default_rng(sc.seed)
    dt = np.dtype(sc.obs_dtype)
    for _ in range(sc.steps):
        if dt.kind == "f":
            obs = rng.uniform(-1.0, 1.0, sc.obs_shape).astype(dt)
        else:
            obs = rng.integers(0, 256, sc.obs_shape).astype(dt)
            obs = obs * (rng.random(sc.obs_shape) >= sc.sparse).astype(dt)
        action = int(rng.integers(9))
        reward = int(rng.integers(-1, 2)) if sc.int_rewards else float(rng.integers(-1, 2))
        u = float(rng.random())
        terminal = u < sc.p_terminal
        episode_end = terminal or (u < sc.p_terminal + sc.p_truncate)
        prio = float(abs(rng.standard_normal()) + 1e-3)
        if rng.random() < 0.05:
            prio = 0.0  # zero priorities stay zero (samplers.py:73)
        yield obs, action, reward, terminal, episode_end, prio


class Adapter:
    """What a scenario needs from an implementation (subclassed per implementation)."""

    def add(self, obs, action, reward, terminal, episode_end, priority): ...
    def add_count(self) -> int: ...
    def sample(self): ...  # -> (state, action, reward, next_state, is_terminal) numpy arrays
    def sample_keys(self, size) -> np.ndarray: ...
    def update(self, keys, priorities): ...
    def memory_keys(self) -> list: ...
    def index_to_key(self) -> list: ...
    def tree_nodes(self) -> Optional[np.ndarray]: ...


def digest(a) -> str:
    a = np.ascontiguousarray(a)
    h = hashlib.sha256()
    h.update(str(a.dtype).encode())
    h.update(str(a.shape).encode())
    h.update(a.tobytes())
    return h.hexdigest()


FIELDS = ("state", "action", "reward", "next_state", "is_terminal")


def run_scenario(sc: Scenario, ad: Adapter) -> dict:
    """Drives `ad` through the scenario; returns a flat dict of numpy arrays (npz-friendly)."""
    out: dict = {}
    prio_rng = np.random.default_rng(sc.seed + 1000)
    n_batches = 0
    digests = []
    for t, (obs, action, reward, terminal, episode_end, prio) in enumerate(transition_stream(sc)):
        ad.add(obs, action, reward, terminal, episode_end, prio if sc.sampler == "prioritized" else None)
        if ad.add_count() > 0 and (t + 1) % sc.sample_every == 0:
            batch = ad.sample()
            if n_batches < sc.full_batches:
                for f, arr in zip(FIELDS, batch):
                    out[f"batch{n_batches}_{f}"] = np.asarray(arr)
            digests.append("|".join(digest(np.asarray(arr)) for arr in batch))
            n_batches += 1
            if sc.sampler == "prioritized":
                keys = ad.sample_keys(sc.batch)
                out[f"keys{n_batches}"] = np.asarray(keys)
                new_p = np.abs(prio_rng.standard_normal(sc.batch)) + 1e-3
                new_p[prio_rng.random(sc.batch) < 0.1] = 0.0
                ad.update(np.asarray(keys), new_p)
    out["digests"] = np.asarray(digests)
    out["add_count"] = np.asarray(ad.add_count())
    nodes = ad.tree_nodes()
    if sc.digest_only:
        mk = np.asarray(ad.memory_keys(), dtype=np.int64)
        out["memory_keys_span"] = np.asarray([mk[0], mk[-1], mk.size], dtype=np.int64)
        assert (np.diff(mk) == 1).all()
        out["index_to_key_digest"] = np.asarray(digest(np.asarray(ad.index_to_key(), dtype=np.int64)))
        if nodes is not None:
            nodes = np.asarray(nodes, dtype=np.float64)
            out["tree_nodes_digest"] = np.asarray(digest(nodes))
            out["tree_root"] = np.asarray(nodes[0])
        return out
    out["memory_keys"] = np.asarray(ad.memory_keys(), dtype=np.int64)
    out["index_to_key"] = np.asarray(ad.index_to_key(), dtype=np.int64)
    if nodes is not None:
        out["tree_nodes"] = np.asarray(nodes, dtype=np.float64)
    return out


def compare_results(got: dict, want: dict, where: str = "") -> None:
    """Bit-exact comparison of two `run_scenario` outputs (dtype, shape and bytes)."""
    assert set(got) == set(want), f"{where}: key sets differ: {sorted(set(got) ^ set(want))}"
    for k in sorted(want):
        g, w = np.asarray(got[k]), np.asarray(want[k])
        if w.dtype.kind in "US":
            assert g.tolist() == w.tolist(), f"{where}:{k} digests differ"
            continue
        assert g.shape == w.shape, f"{where}:{k} shape {g.shape} != {w.shape}"
        assert g.dtype == w.dtype, f"{where}:{k} dtype {g.dtype} != {w.dtype}"
        assert g.tobytes() == w.tobytes(), f"{where}:{k} bytes differ (max abs diff {np.max(np.abs(g.astype(np.float64) - w.astype(np.float64)))})"


# ----------------------------------------------------------------------------------------- sum tree traces
@dataclass(frozen=True)
class TreeTrace:
    name: str
    seed: int
    capacity: int
    n_ops: int
    max_m: int
    n_queries: int
    store_nodes: bool = True


TREE_TRACES = [
    TreeTrace("cap1", 10, 1, 10, 1, 8),
    TreeTrace("cap4", 11, 4, 40, 4, 64),
    TreeTrace("cap100", 12, 100, 200, 32, 256),
    TreeTrace("cap1000", 13, 1000, 300, 64, 512),
    TreeTrace("cap4101", 14, 4101, 200, 300, 512),
    # the Atari buffer: 1 M leaves (depth 21).  Nodes are compared by digest + root + query answers.
    TreeTrace("cap1M", 15, 1_000_000, 300, 32, 4096, store_nodes=False),
]


def tree_trace_ops(tt: TreeTrace):
    """Yields ("set", idx int32[m], val float64[m]) ops; values have duplicates and zeros."""
    rng = np.random.default_rng(tt.seed)
    for i in range(tt.n_ops):
        m = int(rng.integers(1, tt.max_m + 1))
        idx = rng.integers(0, tt.capacity, m).astype(np.int32)
        if m > 2 and rng.random() < 0.5:
            idx[rng.integers(0, m, m // 3)] = idx[0]  # duplicates: first one must win
        val = np.abs(rng.standard_normal(m)) * float(rng.choice([1e-3, 1.0, 1e3]))
        val[rng.random(m) < 0.1] = 0.0
        yield idx, val


def run_tree_trace(tt: TreeTrace, tree) -> dict:
    """`tree` offers set/get/root/query/_nodes/max_recorded_priority (reference API)."""
    rng = np.random.default_rng(tt.seed + 500)
    roots = []
    for idx, val in tree_trace_ops(tt):
        if idx.size == 1 and rng.random() < 0.5:
            tree.set(int(idx[0]), float(val[0]))  # scalar call form
        else:
            tree.set(idx, val)
        roots.append(float(tree.root))
    out = {"roots": np.asarray(roots, dtype=np.float64), "max_recorded_priority": np.asarray(float(tree.max_recorded_priority))}
    nodes = np.asarray(tree._nodes, dtype=np.float64)
    out["nodes_digest"] = np.asarray(digest(nodes))
    if tt.store_nodes:
        out["nodes"] = nodes.copy()
    root = float(tree.root)
    u = rng.random(tt.n_queries)
    targets = root * u
    # exact node boundaries are the interesting targets for the strict `<` rule
    if nodes.size >= 3:
        targets[0] = nodes[1] if nodes[1] < root else targets[0]
        targets[1] = 0.0
    out["query_targets"] = targets
    out["query_out"] = np.asarray(tree.query(targets))
    return out
