"""GPU: the batched / device-resident add path and the prioritized training driver, and
BASELINE.json configs[2] at its real size (1 M-transition prioritized buffer).

  * every golden scenario (fixtures produced by the UNMODIFIED reference) replayed through `ReplayBuffer.add_batch`,
    from host arrays and from CUDA tensors: bit-exact;
  * the 1 M-capacity prioritized scenario (fixture: the reference itself at capacity 1 M, 131 k evictions): maps,
    sum-tree digest, sampled batches, keys;
  * a prioritized training loop (add at max priority -> sample with keys + importance weights -> update) against the
    oracle driver: keys, batches, tree nodes and maps bit-exact, weights to 1e-6; host and device update paths.
"""
import os

import numpy as np
import pytest

from tests import scenarios as S

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def mods():
    from isdqn_b200.sample_collection import replay_buffer, samplers

    return replay_buffer, samplers


class BatchedAdapter(S.Adapter):
    """Buffers the scenario's transitions and hands them to `add_batch` whenever the scenario is about to look."""

    def __init__(self, sc, mods, device_obs=False, chunk=None, **kw):
        import torch

        self.torch = torch
        replay_buffer, samplers = mods
        if sc.sampler == "uniform":
            sampler = samplers.UniformSamplingDistribution(seed=sc.seed)
        else:
            sampler = samplers.PrioritizedSamplingDistribution(sc.seed, sc.capacity, sc.priority_exponent)
        self.rb = replay_buffer.ReplayBuffer(sampler, sc.batch, sc.capacity, stack_size=sc.stack,
                                             update_horizon=sc.horizon, gamma=sc.gamma, compress=False, **kw)
        self.sc, self.device_obs, self.chunk = sc, device_obs, chunk
        self.pending = []
        self.t = 0

    def add(self, obs, action, reward, terminal, episode_end, priority):
        self.pending.append((obs, action, reward, terminal, episode_end, priority))
        self.t += 1

    def flush(self):
        if not self.pending:
            return
        obs, a, r, term, end, prio = zip(*self.pending)
        self.pending = []
        obs = np.stack(obs)
        if self.device_obs:
            obs = self.torch.from_numpy(obs).cuda()
        prios = None if prio[0] is None else list(prio)
        n = len(a)
        step = self.chunk or n
        for i in range(0, n, step):
            self.rb.add_batch(obs[i : i + step], list(a[i : i + step]), list(r[i : i + step]), list(term[i : i + step]),
                              list(end[i : i + step]), priorities=None if prios is None else prios[i : i + step])

    def add_count(self):
        if self.t % self.sc.sample_every == 0 or self.t == self.sc.steps:
            self.flush()
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


@pytest.mark.parametrize("sc", S.SCENARIOS, ids=lambda s: s.name)
def test_add_batch_golden_scenarios_bit_exact(mods, sc):
    want = dict(np.load(os.path.join(GOLDEN, f"replay_{sc.name}.npz")))
    got = S.run_scenario(sc, BatchedAdapter(sc, mods))
    S.compare_results(got, want, where=sc.name + "/add_batch")


@pytest.mark.parametrize("name", ["atari_u8_uniform", "nstep3_u8_uniform", "prioritized_cap37", "lunar_f32_vectors"])
def test_add_batch_device_observations_bit_exact(mods, name):
    """Observations handed over as CUDA tensors (a device-resident environment): frames go ring-to-ring on the device."""
    sc = S.scenario_by_name(name)
    want = dict(np.load(os.path.join(GOLDEN, f"replay_{sc.name}.npz")))
    got = S.run_scenario(sc, BatchedAdapter(sc, mods, device_obs=True, chunk=7))
    S.compare_results(got, want, where=sc.name + "/add_batch(cuda)")


def test_add_batch_small_staging(mods):
    sc = S.scenario_by_name("nstep3_u8_uniform")
    want = dict(np.load(os.path.join(GOLDEN, f"replay_{sc.name}.npz")))
    got = S.run_scenario(sc, BatchedAdapter(sc, mods, staging_frames=3))
    S.compare_results(got, want, where=sc.name + "/add_batch/staging3")


def test_config2_prioritized_capacity_1M(mods):
    """BASELINE.json configs[2]: sum-tree prioritized replay with a 1 M-transition buffer (depth-21 tree), 131 k evictions
    through leaf index = capacity, priority updates in between.  The fixture is the UNMODIFIED reference run at this size."""
    sc = S.scenario_by_name("prioritized_cap1M")
    path = os.path.join(GOLDEN, f"replay_{sc.name}.npz")
    want = dict(np.load(path))
    got = S.run_scenario(sc, BatchedAdapter(sc, mods, chunk=200_000))
    S.compare_results(got, want, where=sc.name)


def _driver_stream(seed, steps, obs_shape):
    rng = np.random.default_rng(seed)
    for _ in range(steps):
        obs = rng.integers(0, 256, obs_shape).astype(np.uint8)
        u = float(rng.random())
        yield obs, int(rng.integers(9)), float(rng.integers(-1, 2)), u < 0.01, u < 0.015


def _synthetic_td(keys, step):
    """Deterministic stand-in for |TD| (a function of what both sides agree on bit for bit)."""
    return np.abs(np.sin(np.asarray(keys, dtype=np.float64) * 0.37 + step)) + 1e-3


@pytest.mark.parametrize("device_update", [False, True])
def test_prioritized_training_driver_matches_oracle(mods, device_update):
    """10 k environment steps, one prioritized update every 4 (2.5 k updates): new transitions enter at
    max_recorded_priority (resolved on the device), `sample(beta=)` returns keys and importance weights, `update` /
    `update_device` write the new priorities back.  Keys, batches, maps and every sum-tree node equal the oracle's."""
    import torch

    from oracle.prioritized_driver_oracle import PrioritizedDriverOracle

    replay_buffer, samplers = mods
    cap, B, steps, beta = 3001, 32, 10_000, 0.5
    ref = PrioritizedDriverOracle(3, cap, B, 4, 1, 0.99)
    rb = replay_buffer.ReplayBuffer(samplers.PrioritizedSamplingDistribution(3, cap), B, cap, stack_size=4, update_horizon=1,
                                    gamma=0.99, compress=False)
    pend = []
    n_updates = 0
    for t, (obs, a, r, term, end) in enumerate(_driver_stream(11, steps, (8, 8))):
        ref.add(obs, a, r, term, end)
        pend.append((obs, a, r, term, end))
        if (t + 1) % 4:
            continue
        o, aa, rr, tt, ee = zip(*pend)
        pend = []
        rb.add_batch(np.stack(o), list(aa), list(rr), list(tt), list(ee), priorities="max")
        if rb.add_count < 200:
            continue
        want_b, want_k, want_w = ref.sample(beta)
        if device_update:
            got_b, d_keys, d_w = rb.sample_device(beta=beta)
            got_k, got_w = d_keys.cpu().numpy(), d_w.cpu().numpy()
            got = tuple(np.asarray(x.cpu().numpy()) for x in got_b)
        else:
            got_b, got_k, got_w = rb.sample(beta=beta)
            got = tuple(np.asarray(x) for x in got_b)
        np.testing.assert_array_equal(got_k, want_k)
        np.testing.assert_allclose(got_w, want_w, rtol=2e-6)
        for g, w in zip(got, want_b):
            assert np.asarray(g).reshape(-1).astype(np.float64).tolist() == np.asarray(w).reshape(-1).astype(np.float64).tolist()
        p = _synthetic_td(want_k, n_updates)
        ref.update(want_k, p)
        if device_update:
            rb.update_device(d_keys, torch.from_numpy(p).cuda())
        else:
            rb.update(got_k, priorities=p)
        n_updates += 1
    assert n_updates >= 2400
    sd = rb._sampling_distribution
    sd.check_status()
    assert list(sd._index_to_key) == list(ref.sampler.index_to_key)
    np.testing.assert_array_equal(sd._sum_tree._nodes, ref.sampler.tree._nodes)
    assert sd._sum_tree.max_recorded_priority == ref.sampler.tree.max_recorded_priority


def test_update_device_missing_key_raises_and_leaves_tree(mods):
    import torch

    _, samplers = mods
    sd = samplers.PrioritizedSamplingDistribution(0, 100)
    for k in range(50):
        sd.add(k, 1.0 + k)
    sd.remove(10)
    before = sd._sum_tree._nodes.copy()
    sd.update_device(torch.tensor([3, 10, 4], dtype=torch.int32, device="cuda"), torch.tensor([9.0, 9.0, 9.0], dtype=torch.float64, device="cuda"))
    with pytest.raises(KeyError):
        sd.check_status()
    np.testing.assert_array_equal(sd._sum_tree._nodes, before)
    # per-head |TD| matrix form: mean over the rows, in float32
    td = torch.tensor([[1.0, 2.0], [3.0, 5.0]], dtype=torch.float32, device="cuda")
    sd.update_device(torch.tensor([3, 4], dtype=torch.int32, device="cuda"), td, prio_rows=2, offset=0.5)
    sd.check_status()
    assert sd._sum_tree.get(sd._key_to_index[3]) == 2.5 and sd._sum_tree.get(sd._key_to_index[4]) == 4.0


def test_prioritized_sampler_clamps_past_the_live_keys(mods):
    """a descent that ends past the live keys (empty tree) must not index the tables out of range on the
    device path; the status is raised lazily as the reference's IndexError / AttributeError."""
    _, samplers = mods
    sd = samplers.PrioritizedSamplingDistribution(0, 1000)
    for k in range(5):
        sd.add(k, 0.0)  # all-zero tree: every descent goes right, to leaf 1023
    idx, key, slot = sd.sample_device(64, 1001)
    assert int(key.max()) <= 4 and int(key.min()) >= 0 and int(slot.max()) <= 4
    assert int(idx.min()) >= 5
    with pytest.raises((IndexError, AttributeError)):
        sd.check_status()


def test_agent_prioritized_update_runs_on_device(mods):
    """`update_online_params` with `agent.prioritized_beta` set: importance-weighted loss, |TD| written back as priorities."""
    import torch

    from isdqn_b200.networks.isdqn import iSDQN

    replay_buffer, samplers = mods
    cap, B, K = 500, 32, 3
    rb = replay_buffer.ReplayBuffer(samplers.PrioritizedSamplingDistribution(1, cap), B, cap, stack_size=4, update_horizon=1,
                                    gamma=0.99, compress=False)
    rng = np.random.default_rng(0)
    N = 700
    rb.add_batch(rng.integers(0, 256, (N, 84, 84), dtype=np.uint8), rng.integers(0, 4, N), rng.integers(-1, 2, N).astype(np.float64),
                 rng.random(N) < 0.01, priorities="max")
    agent = iSDQN(0, (84, 84, 4), 4, K, [32, 64, 64, 512], True, False, "cnn", 1e-4, 0.99, 1, 1, 100, compute_dtype="bfloat16")
    agent.prioritized_beta = 0.4
    sd = rb._sampling_distribution
    root0 = sd._sum_tree.root
    for step in range(1, 6):
        agent.update_online_params(step, rb)
    torch.cuda.synchronize()
    sd.check_status()
    td = agent.td_abs(B).cpu().numpy()
    assert td.shape == (K, B) and np.isfinite(td).all() and (td >= 0).all()
    assert sd._sum_tree.root != root0  # priorities moved from 1.0 to |TD|
    # the weighted loss: with weights of one it equals the plain loss, with other weights it is the weighted mean
    batch = rb.sample()
    _, (l_plain, _) = agent.loss_on_batch(agent.params, batch)
    w = rng.random(B).astype(np.float32) + 0.5
    ctx = agent._context(B)
    q = agent.last_all_q_values.reshape(2 * B, -1).contiguous()
    from isdqn_b200 import _lib

    losses = torch.empty(K, dtype=torch.float32, device="cuda")
    tdm = torch.empty((K, B), dtype=torch.float32, device="cuda")
    d_w = torch.from_numpy(w).cuda()
    _lib.check(_lib.load().isdqn_heads_td_loss_weighted(
        q.data_ptr(), ctx["action"].data_ptr(), ctx["reward"].data_ptr(), ctx["terminal"].data_ptr(), float(0.99), B, B, K, 4,
        d_w.data_ptr(), losses.data_ptr(), None, tdm.data_ptr(), _lib.stream_ptr()), "weighted loss")
    want = (torch.from_numpy(w).cuda()[None, :] * tdm * tdm).mean(dim=1)
    np.testing.assert_allclose(losses.cpu().numpy(), want.cpu().numpy(), rtol=1e-5)
    np.testing.assert_allclose((tdm * tdm).mean(dim=1).cpu().numpy(), l_plain.cpu().numpy(), rtol=1e-5)


def test_captured_prioritized_step_equals_the_launch_by_launch_driver(mods):
    """The prioritized training step as ONE graph replay (draw with importance weights -> gather -> weighted step -> priority
    write-back, `isdqn_sample_prioritized_train`) against the same loop run launch by launch (`ISDQN_GRAPH_SAMPLE=0` path:
    `sample_device(beta=)` + `learn_on_batch` + `update_device`): same draws, same weights, same trees, same parameters —
    with adds and evictions between the updates and an annealed beta."""
    import torch

    from isdqn_b200.networks.isdqn import iSDQN

    replay_buffer, samplers = mods
    cap, B, K = 400, 32, 3
    rng = np.random.default_rng(5)
    N0, N1 = 380, 8
    frames = rng.integers(0, 256, (N0 + 30 * N1, 84, 84), dtype=np.uint8)
    acts, rews = rng.integers(0, 4, len(frames)), rng.integers(-1, 2, len(frames)).astype(np.float64)
    terms = rng.random(len(frames)) < 0.01

    def run(captured: bool):
        rb = replay_buffer.ReplayBuffer(samplers.PrioritizedSamplingDistribution(3, cap), B, cap, stack_size=4, update_horizon=1,
                                        gamma=0.99, compress=False)
        rb.add_batch(frames[:N0], acts[:N0], rews[:N0], terms[:N0], priorities="max")
        agent = iSDQN(0, (84, 84, 4), 4, K, [32, 64, 64, 512], True, False, "cnn", 1e-4, 0.99, 1, 1, 10**9, compute_dtype="bfloat16")
        weights = []
        for step in range(1, 31):
            agent.prioritized_beta = 0.4 + 0.02 * step  # annealed: the captured step reads beta from device memory
            if captured:
                assert agent._learn_from_replay(rb, prioritized=True)
            else:
                batch, d_keys, d_w = rb.sample_device(out=agent.batch_buffers(B), beta=agent.prioritized_beta)
                agent.params, agent.optimizer_state, _ = agent.learn_on_batch(agent.params, agent.optimizer_state, batch,
                                                                              _accumulate=True, is_weights=d_w)
                rb.update_device(d_keys, agent.td_abs(B), prio_rows=K, offset=agent.prioritized_eps)
            weights.append(agent._context(B)["is_weights"].cpu().numpy().copy())
            lo = N0 + (step - 1) * N1  # the buffer keeps moving: 8 adds (evictions from step 3 on) per update
            rb.add_batch(frames[lo:lo + N1], acts[lo:lo + N1], rews[lo:lo + N1], terms[lo:lo + N1], priorities="max")
        torch.cuda.synchronize()
        sd = rb._sampling_distribution
        sd.check_status()
        return np.stack(weights), sd._sum_tree._nodes.copy(), agent.params.flat.cpu().numpy(), list(sd._index_to_key)

    w_a, tree_a, p_a, keys_a = run(True)
    w_b, tree_b, p_b, keys_b = run(False)
    assert keys_a == keys_b
    np.testing.assert_allclose(w_a, w_b, rtol=2e-6)
    assert (w_a.max(axis=1) == 1.0).all() and (w_a > 0).all()
    np.testing.assert_allclose(tree_a, tree_b, rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(p_a, p_b, rtol=0, atol=1e-5 * np.abs(p_b).max())
