"""GPU: the device replay buffer through the reference API.  Golden scenarios (fixtures produced by the
unmodified reference) must match bit for bit; the reference's own replay tests are restated; the fused
normalising gathers and the device-resident sample path are checked against the raw path."""
import os

import numpy as np
import pytest

from tests import scenarios as S

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

OBSERVATION_SHAPE = (84, 84)
STACK_SIZE = 4
BATCH_SIZE = 32


@pytest.fixture(scope="module")
def mods():
    from isdqn_b200.sample_collection import replay_buffer, samplers

    return replay_buffer, samplers


class ProductAdapter(S.Adapter):
    def __init__(self, sc, mods, **kw):
        replay_buffer, samplers = mods
        self.TE = replay_buffer.TransitionElement
        if sc.sampler == "uniform":
            sampler = samplers.UniformSamplingDistribution(seed=sc.seed)
        else:
            sampler = samplers.PrioritizedSamplingDistribution(sc.seed, sc.capacity, sc.priority_exponent)
        self.rb = replay_buffer.ReplayBuffer(sampler, sc.batch, sc.capacity, stack_size=sc.stack,
                                             update_horizon=sc.horizon, gamma=sc.gamma, compress=False, **kw)

    def add(self, obs, action, reward, terminal, episode_end, priority):
        t = self.TE(obs, action, reward, terminal, episode_end)
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


@pytest.mark.parametrize("sc", S.SCENARIOS, ids=lambda s: s.name)
def test_golden_scenarios_bit_exact(mods, sc):
    want = dict(np.load(os.path.join(GOLDEN, f"replay_{sc.name}.npz")))
    got = S.run_scenario(sc, ProductAdapter(sc, mods))
    S.compare_results(got, want, where=sc.name)


def test_small_staging_and_tight_ring(mods):
    """Same scenario with a 3-frame pinned staging buffer (forces mid-stream flushes and ring wrap-around)."""
    sc = S.scenario_by_name("nstep3_u8_uniform")
    want = dict(np.load(os.path.join(GOLDEN, f"replay_{sc.name}.npz")))
    got = S.run_scenario(sc, ProductAdapter(sc, mods, staging_frames=3))
    S.compare_results(got, want, where=sc.name + "/staging3")


def test_frame_ring_overflow_is_loud(mods):
    replay_buffer, samplers = mods
    from isdqn_b200 import _lib

    rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(0), 4, 50, stack_size=4, frame_capacity=8)
    with pytest.raises(_lib.IsdqnNativeError):
        for i in range(40):
            rb.add(replay_buffer.TransitionElement(np.full((4, 4), i, np.uint8), 0, 0.0, False))


# ---- reference tests/test_replay_buffer.py restated on the product -------------------------------------------------
def test_element_pack_unpack(mods):
    replay_buffer, _ = mods
    state = np.zeros(OBSERVATION_SHAPE + (STACK_SIZE,), dtype=np.uint8)
    next_state = np.ones(OBSERVATION_SHAPE + (STACK_SIZE,), dtype=np.uint8)
    el = replay_buffer.ReplayElement(state=state, action=1, reward=1.0, next_state=next_state, is_terminal=False)
    un = el.pack().unpack()
    assert un.action == 1 and un.reward == 1.0 and un.is_terminal is False
    np.testing.assert_array_equal(un.state, state)
    np.testing.assert_array_equal(un.next_state, next_state)


def test_add_up_to_capacity(mods):
    replay_buffer, samplers = mods
    capacity = 10
    rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(seed=0), BATCH_SIZE, capacity,
                                    stack_size=STACK_SIZE, update_horizon=1, gamma=1.0, compress=False)
    transitions = []
    for i in range(16):
        transitions.append(replay_buffer.TransitionElement(np.full(OBSERVATION_SHAPE, i), i, i, False, False))
        rb.add(transitions[-1])
    assert len(rb._memory) == capacity
    expected_keys = list(range(5, 5 + capacity))
    assert list(rb._memory.keys()) == expected_keys
    for i in expected_keys:
        np.testing.assert_array_equal(
            rb._memory[i].state,
            np.array([t.observation for t in transitions[i - STACK_SIZE + 1 : i + 1]]).transpose(1, 2, 0))
        np.testing.assert_array_equal(
            rb._memory[i].next_state,
            np.array([t.observation for t in transitions[i - STACK_SIZE + 2 : i + 2]]).transpose(1, 2, 0))
        assert rb._memory[i].action == transitions[i].action
        assert rb._memory[i].reward == transitions[i].reward
        assert rb._memory[i].is_terminal == int(transitions[i].is_terminal)


def test_n_step_rewards(mods):
    replay_buffer, samplers = mods
    rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(seed=0), BATCH_SIZE, 10,
                                    stack_size=STACK_SIZE, update_horizon=5, gamma=1.0, compress=False)
    for i in range(50):
        rb.add(replay_buffer.TransitionElement(np.full(OBSERVATION_SHAPE, i), 0, 2.0, False))
    for _ in range(20):
        np.testing.assert_array_equal(rb.sample().reward, np.ones(BATCH_SIZE) * 10.0)


def test_get_stack(mods):
    replay_buffer, samplers = mods
    rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(seed=0), BATCH_SIZE, 50,
                                    stack_size=STACK_SIZE, update_horizon=1, gamma=1.0, compress=False)
    for i in range(11):
        rb.add(replay_buffer.TransitionElement(np.full(OBSERVATION_SHAPE, i), 0, 0, False))
    for i in rb._memory:
        np.testing.assert_array_equal(rb._memory[i].state.shape, OBSERVATION_SHAPE + (4,))
    np.testing.assert_array_equal(np.zeros(OBSERVATION_SHAPE + (3,)), rb._memory[0].state[:, :, :3])
    state = rb._memory[STACK_SIZE - 1].state
    for i in range(STACK_SIZE):
        np.testing.assert_array_equal(np.full(OBSERVATION_SHAPE, i), state[:, :, i])


def test_key_mappings_for_sampling(mods):
    replay_buffer, samplers = mods
    capacity = 10
    rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(seed=0), BATCH_SIZE, capacity,
                                    stack_size=1, update_horizon=1, gamma=0.99, compress=False)
    sampler = rb._sampling_distribution
    for i in range(capacity + 1):
        rb.add(replay_buffer.TransitionElement(np.full(OBSERVATION_SHAPE, i), i, i, False, False))
    for i in range(capacity):
        assert i in sampler._key_to_index
        index = sampler._key_to_index[i]
        assert i == index and i == sampler._index_to_key[index]
    next_key = capacity
    rb.add(replay_buffer.TransitionElement(np.full(OBSERVATION_SHAPE, next_key + 1), next_key + 1, next_key + 1, False, False))
    assert 0 not in sampler._key_to_index
    assert sampler._index_to_key[0] != 0
    assert next_key in sampler._key_to_index
    assert next_key == sampler._index_to_key[sampler._key_to_index[next_key]]
    indices = np.random.default_rng(seed=0).integers(len(sampler._index_to_key), size=BATCH_SIZE)
    keys = [sampler._index_to_key[index] for index in indices]
    samples = rb.sample()
    for i, key in enumerate(keys):
        np.testing.assert_array_equal(samples.state[i, ...], np.full(OBSERVATION_SHAPE, key)[..., None])
        np.testing.assert_array_equal(samples.next_state[i, ...], np.full(OBSERVATION_SHAPE, key + 1)[..., None])
        assert samples.action[i] == key and samples.reward[i] == key and samples.is_terminal[i] == 0


# ---- device-resident path and fused normalisation ---------------------------------------------------------------------
def test_device_sample_equals_host_sample_and_fused_normalise(mods):
    import torch

    from isdqn_b200 import _lib

    replay_buffer, samplers = mods
    rbs = []
    for _ in range(4):
        rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(seed=11), 64, 300, stack_size=4,
                                        update_horizon=1, gamma=0.99)
        sc = S.Scenario("x", 21, 300, 64, 4, 1, 0.99, 700, (84, 84), "uint8", 0.02, 0.01, "uniform", 10**9)
        for obs, a, r, d, e, _ in S.transition_stream(sc):
            rb.add(replay_buffer.TransitionElement(obs, a, r, d, e))
        rbs.append(rb)
    host = rbs[0].sample()
    dev = rbs[1].sample_device()
    f32 = rbs[2].sample_device(out_dtype=_lib.OUT_F32)
    bf16 = rbs[3].sample_device(out_dtype=_lib.OUT_BF16)
    assert dev.state.shape == (64, 84, 84, 4) and dev.state.dtype == torch.uint8
    np.testing.assert_array_equal(dev.state.cpu().numpy(), host.state)
    np.testing.assert_array_equal(dev.next_state.cpu().numpy(), host.next_state)
    np.testing.assert_array_equal(dev.action.cpu().numpy(), host.action)
    np.testing.assert_array_equal(dev.reward.cpu().numpy(), host.reward)
    np.testing.assert_array_equal(dev.is_terminal.cpu().numpy().astype(bool), host.is_terminal)
    want = host.state.astype(np.float32) / np.float32(255.0)  # architectures/dqn.py:51
    np.testing.assert_array_equal(f32.state.cpu().numpy(), want)
    np.testing.assert_array_equal(f32.next_state.cpu().numpy(), host.next_state.astype(np.float32) / np.float32(255.0))
    np.testing.assert_array_equal(bf16.state.float().cpu().numpy(), torch.from_numpy(want).to(torch.bfloat16).float().numpy())


def test_large_launch_roundtrip_property(mods):
    """Full-size property check (no oracle needed): every gathered stack must equal the frames it references —
    state[..., j] of element k is the frame stored for observation k - 3 + j of a long single episode."""
    replay_buffer, samplers = mods
    cap = 5000
    rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(seed=2), 32, cap, stack_size=4)
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, (cap + 500, 84, 84), dtype=np.uint8)
    for i in range(cap + 500):
        rb.add(replay_buffer.TransitionElement(frames[i], i % 9, float(i % 3 - 1), False))
    keys = rb._sampling_distribution.sample(4096)
    batch = rb._gather_keys(keys)
    for j in range(4):
        idx = keys.astype(np.int64) - 3 + j  # element k: state ends at observation k
        np.testing.assert_array_equal(batch.state[..., j], np.where((idx >= 0)[:, None, None], frames[np.maximum(idx, 0)], 0))
        np.testing.assert_array_equal(batch.next_state[..., j], frames[idx + 1])
    np.testing.assert_array_equal(batch.action, keys % 9)
