"""CPU: host logic of the product's n-step accumulator (frame references) against the oracle (materialised
stacks), on the same seeded transition streams as the golden scenarios."""
import numpy as np
import pytest

from isdqn_b200.sample_collection.accumulator import NStepAccumulator
from oracle.replay_oracle import ReplayOracle
from oracle.samplers_oracle import UniformSamplingOracle
from tests import scenarios as S


@pytest.mark.parametrize("sc", S.SCENARIOS, ids=lambda s: s.name)
def test_records_rebuild_the_reference_elements(sc):
    frames = []

    def commit(fr):
        fr.frame_id = len(frames)
        frames.append(np.array(fr.observation))
        return fr.frame_id

    acc = NStepAccumulator(sc.stack, sc.horizon, sc.gamma, commit)
    oracle = ReplayOracle(UniformSamplingOracle(0), sc.batch, 10**9, sc.stack, sc.horizon, sc.gamma)
    n_elems = 0
    for obs, action, reward, terminal, episode_end, _ in S.transition_stream(sc):
        records = list(acc.accumulate(obs, action, reward, terminal, episode_end))
        oracle.add(obs, action, reward, terminal, episode_end)
        assert oracle.add_count == n_elems + len(records)
        for refs, a, r, d in records:
            want = oracle.memory[n_elems]
            zero = np.zeros_like(obs)
            state = np.stack([frames[i] if i >= 0 else zero for i in refs[: sc.stack]], axis=-1)
            nxt = np.stack([frames[i] if i >= 0 else zero for i in refs[sc.stack :]], axis=-1)
            np.testing.assert_array_equal(state, want.state)
            np.testing.assert_array_equal(nxt, want.next_state)
            assert a == want.action and d == want.is_terminal
            assert np.float64(r).tobytes() == np.float64(want.reward).tobytes()
            n_elems += 1
    # lazily committed frames: never more than (1+n) per element plus the padding bound used for the ring size
    assert len(frames) <= (1 + sc.horizon) * max(n_elems, 1) + sc.stack + sc.horizon


def test_truncation_drops_the_last_transition():
    acc = NStepAccumulator(4, 1, 0.99, lambda fr: 0)
    out = []
    for t in range(1, 6):
        out += list(acc.accumulate(np.full((2, 2), t), t, float(t), False, t == 5))
    assert len(out) == 4 and len(acc.trajectory) == 0  # obs 5's own transition never appears
