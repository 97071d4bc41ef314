"""GPU: model export/import and resume through the agent."""
import pickle

import numpy as np
import pytest
import torch

from oracle import learner_oracle as L
from tests.learner_utils import batch_as_element, make_agent, oracle_params_for, push_params

pytestmark = pytest.mark.gpu

CFG = dict(obs_dim=(84, 84, 4), A=6, K=3, features=[32, 64, 64, 512], layer_norm=True, arch="cnn")


def test_model_pickle_round_trip(tmp_path):
    from isdqn_b200 import checkpoint as ck

    a = make_agent(51, **CFG)
    push_params(a, oracle_params_for(a, 51))
    path = tmp_path / "51"
    ck.save_model(a, path)
    with open(path, "rb") as f:  # the reference's layout: {"params": {"params": {module: {leaf: ndarray}}}}
        raw = pickle.load(f)
    assert set(raw) == {"params"} and set(raw["params"]) == {"params"} and "Dense_1" in raw["params"]["params"]
    b = make_agent(52, **CFG, compute_dtype="bfloat16")
    b.load_model(path)
    assert torch.equal(a.params.flat, b.params.flat)
    assert b.params.shadow_dirty  # the bf16 shadow is rebuilt before the next tensor-core step
    wrong = make_agent(53, **dict(CFG, A=5))
    with pytest.raises(ValueError):
        wrong.load_model(path)


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_resume_continues_bit_identically(tmp_path, dtype):
    from isdqn_b200 import checkpoint as ck

    batches = [batch_as_element(L.make_batch(600 + i, 32, CFG["obs_dim"], CFG["A"], "cnn")) for i in range(5)]
    a = make_agent(54, **CFG, compute_dtype=dtype)
    push_params(a, oracle_params_for(a, 54))
    for el in batches[:3]:
        a.params, a.optimizer_state, _ = a.learn_on_batch(a.params, a.optimizer_state, el, _accumulate=True)
    path = tmp_path / "state.npz"
    ck.save_agent_state(a, path)
    b = make_agent(99, **CFG, compute_dtype=dtype)
    ck.load_agent_state(b, path)
    assert int(b.optimizer_state["count"].item()) == 3
    assert np.array_equal(a.cumulated_losses, b.cumulated_losses)
    for el in batches[3:]:
        a.params, a.optimizer_state, la = a.learn_on_batch(a.params, a.optimizer_state, el, _accumulate=True)
        b.params, b.optimizer_state, lb = b.learn_on_batch(b.params, b.optimizer_state, el, _accumulate=True)
        assert torch.equal(la, lb)
    torch.cuda.synchronize()
    assert torch.equal(a.params.flat, b.params.flat)
    assert torch.equal(a.optimizer_state["nu"].flat, b.optimizer_state["nu"].flat)
    assert np.array_equal(a.cumulated_losses, b.cumulated_losses)
