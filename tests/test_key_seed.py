"""CPU: the key -> NumPy seed mapping of the agent (head draw of best_action, parameter initialisation).  The fast form
must give the entropy pool of the original `SeedSequence(list of the key's bytes)` for every kind of key a caller passes."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_key_to_seed():
    # the function lives in the agent module, which needs CUDA to import fully: read it out of the source instead
    src = open(os.path.join(ROOT, "is-dqn_b200", "networks", "isdqn.py")).read()
    start = src.index("def _key_to_seed")
    end = src.index("\ndef ", start + 1)
    ns = {"np": np}
    exec(src[start:end], ns)
    return ns["_key_to_seed"]


def _original(key):
    return np.random.SeedSequence(np.frombuffer(np.asarray(key).tobytes(), dtype=np.uint8).tolist() or [0])


@pytest.mark.parametrize("key", [0, 1, 7, 2**31 - 1, 2**40 + 5, np.uint32(9), np.int64(-3), np.array([1, 2], dtype=np.uint32),
                                 np.array([0, 0], dtype=np.uint32), np.array([], dtype=np.uint32), np.arange(4, dtype=np.uint8)])
def test_same_entropy_pool_and_draws(key):
    fast = _load_key_to_seed()
    a, b = fast(key), _original(key)
    assert np.array_equal(a.generate_state(8), b.generate_state(8))
    for K in (1, 3, 9, 50):
        assert int(np.random.default_rng(fast(key)).integers(K)) == int(np.random.default_rng(_original(key)).integers(K))
