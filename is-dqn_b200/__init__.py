"""isdqn_b200 — B200-native (sm_100a) learner hot path of iS-DQN behind the reference's `slimdqn` API.

    from isdqn_b200.sample_collection.replay_buffer import ReplayBuffer, TransitionElement, ReplayElement
    from isdqn_b200.sample_collection.samplers import UniformSamplingDistribution, PrioritizedSamplingDistribution
    from isdqn_b200.sample_collection.sum_tree import SumTree
    from isdqn_b200.networks.isdqn import iSDQN

The directory is named `is-dqn_b200/` (not importable as such); `isdqn_b200/__init__.py` at the repo root is the
import alias.  Everything computes in `lib/libisdqn_b200.so` (C-ABI: include/isdqn_b200.h); there is no CPU path.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
