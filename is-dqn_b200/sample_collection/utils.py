"""Acting glue — drop-in for `slimdqn/sample_collection/utils.py` (`select_action`, `collect_single_sample`), plus the
batched form for N environments.

The reference jits `select_action` and draws with JAX's threefry PRNG: `uniform_key, action_key, kwargs_key = split(key, 3)`,
explore iff `uniform(uniform_key) <= epsilon_fn(n_training_steps)`, the random action is `randint(action_key, (), 0,
n_actions)`, the greedy one `best_action_fn(params, state, key=kwargs_key)` (utils.py:8-15).  Here the three draws are
the host restatements of those JAX functions (`isdqn_threefry_split / _uniform / _randint`, csrc/runtime.cu — the same key
gives the same decisions; block function pinned, composition unpinned, see oracle/threefry_oracle.py) and only the greedy
branch touches the GPU, when it is taken.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from ..networks.isdqn import _raw_key
from .replay_buffer import ReplayBuffer, TransitionElement


def split(key, num: int = 2) -> np.ndarray:
    """`jax.random.split(key, num)` on the host: uint32[num][2] raw keys."""
    k0, k1 = _raw_key(key)
    out = (C.c_uint32 * (2 * num))()
    _lib.load().isdqn_threefry_split(k0, k1, num, out)
    return np.ctypeslib.as_array(out).reshape(num, 2).copy()


def linear_schedule(init_value: float, end_value: float, transition_steps: int):
    """`optax.linear_schedule` (what experiments/base/dqn.py:20 builds for epsilon)."""

    def schedule(count):
        frac = min(max(count / transition_steps, 0.0), 1.0) if transition_steps > 0 else 1.0
        return init_value + (end_value - init_value) * frac

    return schedule


def select_action(best_action_fn, params, state, key, n_actions, epsilon_fn, n_training_steps):
    """utils.py:8-15.  Returns an int32 scalar (`.item()` works on it)."""
    lib = _lib.load()
    keys = split(key, 3)
    if lib.isdqn_threefry_uniform(int(keys[0][0]), int(keys[0][1])) <= epsilon_fn(n_training_steps):
        return np.int32(lib.isdqn_threefry_randint(int(keys[1][0]), int(keys[1][1]), 0, int(n_actions)))
    return best_action_fn(params, state, key=keys[2])


def collect_single_sample(key, env, agent, rb: ReplayBuffer, p, epsilon_schedule, n_training_steps: int):
    """utils.py:18-40, line for line: act, step the environment, add the transition (reward clipped by `rb._clipping`)."""
    action = select_action(
        agent.best_action, agent.params, env.state, key, env.n_actions, epsilon_schedule, n_training_steps
    ).item()

    obs = env.observation
    reward, absorbing = env.step(action)

    episode_end = absorbing or env.n_steps >= p["horizon"]
    rb.add(
        TransitionElement(
            observation=obs,
            action=action,
            reward=reward if rb._clipping is None else rb._clipping(reward),
            is_terminal=absorbing,
            episode_end=episode_end,
        )
    )

    if episode_end:
        env.reset()

    return reward, episode_end


def select_actions(agent, params, states, keys, n_actions, epsilon_fn, n_training_steps) -> np.ndarray:
    """`select_action` for N environments with ONE forward (new: the reference acts on one environment): row i makes the
    decisions `select_action(..., states[i], keys[i], ...)` makes; the greedy rows share a `best_actions` call."""
    lib = _lib.load()
    n = len(keys)
    eps = epsilon_fn(n_training_steps)
    actions = np.empty(n, dtype=np.int32)
    greedy, greedy_keys = [], []
    for i in range(n):
        k = split(keys[i], 3)
        if lib.isdqn_threefry_uniform(int(k[0][0]), int(k[0][1])) <= eps:
            actions[i] = lib.isdqn_threefry_randint(int(k[1][0]), int(k[1][1]), 0, int(n_actions))
        else:
            greedy.append(i)
            greedy_keys.append(k[2])
    if greedy:
        st = np.ascontiguousarray(np.asarray(states)[greedy])
        actions[greedy] = agent.best_actions(params, st, greedy_keys)
    return actions
