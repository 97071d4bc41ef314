"""On-disk formats.

The reference saves `agent.get_model()` = `{"params": {"params": {"Conv_0": {"kernel", "bias"}, "LayerNorm_0":
{"scale", "bias"}, ..., "Dense_1": {...}}}}` with `pickle.dump` (experiments/base/utils.py:123-135, isdqn.py:137-138).
Its leaves are `jax.Array`s, whose pickle stream calls `jax._src.array._reconstruct_array(fun, args, arr_state,
aval_state)` — i.e. NumPy's own reconstruction triple plus a device_put.  `load_model_pickle` reads such a file WITHOUT
jax by resolving exactly that global to a NumPy-only stand-in (every other global is refused except NumPy's own array /
dtype / scalar reconstruction and plain containers: a checkpoint is data, not code), and also reads pickles whose leaves
already are NumPy arrays (`jax.device_get`, or this package's `save_model`).

`agent_state` / `load_agent_state` are the resume format the reference lacks (it saves parameters only): parameters,
both Adam moments, the step counter and the per-head loss sums, as one `.npz`.
"""
from __future__ import annotations

import io
import pickle
from typing import Any, Dict

import numpy as np


def module_at(inner: dict, mod: str) -> dict:
    """The module dict at the path `mod` ("Dense_0", or "Stack_1/Conv_2" for impala's nested sub-modules)."""
    for part in mod.split("/"):
        inner = inner[part]
    return inner

# what a parameter checkpoint legitimately refers to: NumPy's array / dtype / scalar reconstruction and plain containers
_NUMPY_NAMES = {"_reconstruct", "ndarray", "dtype", "scalar", "_frombuffer"}
_BUILTIN_NAMES = {"dict", "list", "tuple", "set", "frozenset", "int", "float", "complex", "bool", "str", "bytes", "bytearray",
                  "slice", "range"}
_COLLECTIONS_NAMES = {"OrderedDict", "defaultdict"}


def _reconstruct_array(fun, args, arr_state, aval_state=None):
    """NumPy-only stand-in for jax._src.array._reconstruct_array (jax 0.4.30): `fun(*args)` is numpy's
    `_reconstruct`, `arr_state` its `__setstate__` tuple; the device_put and the weak-type flag are dropped."""
    value = fun(*args)
    value.__setstate__(arr_state)
    return value


class _FrozenDictStandIn(dict):
    """flax.core.frozen_dict.FrozenDict pickles as a class + state {'_dict': {...}}"""

    def __setstate__(self, state):
        self.update(state.get("_dict", state))


class _CheckpointUnpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):
        if module in ("jax._src.array", "jax.interpreters.xla", "jaxlib.xla_extension") and name == "_reconstruct_array":
            return _reconstruct_array
        if module == "flax.core.frozen_dict" and name == "FrozenDict":
            return _FrozenDictStandIn
        if (module == "numpy" or module.startswith("numpy.")) and name in _NUMPY_NAMES:
            return super().find_class(module, name)
        if module == "builtins" and name in _BUILTIN_NAMES:
            return super().find_class(module, name)
        if module == "collections" and name in _COLLECTIONS_NAMES:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"checkpoint refers to {module}.{name}: only numpy / jax-array leaves are accepted")


def load_model_pickle(path_or_bytes) -> Dict[str, Any]:
    """A reference (or isdqn_b200) model pickle -> nested dicts with NumPy leaves."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        f = io.BytesIO(path_or_bytes)
        return _CheckpointUnpickler(f).load()
    with open(path_or_bytes, "rb") as f:
        return _CheckpointUnpickler(f).load()


def flax_leaves(model: Dict[str, Any]) -> Dict[str, Dict[str, np.ndarray]]:
    """`{"params": {"params": {...}}}` (get_model), `{"params": {...}}` (agent.params) or the bare module dict ->
    {module: {leaf: ndarray}}.  A `batch_stats` collection next to "params" is rejected (batch_norm is out of scope)."""
    node = model
    for _ in range(2):
        if isinstance(node, dict) and "params" in node:
            if "batch_stats" in node and node["batch_stats"]:
                raise ValueError("checkpoint carries batch_stats: batch_norm networks are out of scope for isdqn_b200")
            node = node["params"]
    if not isinstance(node, dict) or not all(isinstance(v, dict) for v in node.values()):
        raise ValueError("not a flax-shaped parameter tree: expected {module: {leaf: array}}")
    out: Dict[str, Dict[str, np.ndarray]] = {}

    def walk(prefix: str, mod: dict) -> None:
        # a sub-module (impala's Stack_i) holds modules, a module holds array leaves; paths are joined with "/"
        for name, v in mod.items():
            if isinstance(v, dict):
                walk(f"{prefix}/{name}" if prefix else name, v)
            else:
                out.setdefault(prefix, {})[name] = np.asarray(v)

    walk("", node)
    return out


def check_against(specs, leaves: Dict[str, Dict[str, np.ndarray]]) -> None:
    """Names and shapes of a checkpoint against the network's leaf list [(module, leaf, shape)]; raises ValueError with
    every mismatch listed (a silently truncated or transposed kernel is worse than a refusal)."""
    want = {(m, l): tuple(s) for m, l, s in specs}
    got = {(m, l): tuple(v.shape) for m, lv in leaves.items() for l, v in lv.items()}
    problems = [f"missing {m}.{l} {s}" for (m, l), s in want.items() if (m, l) not in got]
    problems += [f"unexpected {m}.{l} {s}" for (m, l), s in got.items() if (m, l) not in want]
    problems += [f"{m}.{l}: checkpoint {got[(m, l)]} != network {s}" for (m, l), s in want.items()
                 if (m, l) in got and got[(m, l)] != s]
    if problems:
        raise ValueError("checkpoint does not fit this network: " + "; ".join(problems))


def save_model(agent, path) -> None:
    """What the reference's save_data does with the model (experiments/base/utils.py:134-135)."""
    with open(path, "wb") as f:
        pickle.dump(agent.get_model(), f)


def load_model(agent, model_or_path) -> None:
    """Parameters of a reference / isdqn_b200 model (dict or pickle path) into `agent.params` (fp32, in place)."""
    model = model_or_path if isinstance(model_or_path, dict) else load_model_pickle(model_or_path)
    leaves = flax_leaves(model)
    check_against(agent.params.specs, leaves)
    for mod, lv in leaves.items():
        for leaf, v in lv.items():
            module_at(agent.params["params"], mod)[leaf] = np.asarray(v, dtype=np.float32)


# ---------------------------------------------------------------------------------------------------- resume
def pack_state(params: Dict[str, Dict[str, np.ndarray]], mu, nu, count: int, cumulated: np.ndarray) -> Dict[str, np.ndarray]:
    out = {"count": np.asarray(count, dtype=np.int32), "cumulated_losses": np.asarray(cumulated, dtype=np.float64)}
    for prefix, tree in (("params", params), ("mu", mu), ("nu", nu)):
        for mod, lv in tree.items():
            for leaf, v in lv.items():
                out[f"{prefix}/{mod}/{leaf}"] = np.asarray(v, dtype=np.float32)
    return out


def unpack_state(npz) -> Dict[str, Any]:
    trees: Dict[str, Any] = {"params": {}, "mu": {}, "nu": {}}
    for key in npz.files if hasattr(npz, "files") else npz.keys():
        parts = key.split("/")
        if len(parts) >= 3 and parts[0] in trees:  # prefix / module path (may be nested) / leaf
            trees[parts[0]].setdefault("/".join(parts[1:-1]), {})[parts[-1]] = np.asarray(npz[key])
    trees["count"] = int(np.asarray(npz["count"]))
    trees["cumulated_losses"] = np.asarray(npz["cumulated_losses"], dtype=np.float64)
    return trees


def _host_tree(tree) -> Dict[str, Dict[str, np.ndarray]]:
    out: Dict[str, Dict[str, np.ndarray]] = {}
    for mod, leaf, v in tree.leaves():
        out.setdefault(mod, {})[leaf] = v.detach().cpu().numpy()
    return out


def save_agent_state(agent, path) -> None:
    """Everything `learn_on_batch` carries from one step to the next, as one .npz."""
    state = pack_state(_host_tree(agent.params), _host_tree(agent.optimizer_state["mu"]), _host_tree(agent.optimizer_state["nu"]),
                       int(agent.optimizer_state["count"].item()), agent.cumulated_losses)
    with open(path, "wb") as f:
        np.savez(f, **state)


def load_agent_state(agent, path) -> None:
    with np.load(path) as npz:
        st = unpack_state(npz)
    for name in ("params", "mu", "nu"):
        check_against(agent.params.specs, st[name])
    for name, tree in (("params", agent.params), ("mu", agent.optimizer_state["mu"]), ("nu", agent.optimizer_state["nu"])):
        for mod, lv in st[name].items():
            for leaf, v in lv.items():
                module_at(tree["params"], mod)[leaf] = v
    agent.optimizer_state["count"].fill_(st["count"])
    agent.cumulated_losses = st["cumulated_losses"]
