"""TEST INFRASTRUCTURE ONLY (oracle/): import shim for the *unmodified* reference modules.

The reference's replay side (`slimdqn/sample_collection/{sum_tree,samplers,replay_buffer}.py`) is
pure NumPy except for three import-time dependencies that are not installed in this image:
`jax` (only `jax.tree_util.tree_map`, replay_buffer.py:212), `flax.struct.PyTreeNode`
(replay_buffer.py:26,59,65) and `snappy` (replay_buffer.py:39,55).  This module registers ~20-line
stand-ins for those three in `sys.modules` so that the reference files import *as they lie* under
`/root/reference` (read-only).  snappy is lossless, so an identity codec is result-equivalent.

Used only by `oracle/make_golden.py` (fixture generation + pinning the restatements against the real
reference) inside the build container; `/root/reference` does not exist on the GPU box and nothing in
the product package imports this file.
"""
from __future__ import annotations

import dataclasses
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ISDQN_REFERENCE_ROOT", "/root/reference")


def _tree_map(f, *trees):
    first = trees[0]
    if dataclasses.is_dataclass(first):
        kw = {fld.name: _tree_map(f, *[getattr(t, fld.name) for t in trees]) for fld in dataclasses.fields(first)}
        return type(first)(**kw)
    return f(*trees)


class _PyTreeNode:
    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        dataclasses.dataclass(frozen=True)(cls)

    def replace(self, **updates):
        return dataclasses.replace(self, **updates)


def install() -> bool:
    """Register the stand-ins and put the reference on sys.path. Returns False if it is absent."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "slimdqn")):
        return False
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if "jax" not in sys.modules:
        jax = types.ModuleType("jax")
        tu = types.ModuleType("jax.tree_util")
        tu.tree_map = _tree_map
        jax.tree_util = tu
        sys.modules["jax"] = jax
        sys.modules["jax.tree_util"] = tu
    if "flax" not in sys.modules:
        flax = types.ModuleType("flax")
        struct = types.ModuleType("flax.struct")
        struct.PyTreeNode = _PyTreeNode
        flax.struct = struct
        sys.modules["flax"] = flax
        sys.modules["flax.struct"] = struct
    if "snappy" not in sys.modules:
        snappy = types.ModuleType("snappy")
        snappy.compress = lambda b: bytes(memoryview(b).cast("B")) if not isinstance(b, bytes) else b
        snappy.uncompress = lambda b: bytes(b)
        sys.modules["snappy"] = snappy
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return True


def pytest_configure(config):  # allows `pytest -p refshim` on the reference's own tests
    install()


if __name__ != "__main__":
    pass
