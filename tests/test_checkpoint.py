"""CPU: the on-disk formats — reference model pickles (jax-array leaves read without jax), name/shape
validation, and the resume archive's key layout."""
import io
import pickle
import sys
import types

import numpy as np
import pytest

from isdqn_b200 import checkpoint as ck


def _fake_jax_pickle(tree):
    """Pickle `tree` the way jax 0.4.30 does: every leaf reduces to jax._src.array._reconstruct_array(fun, args,
    arr_state, aval_state) with numpy's own reconstruction triple (jax/_src/array.py, ArrayImpl.__reduce__)."""
    mods = {name: types.ModuleType(name) for name in ("jax", "jax._src", "jax._src.array")}

    def _reconstruct_array(fun, args, arr_state, aval_state):  # never called here: only its qualified name is pickled
        raise AssertionError

    _reconstruct_array.__module__ = "jax._src.array"
    _reconstruct_array.__qualname__ = "_reconstruct_array"
    mods["jax._src.array"]._reconstruct_array = _reconstruct_array

    class FakeJaxArray:
        def __init__(self, value):
            self._value = np.asarray(value)

        def __reduce__(self):
            fun, args, arr_state = self._value.__reduce__()
            return _reconstruct_array, (fun, args, arr_state, {"weak_type": False})

    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        wrapped = {m: {l: FakeJaxArray(v) for l, v in lv.items()} for m, lv in tree.items()}
        data = pickle.dumps({"params": {"params": wrapped}})
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert b"jax._src.array" in data and "jax" not in sys.modules
    return data


def _tree(seed=0):
    g = np.random.default_rng(seed)
    return {
        "Conv_0": {"kernel": g.standard_normal((8, 8, 4, 32)).astype(np.float32), "bias": g.standard_normal(32).astype(np.float32)},
        "LayerNorm_0": {"scale": g.standard_normal(32).astype(np.float32), "bias": g.standard_normal(32).astype(np.float32)},
        "Dense_0": {"kernel": g.standard_normal((16, 6)).astype(np.float32), "bias": np.zeros(6, np.float32)},
    }


def _specs(tree):
    return [(m, l, v.shape) for m, lv in tree.items() for l, v in lv.items()]


def test_reference_pickle_with_jax_leaves_loads_without_jax():
    tree = _tree()
    model = ck.load_model_pickle(_fake_jax_pickle(tree))
    leaves = ck.flax_leaves(model)
    ck.check_against(_specs(tree), leaves)
    for m, lv in tree.items():
        for l, v in lv.items():
            assert isinstance(leaves[m][l], np.ndarray) and leaves[m][l].dtype == v.dtype
            assert np.array_equal(leaves[m][l], v)


def test_numpy_scalars_and_dtypes_round_trip():
    model = {"params": {"params": {"Dense_0": {"kernel": np.arange(6, dtype=np.float16).reshape(2, 3), "bias": np.float32(1.5)}}},
             "step": np.int64(7)}
    got = ck.load_model_pickle(pickle.dumps(model))
    assert got["step"] == 7 and got["params"]["params"]["Dense_0"]["bias"] == np.float32(1.5)
    assert got["params"]["params"]["Dense_0"]["kernel"].dtype == np.float16


def test_numpy_pickle_and_file_path(tmp_path):
    tree = _tree(1)
    path = tmp_path / "model"
    with open(path, "wb") as f:
        pickle.dump({"params": {"params": tree}}, f)  # what save_data writes after jax.device_get
    leaves = ck.flax_leaves(ck.load_model_pickle(path))
    assert np.array_equal(leaves["Dense_0"]["kernel"], tree["Dense_0"]["kernel"])
    # the three nesting levels a caller may hold
    for model in ({"params": {"params": tree}}, {"params": tree}, tree):
        assert set(ck.flax_leaves(model)) == set(tree)


def test_code_in_a_checkpoint_is_refused():
    class Evil:
        def __reduce__(self):
            import os

            return os.system, ("true",)

    with pytest.raises(pickle.UnpicklingError):
        ck.load_model_pickle(pickle.dumps({"params": {"params": {"Dense_0": {"kernel": Evil()}}}}))

    class EvilBuiltin:  # builtins.eval is as much code as os.system
        def __reduce__(self):
            return eval, ("1 + 1",)

    with pytest.raises(pickle.UnpicklingError):
        ck.load_model_pickle(pickle.dumps({"params": EvilBuiltin()}))

    class EvilNumpy:  # so is anything in numpy that is not array reconstruction
        def __reduce__(self):
            return np.load, ("/etc/passwd",)

    with pytest.raises(pickle.UnpicklingError):
        ck.load_model_pickle(pickle.dumps({"params": EvilNumpy()}))


def test_mismatches_are_listed():
    tree = _tree(2)
    specs = _specs(tree)
    bad = {m: dict(lv) for m, lv in tree.items()}
    bad["Dense_0"]["kernel"] = bad["Dense_0"]["kernel"].T.copy()
    del bad["LayerNorm_0"]["scale"]
    bad["Dense_9"] = {"bias": np.zeros(3, np.float32)}
    with pytest.raises(ValueError) as e:
        ck.check_against(specs, bad)
    msg = str(e.value)
    assert "Dense_0.kernel" in msg and "missing LayerNorm_0.scale" in msg and "unexpected Dense_9.bias" in msg
    with pytest.raises(ValueError):
        ck.flax_leaves({"params": tree, "batch_stats": {"BatchNorm_0": {"mean": np.zeros(3)}}})


def test_resume_archive_round_trip():
    p, mu, nu = _tree(3), _tree(4), _tree(5)
    state = ck.pack_state(p, mu, nu, 1234, np.arange(9.0))
    buf = io.BytesIO()
    np.savez(buf, **state)
    buf.seek(0)
    with np.load(buf) as npz:
        st = ck.unpack_state(npz)
    assert st["count"] == 1234 and np.array_equal(st["cumulated_losses"], np.arange(9.0))
    for name, src in (("params", p), ("mu", mu), ("nu", nu)):
        for m, lv in src.items():
            for l, v in lv.items():
                assert np.array_equal(st[name][m][l], v)
