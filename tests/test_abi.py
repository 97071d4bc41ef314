"""CPU: the C-ABI library loads without a GPU and exports every symbol include/isdqn_b200.h declares; host-only
entry points (layout, workspace sizes, error strings) answer; device objects refuse to exist without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest

from isdqn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "isdqn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(isdqn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/isdqn_b200.h but not exported"
        assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype in _lib.py"
    assert set(_lib.PROTOTYPES) == set(names)


def test_abi_version_and_error_strings():
    lib = _lib.load()
    assert lib.isdqn_abi_version() == _lib.ABI_VERSION
    assert lib.isdqn_strerror(0) == b"ok"
    assert b"invalid" in lib.isdqn_strerror(-1)


def test_layout_matches_flax_parameter_count():
    from isdqn_b200.networks.architectures.dqn import DQNNet

    net = DQNNet([32, 64, 64, 512], "cnn", 90, layer_norm=True)
    net.configure((84, 84, 4), 9, 9)
    assert net.n_params == 4_090_938
    names = [f"{m}.{l}" for m, l, _ in net._specs]
    assert names[:4] == ["Conv_0.kernel", "Conv_0.bias", "LayerNorm_0.scale", "LayerNorm_0.bias"]
    assert names[-2:] == ["Dense_1.kernel", "Dense_1.bias"]
    assert all(o % 4 == 0 for o in net._offsets)
    fc = DQNNet([100, 100], "fc", 16, layer_norm=False)
    fc.configure((8,), 3, 4)
    assert fc.n_params == 8 * 100 + 100 + 100 * 100 + 100 + 100 * 16 + 16
    assert _lib.load().isdqn_learn_workspace_bytes(net._net, 32) > 0


def test_out_of_scope_configurations_raise():
    from isdqn_b200.networks.architectures.dqn import DQNNet

    with pytest.raises(NotImplementedError):
        DQNNet([32, 64, 64, 512], "cnn", 90, layer_norm=True, batch_norm=True)
    with pytest.raises(NotImplementedError):
        DQNNet([32, 64, 64, 512], "resnet", 90)


@pytest.mark.parametrize("layer_norm", [True, False])
def test_impala_layout_matches_the_flax_leaf_list(layer_norm):
    """host-only: the native layout of `impala` (isdqn_net_layout) against the flax-named leaf list, Stack by Stack"""
    from isdqn_b200.networks.architectures.dqn import DQNNet, leaf_specs

    net = DQNNet([16, 32, 32, 256], "impala", 5 * 6, layer_norm=layer_norm)
    net.configure((84, 84, 4), 4, 6)
    specs = leaf_specs("impala", (84, 84, 4), [16, 32, 32, 256], 30, layer_norm)
    assert net._layout.n_leaves == len(specs) == (50 if layer_norm else 34)
    names = [m for m, _, _ in specs]
    assert names[0] == "Stack_0/Conv_0" and "Stack_2/Conv_4" in names and names[-1] == "Dense_1"
    assert ("LayerNorm_0" in names) == layer_norm and ("Stack_1/LayerNorm_1" in names) == layer_norm
    # 84 -> 42 -> 21 -> 11 through the three poolings: the Dense tail sees 11 * 11 * 32 features
    dense0 = [s for m, l, s in specs if m == "Dense_0" and l == "kernel"][0]
    assert dense0 == (11 * 11 * 32, 256)
    offs = [int(net._layout.offset[i]) for i in range(net._layout.n_leaves)]
    assert offs == sorted(offs) and all(o % 8 == 0 for o in offs)
    assert int(net._layout.total) >= offs[-1] + 30


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from isdqn_b200.sample_collection.sum_tree import SumTree

    with pytest.raises(_lib.IsdqnNativeError):
        SumTree(16)


def test_invalid_arguments_are_rejected_without_touching_the_gpu():
    lib = _lib.load()
    assert lib.isdqn_sumtree_query(None, 3, None, 4, None, None, None) == -1
    assert lib.isdqn_sumtree_set(None, 3, None, None, 4, None, None, None) == -1
    assert lib.isdqn_gather_stacks(None, 16, 4, 1, 4, None, None, None, None, None, 1, 0, None, None, None, None, None, None) == -1
    assert lib.isdqn_adam_step(None, None, None, None, None, 0.1, 0.9, 0.999, 1e-8, 4, None) == -1
