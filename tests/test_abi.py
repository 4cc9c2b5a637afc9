"""CPU-side checks of the C-ABI boundary: the library builds, loads and exports every symbol the
header declares; host-side helpers behave (no kernel is launched here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "mapf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mapf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mapf_gnn_b200 import _lib
    lib = _lib.load()
    declared = header_functions()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), name
    assert set(_lib.exported_symbols()) == set(declared)


def test_error_strings_and_version():
    from mapf_gnn_b200 import _lib
    lib = _lib.load()
    assert lib.mapf_version().startswith(b"mapf_b200")
    assert lib.mapf_last_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert len(lib.mapf_last_error_string(code)) > 3


def test_argument_validation_without_gpu():
    """Null / shape errors are reported by status code before any CUDA call."""
    from mapf_gnn_b200 import _lib
    lib = _lib.load()
    a = _lib.EnvStepArgs()
    assert lib.mapf_env_step(a, None) == -1            # null pointers
    assert lib.mapf_env_step(None, None) == -1
    p = _lib.NetParams()
    assert lib.mapf_packed_weights_size(p) == -1        # num_conv == 0 unsupported
    with pytest.raises(RuntimeError):
        _lib.check(-3, "x")


def test_packed_size_matches_layout():
    from mapf_gnn_b200 import _lib
    from mapf_gnn_b200.ops import net_params
    lib = _lib.load()
    dims = [3, 2, 16, 16, 16, 0, 400, 64, 3, 128, 5]
    n = lib.mapf_packed_weights_size(net_params(dims))
    expect = (288 + 16) + (2304 + 16) * 2 + 400 * 64 + 64 + 4 * 64 * 128 + 50 * 2 * 64 * 8 + 32 * 2 * 128 * 8 + 5 * 128 + 8
    expect += 4 + (1 + 3 + 3) * 2 * 48 * 8 + 4 + 25 * 64 * 16   # tensor-core conv B operands + pixel-major FC operand
    assert n == expect


def test_ops_are_cuda_only():
    import torch
    import mapf_gnn_b200  # noqa: F401
    with pytest.raises(NotImplementedError):
        torch.ops.mapf.head_fwd(torch.zeros(4, 64), torch.zeros(5, 64), torch.zeros(5))


def test_d2_limit():
    from mapf_gnn_b200 import d2_limit_from_range
    assert [d2_limit_from_range(r) for r in (6, 4, 2.5, 2, 1, 0, 1.5)] == [36, 16, 7, 4, 1, 0, 3]


def test_state_dict_layout_matches_shipped_weights():
    import numpy as np
    import torch
    import mapf_gnn_b200 as m
    from conftest import load_weights
    from helpers import model_config
    for name in ("gnn_k3", "gnn_k2", "gnn_msg_k3", "baseline", "gnn_k4_seed4"):
        net = m.build_network(model_config(name))
        sd = {k: torch.from_numpy(np.asarray(v)) for k, v in load_weights(name).items()}
        missing, unexpected = net.load_state_dict(sd, strict=True)
        assert not missing and not unexpected


def test_tensor_core_encoder_selection_is_host_logic():
    """Which encoder mapf_policy_fwd gets a workspace for is decided on the host from the channel widths (no GPU call):
    a first layer of 16 or 32 channels with 16-channel layers behind it -> tensor-core conv stack, else FFMA2 kernel."""
    from mapf_gnn_b200 import _lib, ops
    from mapf_gnn_b200.ops import net_params
    lib = _lib.load()
    gnn = net_params([3, 2, 16, 16, 16, 0, 400, 64, 3, 128, 5])
    base = net_params([3, 2, 32, 16, 16, 0, 400, 64, 0, 0, 5])
    other = net_params([3, 2, 16, 32, 16, 0, 400, 64, 0, 0, 5])
    assert lib.mapf_encoder_tc_supported(gnn) == 1 and lib.mapf_encoder_tc_supported(base) == 1
    assert lib.mapf_encoder_tc_supported(other) == 0
    assert ops.wants_encoder_workspace(gnn) == (ops.USE_TENSOR_CORES and (ops.USE_TENSOR_CORE_ENCODER or ops.USE_TENSOR_CORE_FC))
    assert ops.wants_encoder_workspace(other) == (ops.USE_TENSOR_CORES and ops.USE_TENSOR_CORE_FC)
    assert lib.mapf_tc_packed_b_f16_size(64, 400) == 4 + 25 * 64 * 16
    assert lib.mapf_tc_packed_b_f16_size(200, 400) == -1
