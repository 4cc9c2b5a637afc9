"""ctypes binding of ``include/mapf_b200.h`` -- the only way Python reaches the kernels.

The structures below mirror the C header field for field.  ``check(rc)`` turns a non-zero status
into ``RuntimeError`` with the library's own message.  There is deliberately no fallback: if the
shared library cannot be loaded the import fails.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

MAX_CONV = 4
# mapf_policy_fwd_args.flags / mapf_train_fwd_args.flags (include/mapf_b200.h)
POLICY_FFMA_CONV, POLICY_FC_UNSPLIT, POLICY_UNFUSED_GRAPH, POLICY_FUSED_ENV_ON, POLICY_FUSED_ENV_OFF, POLICY_NO_PDL, \
    POLICY_FFMA_HOPS = 1, 2, 4, 8, 16, 32, 64
TRAIN_FFMA_FC, TRAIN_FFMA_GRAPH, TRAIN_FFMA_BWD, TRAIN_TC_BWD_ALWAYS, TRAIN_FFMA_CONV, TRAIN_SINGLE_LANE, TRAIN_BN_PASS, \
    TRAIN_PREPARED = 1, 2, 4, 8, 16, 32, 64, 128
c_f32p = C.c_void_p      # device pointers travel as integers
c_i32p = C.c_void_p
c_u8p = C.c_void_p


class EnvResetArgs(C.Structure):
    _fields_ = [("episodes", C.c_int32), ("agents", C.c_int32), ("board", C.c_int32), ("num_obstacles", C.c_int32),
                ("cell_stride", C.c_int32),
                ("starts", c_i32p), ("obstacles", c_i32p), ("pos", c_i32p), ("cells", c_u8p),
                ("time", c_i32p), ("done", c_u8p)]


class EnvStepArgs(C.Structure):
    _fields_ = [("episodes", C.c_int32), ("agents", C.c_int32), ("board", C.c_int32), ("cell_stride", C.c_int32),
                ("d2_limit", C.c_int32), ("do_move", C.c_int32),
                ("pos", c_i32p), ("goal", c_i32p), ("actions", c_i32p), ("cells", c_u8p),
                ("time", c_i32p), ("done", c_u8p), ("fov", c_f32p), ("adj", c_f32p),
                ("episodes_per_cta", C.c_int32), ("pos_out", c_i32p), ("done_out", c_u8p)]


class EnvRulesArgs(C.Structure):
    _fields_ = [("episodes", C.c_int32), ("agents", C.c_int32), ("board", C.c_int32), ("cell_stride", C.c_int32),
                ("stages", C.c_int32), ("pos", c_i32p), ("pos_temp", c_i32p), ("cells", c_u8p)]


class EnvMetricsArgs(C.Structure):
    _fields_ = [("episodes", C.c_int32), ("agents", C.c_int32), ("max_time", C.c_int32),
                ("pos", c_i32p), ("goal", c_i32p), ("time", c_i32p), ("success", c_f32p), ("flow_time", c_i32p)]


class NetParams(C.Structure):
    _fields_ = [("num_conv", C.c_int32), ("channels", C.c_int32 * (MAX_CONV + 1)),
                ("fc_in", C.c_int32), ("fc_out", C.c_int32), ("taps", C.c_int32), ("node_dim", C.c_int32),
                ("num_actions", C.c_int32),
                ("conv_w", c_f32p * MAX_CONV), ("conv_b", c_f32p * MAX_CONV),
                ("bn_gamma", c_f32p * MAX_CONV), ("bn_beta", c_f32p * MAX_CONV),
                ("bn_mean", c_f32p * MAX_CONV), ("bn_var", c_f32p * MAX_CONV),
                ("fc_w", c_f32p), ("fc_b", c_f32p), ("gf_w", c_f32p), ("gf_w2", c_f32p),
                ("act_w", c_f32p), ("act_b", c_f32p)]


class EncoderFwdArgs(C.Structure):
    _fields_ = [("rows", C.c_int64), ("states", c_f32p), ("packed", c_f32p), ("x", c_f32p)]


class GraphFilterFwdArgs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("agents", C.c_int32), ("in_dim", C.c_int32), ("out_dim", C.c_int32),
                ("taps", C.c_int32), ("add_self_loop", C.c_int32), ("has_skip", C.c_int32), ("relu", C.c_int32),
                ("num_actions", C.c_int32),
                ("x", c_f32p), ("x_sb", C.c_int64), ("x_sn", C.c_int64), ("x_sf", C.c_int64),
                ("gso", c_f32p), ("wt", c_f32p),
                ("y", c_f32p), ("y_sb", C.c_int64), ("y_sn", C.c_int64), ("y_sg", C.c_int64),
                ("z", c_f32p), ("act_w", c_f32p), ("act_b", c_f32p), ("logits", c_f32p), ("actions", c_i32p)]


class HeadFwdArgs(C.Structure):
    _fields_ = [("rows", C.c_int64), ("in_dim", C.c_int32), ("num_actions", C.c_int32),
                ("x", c_f32p), ("act_w", c_f32p), ("act_b", c_f32p), ("logits", c_f32p), ("actions", c_i32p)]


class PolicyFwdArgs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("agents", C.c_int32), ("message", C.c_int32),
                ("states", c_f32p), ("gso", c_f32p), ("packed", c_f32p), ("workspace_x", c_f32p),
                ("workspace_z", c_f32p), ("workspace_act", c_f32p), ("logits", c_f32p), ("actions", c_i32p),
                ("flags", C.c_int32)]


class NetGrads(C.Structure):
    _fields_ = [("conv_w", c_f32p * MAX_CONV), ("conv_b", c_f32p * MAX_CONV),
                ("bn_gamma", c_f32p * MAX_CONV), ("bn_beta", c_f32p * MAX_CONV),
                ("fc_w", c_f32p), ("fc_b", c_f32p), ("gf_w", c_f32p), ("gf_w2", c_f32p),
                ("act_w", c_f32p), ("act_b", c_f32p)]


class TrainFwdArgs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("agents", C.c_int32), ("message", C.c_int32), ("momentum", C.c_float),
                ("states", c_f32p), ("gso", c_f32p), ("workspace", c_f32p), ("logits", c_f32p),
                ("bn_batches", C.c_void_p * MAX_CONV), ("flags", C.c_int32)]


class TrainBwdArgs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("agents", C.c_int32), ("message", C.c_int32),
                ("states", c_f32p), ("gso", c_f32p), ("workspace", c_f32p), ("dlogits", c_f32p),
                ("grads", NetGrads), ("flags", C.c_int32)]


class CeLossArgs(C.Structure):
    _fields_ = [("rows", C.c_int64), ("num_actions", C.c_int32), ("logits", c_f32p), ("targets", c_i32p),
                ("loss", c_f32p), ("dlogits", c_f32p)]


class ClipAdamArgs(C.Structure):
    _fields_ = [("n", C.c_int64), ("params", c_f32p), ("grads", c_f32p), ("exp_avg", c_f32p),
                ("exp_avg_sq", c_f32p), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("weight_decay", C.c_float), ("max_norm", C.c_float), ("step", C.c_int32),
                ("grad_norm", c_f32p), ("scratch", C.c_void_p), ("step_dev", C.c_void_p), ("lr_dev", C.c_void_p)]


class PeerReduceArgs(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("peer_grads", C.c_void_p * 8), ("peer_flags", C.c_void_p * 8)]


class CbsArgs(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("num_agents", C.c_int32),
                ("starts", C.POINTER(C.c_int32)), ("goals", C.POINTER(C.c_int32)),
                ("num_obstacles", C.c_int32), ("obstacles", C.POINTER(C.c_int32)),
                ("max_high_level_iterations", C.c_int32), ("max_low_level_iterations", C.c_int32),
                ("max_path", C.c_int32), ("path_len", C.POINTER(C.c_int32)), ("paths", C.POINTER(C.c_int32)),
                ("solved", C.c_int32), ("truncated", C.c_int32), ("expanded", C.c_int32), ("generated", C.c_int32),
                ("cost", C.c_int64)]


class GraphFilterBwdArgs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("agents", C.c_int32), ("in_dim", C.c_int32), ("out_dim", C.c_int32),
                ("taps", C.c_int32), ("add_self_loop", C.c_int32),
                ("x", c_f32p), ("gso", c_f32p), ("w", c_f32p), ("w2", c_f32p), ("dy", c_f32p),
                ("workspace", c_f32p), ("dx", c_f32p), ("dw", c_f32p), ("dw2", c_f32p)]


class RolloutArgs(C.Structure):
    _fields_ = [("env", EnvStepArgs), ("policy", PolicyFwdArgs), ("steps", C.c_int32)]


_SIGNATURES = {
    "mapf_last_error_string": (C.c_char_p, [C.c_int]),
    "mapf_version": (C.c_char_p, []),
    "mapf_graph_launch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mapf_event_wait": (C.c_int, [C.c_void_p]),
    "mapf_make_scenarios": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "mapf_env_reset": (C.c_int, [C.POINTER(EnvResetArgs), C.c_void_p]),
    "mapf_env_step": (C.c_int, [C.POINTER(EnvStepArgs), C.c_void_p]),
    "mapf_env_metrics": (C.c_int, [C.POINTER(EnvMetricsArgs), C.c_void_p]),
    "mapf_env_apply_rules": (C.c_int, [C.POINTER(EnvRulesArgs), C.c_void_p]),
    "mapf_packed_weights_size": (C.c_int64, [C.POINTER(NetParams)]),
    "mapf_pack_weights": (C.c_int, [C.POINTER(NetParams), C.c_void_p, C.c_void_p]),
    "mapf_pack_graph_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                          C.c_void_p]),
    "mapf_encoder_fwd": (C.c_int, [C.POINTER(NetParams), C.POINTER(EncoderFwdArgs), C.c_void_p]),
    "mapf_encoder_conv_fwd": (C.c_int, [C.POINTER(NetParams), C.POINTER(EncoderFwdArgs), C.c_void_p]),
    "mapf_graph_filter_fwd": (C.c_int, [C.POINTER(GraphFilterFwdArgs), C.c_void_p]),
    "mapf_head_fwd": (C.c_int, [C.POINTER(HeadFwdArgs), C.c_void_p]),
    "mapf_policy_fwd": (C.c_int, [C.POINTER(NetParams), C.POINTER(PolicyFwdArgs), C.c_void_p]),
    "mapf_rollout": (C.c_int, [C.POINTER(NetParams), C.POINTER(RolloutArgs), C.c_void_p]),
    "mapf_train_workspace_size": (C.c_int64, [C.POINTER(NetParams), C.c_int32, C.c_int32]),
    "mapf_train_forward": (C.c_int, [C.POINTER(NetParams), C.POINTER(TrainFwdArgs), C.c_void_p]),
    "mapf_train_backward": (C.c_int, [C.POINTER(NetParams), C.POINTER(TrainBwdArgs), C.c_void_p]),
    "mapf_ce_loss": (C.c_int, [C.POINTER(CeLossArgs), C.c_void_p]),
    "mapf_clip_adam": (C.c_int, [C.POINTER(ClipAdamArgs), C.c_void_p]),
    "mapf_peer_reduce_clip_adam": (C.c_int, [C.POINTER(PeerReduceArgs), C.POINTER(ClipAdamArgs), C.c_void_p]),
    "mapf_peer_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]),
    "mapf_peer_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "mapf_peer_close": (C.c_int, [C.c_void_p]),
    "mapf_peer_free": (C.c_int, [C.c_void_p]),
    "mapf_graph_filter_bwd_workspace_size": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                          C.c_int32]),
    "mapf_graph_filter_bwd": (C.c_int, [C.POINTER(GraphFilterBwdArgs), C.c_void_p]),
    "mapf_tc_packed_b_size": (C.c_int64, [C.c_int32, C.c_int32]),
    "mapf_tc_pack_b": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mapf_tc_gemm": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                               C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "mapf_tc_gemm_w3": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "mapf_cbs_solve": (C.c_int, [C.POINTER(CbsArgs)]),
    "mapf_cbs_solve_batch": (C.c_int, [C.POINTER(CbsArgs), C.c_int32, C.c_int32]),
    "mapf_cbs_a_star": (C.c_int, [C.POINTER(CbsArgs), C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32),
                                  C.c_int32, C.POINTER(C.c_int32)]),
    "mapf_tc_pack_b_tiles": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mapf_tc_gemm_tiles": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "mapf_tc_gemm_tn": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "mapf_tc_presplit_a_size": (C.c_int64, [C.c_int64, C.c_int32]),
    "mapf_encoder_conv_split_fwd": (C.c_int, [C.POINTER(NetParams), C.POINTER(EncoderFwdArgs), C.c_void_p]),
    "mapf_tc_fc_presplit": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_int32, C.c_void_p]),
    "mapf_encoder_tc_supported": (C.c_int, [C.POINTER(NetParams)]),
    "mapf_encoder_tc_conv_fwd": (C.c_int, [C.POINTER(NetParams), C.POINTER(EncoderFwdArgs), C.c_void_p]),
    "mapf_tc_packed_b_f16_size": (C.c_int64, [C.c_int32, C.c_int32]),
    "mapf_tc_pack_b_f16_pixel_major": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mapf_tc_fc_presplit_f16": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_int32, C.c_void_p]),
    "mapf_tc_mix_head": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mapf_tc_graph_supported": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "mapf_tc_graph_head": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "mapf_graph_taps": (C.c_int, [C.POINTER(GraphFilterFwdArgs), C.c_void_p]),
    "mapf_graph_taps_ffma": (C.c_int, [C.POINTER(GraphFilterFwdArgs), C.c_void_p]),
}

_lib = None


def exported_symbols():
    """Every entry point ``include/mapf_b200.h`` declares."""
    return sorted(_SIGNATURES)


def library_path() -> str:
    return _build.LIB


def load():
    """Load (building first if the sources changed) libmapf_b200.so.  Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build_library()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -m mapf_gnn_b200._build`")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().mapf_last_error_string(rc).decode()
        hint = ""
        if rc == -5:      # MAPF_ERR_CUDA: the most common cause outside a B200 box is the wrong architecture
            try:
                import torch
                if torch.cuda.is_available():
                    cap = torch.cuda.get_device_capability()
                    if cap != (10, 0):
                        hint = (f"; the library is built for sm_100a only (tcgen05 / TMEM kernels) and this device is "
                                f"sm_{cap[0]}{cap[1]}: there is no other code path")
            except Exception:      # noqa: BLE001
                pass
        raise RuntimeError(f"mapf_b200 {what}: {msg} (status {rc}){hint}")


def ptr(t):
    """Device pointer of a tensor (or None) as an integer for a ctypes c_void_p field."""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
