"""``mapf::`` torch custom ops: typed schemas, CUDA-only dispatch, bodies = one C-ABI call each.

The ops exist so that the facade modules (``Network``, ``GCNLayer`` ...) stay ordinary PyTorch
modules (device placement, streams, autograd) while every number is produced by libmapf_b200.so.
They are registered for the CUDA dispatch key only: calling one with CPU tensors raises
``NotImplementedError`` from the dispatcher -- there is no CPU path.

``dims`` is the flat description of a network that every model op takes:
``[num_conv, c0, c1, c2, c3, c4, fc_in, fc_out, taps, node_dim, num_actions]``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib

_LIB = torch.library.Library("mapf", "DEF")

# graph-filter mix on tcgen05 (3xTF32) when True; the fused fp32-FFMA kernel when False (development switch)
USE_TENSOR_CORES = os.environ.get("MAPF_B200_TENSOR_CORES", "1") != "0"
# encoder on tcgen05: the conv stack (first layer 16 or 32 channels, 16-channel layers behind it) runs on the tensor cores
# (enc_tc.cu) with the FC behind it as an all-TMA tcgen05 GEMM; other channel widths keep the fused FFMA2 encoder kernel.  MAPF_B200_TENSOR_CORE_ENCODER=0
# forces the FFMA2 kernel; MAPF_B200_TENSOR_CORE_FC=1 hands the workspace to every network (development switch: FFMA2
# conv kernel emitting the pre-split operand + tensor-core FC where the tensor-core conv stack does not apply).
USE_TENSOR_CORE_ENCODER = os.environ.get("MAPF_B200_TENSOR_CORE_ENCODER", "1") != "0"
USE_TENSOR_CORE_FC = os.environ.get("MAPF_B200_TENSOR_CORE_FC", "0") == "1"


def policy_flags() -> int:
    """Development switches of the policy launch sequence.  They live HERE (the library never reads the environment)
    and are read per call: MAPF_B200_ENC_FFMA_CONV / FC_UNSPLIT / UNFUSED_GRAPH / NO_PDL = 1, MAPF_B200_FUSED_ENV = 0|1."""
    env = os.environ
    f = 0
    if env.get("MAPF_B200_ENC_FFMA_CONV"):
        f |= _lib.POLICY_FFMA_CONV
    if env.get("MAPF_B200_FC_UNSPLIT"):
        f |= _lib.POLICY_FC_UNSPLIT
    if env.get("MAPF_B200_UNFUSED_GRAPH"):
        f |= _lib.POLICY_UNFUSED_GRAPH
    if env.get("MAPF_B200_NO_PDL"):
        f |= _lib.POLICY_NO_PDL
    if env.get("MAPF_B200_FFMA_HOPS"):
        f |= _lib.POLICY_FFMA_HOPS
    fe = env.get("MAPF_B200_FUSED_ENV")
    if fe:
        f |= _lib.POLICY_FUSED_ENV_ON if fe[0] == "1" else _lib.POLICY_FUSED_ENV_OFF
    return f


def train_flags() -> int:
    env = os.environ
    return (_lib.TRAIN_FFMA_FC if env.get("MAPF_B200_TRAIN_FFMA_FC") else 0) | \
           (_lib.TRAIN_FFMA_GRAPH if env.get("MAPF_B200_TRAIN_FFMA_GRAPH") else 0) | \
           (_lib.TRAIN_FFMA_BWD if env.get("MAPF_B200_TRAIN_FFMA_BWD") else 0) | \
           (_lib.TRAIN_TC_BWD_ALWAYS if env.get("MAPF_B200_TRAIN_TC_BWD_ALWAYS") else 0) | \
           (_lib.TRAIN_FFMA_CONV if env.get("MAPF_B200_TRAIN_FFMA_CONV") else 0) | \
           (_lib.TRAIN_SINGLE_LANE if env.get("MAPF_B200_TRAIN_SINGLE_LANE") else 0) | \
           (_lib.TRAIN_BN_PASS if env.get("MAPF_B200_TRAIN_BN_PASS") else 0) | \
           (0 if env.get("MAPF_B200_TRAIN_SELF_CONTAINED") else _lib.TRAIN_PREPARED)   # forward prepares backward's operands


def wants_encoder_workspace(p) -> bool:
    """True when mapf_policy_fwd should get ``workspace_act`` (the pre-split FC operand) for these dimensions."""
    if not USE_TENSOR_CORES:
        return False
    if USE_TENSOR_CORE_FC:
        return True
    return USE_TENSOR_CORE_ENCODER and _lib.load().mapf_encoder_tc_supported(p) == 1

_LIB.define("pack_weights(int[] dims, Tensor[] conv_w, Tensor[] conv_b, Tensor[] bn_gamma, Tensor[] bn_beta, "
            "Tensor[] bn_mean, Tensor[] bn_var, Tensor fc_w, Tensor fc_b, Tensor? gf_w, Tensor? gf_w2, "
            "Tensor act_w, Tensor act_b) -> Tensor")
_LIB.define("policy_fwd(int[] dims, bool message, Tensor states, Tensor? gso, Tensor packed) -> (Tensor, Tensor)")
_LIB.define("encoder_fwd(int[] dims, Tensor states, Tensor packed) -> Tensor")
_LIB.define("pack_graph_weights(Tensor w, Tensor? w2) -> Tensor")
_LIB.define("graph_filter_fwd(Tensor x, Tensor gso, Tensor wt, int taps, bool add_self_loop, bool has_skip, "
            "bool relu) -> Tensor")
_LIB.define("head_fwd(Tensor x, Tensor act_w, Tensor act_b) -> (Tensor, Tensor)")
_LIB.define("linear(Tensor x, Tensor w, Tensor? b, bool relu) -> Tensor")
_LIB.define("env_step(Tensor(a!) pos, Tensor goal, Tensor? actions, Tensor(b!) cells, Tensor(c!) time, "
            "Tensor(d!) done, Tensor(e!) fov, Tensor(f!) adj, int board, int d2_limit, bool do_move) -> ()")


# training step (train.py:82-100): forward / backward of the train-mode network, cross-entropy, clip + Adam
_PARAM_SCHEMA = ("Tensor[] conv_w, Tensor[] conv_b, Tensor[] bn_gamma, Tensor[] bn_beta, {m}[] bn_mean, {v}[] bn_var, "
                 "Tensor fc_w, Tensor fc_b, Tensor? gf_w, Tensor? gf_w2, Tensor act_w, Tensor act_b")
_LIB.define("train_fwd(int[] dims, bool message, Tensor states, Tensor? gso, "
            + _PARAM_SCHEMA.format(m="Tensor(a!)", v="Tensor(b!)")
            + ", Tensor(c!)[] bn_batches, Tensor(d!) workspace, int flags) -> Tensor")
_LIB.define("train_bwd(int[] dims, bool message, Tensor states, Tensor? gso, "
            + _PARAM_SCHEMA.format(m="Tensor", v="Tensor")
            + ", Tensor(a!) workspace, Tensor dlogits, Tensor(b!)[] grads) -> ()")
_LIB.define("ce_loss(Tensor logits, Tensor targets) -> (Tensor, Tensor)")
_LIB.define("clip_adam(Tensor(a!) params, Tensor(b!) grads, Tensor(c!) exp_avg, Tensor(d!) exp_avg_sq, Tensor(e!) step, "
            "Tensor lr, float beta1, float beta2, float eps, float weight_decay, float max_norm, "
            "Tensor(f!) grad_norm, Tensor(g!) scratch) -> ()")


def net_params(dims, tensors=None) -> _lib.NetParams:
    p = _lib.NetParams()
    p.num_conv = int(dims[0])
    for i in range(_lib.MAX_CONV + 1):
        p.channels[i] = int(dims[1 + i])
    p.fc_in, p.fc_out, p.taps, p.node_dim, p.num_actions = (int(v) for v in dims[6:11])
    if tensors is not None:
        for l in range(p.num_conv):
            for name in ("conv_w", "conv_b", "bn_gamma", "bn_beta", "bn_mean", "bn_var"):
                getattr(p, name)[l] = tensors[name][l].data_ptr()
        for name in ("fc_w", "fc_b", "gf_w", "gf_w2", "act_w", "act_b"):
            t = tensors[name]
            setattr(p, name, None if t is None else t.data_ptr())
    return p


def _f32c(t):
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32 tensor, got {t.dtype}")
    return t.contiguous()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _pack_weights(dims, conv_w, conv_b, bn_gamma, bn_beta, bn_mean, bn_var, fc_w, fc_b, gf_w, gf_w2, act_w, act_b):
    lib = _lib.load()
    keep = {"conv_w": [_f32c(t) for t in conv_w], "conv_b": [_f32c(t) for t in conv_b],
            "bn_gamma": [_f32c(t) for t in bn_gamma], "bn_beta": [_f32c(t) for t in bn_beta],
            "bn_mean": [_f32c(t) for t in bn_mean], "bn_var": [_f32c(t) for t in bn_var],
            "fc_w": _f32c(fc_w), "fc_b": _f32c(fc_b), "gf_w": None if gf_w is None else _f32c(gf_w),
            "gf_w2": None if gf_w2 is None else _f32c(gf_w2), "act_w": _f32c(act_w), "act_b": _f32c(act_b)}
    p = net_params(dims, keep)
    size = lib.mapf_packed_weights_size(p)
    if size < 0:
        raise RuntimeError("mapf_b200 pack_weights: unsupported network dimensions")
    packed = torch.empty((size,), dtype=torch.float32, device=fc_w.device)
    with torch.cuda.device(fc_w.device):
        _lib.check(lib.mapf_pack_weights(p, packed.data_ptr(), _stream()), "pack_weights")
    return packed


def _policy_fwd(dims, message, states, gso, packed):
    lib = _lib.load()
    states = _f32c(states)
    B, N = int(states.shape[0]), int(states.shape[1])
    p = net_params(dims)
    dev = states.device
    logits = torch.empty((B, N, p.num_actions), dtype=torch.float32, device=dev)
    actions = torch.empty((B, N), dtype=torch.int32, device=dev)
    work = torch.empty((B * N, p.fc_out), dtype=torch.float32, device=dev)
    a = _lib.PolicyFwdArgs()
    a.batch, a.agents, a.message = B, N, int(message)
    a.states, a.packed, a.workspace_x = states.data_ptr(), packed.data_ptr(), work.data_ptr()
    if wants_encoder_workspace(p):
        worka = torch.empty((lib.mapf_tc_presplit_a_size(B * N, p.fc_in),), dtype=torch.float32, device=dev)
        a.workspace_act = worka.data_ptr()
    if p.taps > 0:
        if gso is None:
            raise ValueError("gso is required for the GNN variants")
        gso = _f32c(gso)
        if tuple(gso.shape) != (B, N, N):
            raise ValueError(f"gso must be [{B},{N},{N}], got {tuple(gso.shape)}")
        a.gso = gso.data_ptr()
        if USE_TENSOR_CORES:
            workz = torch.empty((B * N, (p.taps + 1) * p.fc_out), dtype=torch.float32, device=dev)
            a.workspace_z = workz.data_ptr()
    a.logits, a.actions = logits.data_ptr(), actions.data_ptr()
    a.flags = policy_flags()
    with torch.cuda.device(dev):
        _lib.check(lib.mapf_policy_fwd(p, a, _stream()), "policy_fwd")
    return logits, actions


def _encoder_fwd(dims, states, packed):
    lib = _lib.load()
    states = _f32c(states)
    p = net_params(dims)
    rows = states.numel() // 50
    x = torch.empty((rows, p.fc_out), dtype=torch.float32, device=states.device)
    a = _lib.EncoderFwdArgs(rows, states.data_ptr(), packed.data_ptr(), x.data_ptr())
    with torch.cuda.device(states.device):
        _lib.check(lib.mapf_encoder_fwd(p, a, _stream()), "encoder_fwd")
    return x


def _pack_graph_weights(w, w2):
    lib = _lib.load()
    w = _f32c(w)
    F, K, G = (int(v) for v in w.shape)
    w2c = None if w2 is None else _f32c(w2)
    wt = torch.empty(((K + (w2 is not None)) * F, G), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(lib.mapf_pack_graph_weights(w.data_ptr(), None if w2c is None else w2c.data_ptr(), F, K, G,
                                               wt.data_ptr(), _stream()), "pack_graph_weights")
    return wt


def _graph_filter_fwd(x, gso, wt, taps, add_self_loop, has_skip, relu):
    """x in the reference layout [B,F,N] (gnn.py:72) -> y [B,G,N]."""
    lib = _lib.load()
    x, gso, wt = _f32c(x), _f32c(gso), _f32c(wt)
    B, F, N = (int(v) for v in x.shape)
    G = int(wt.shape[1])
    y = torch.empty((B, G, N), dtype=torch.float32, device=x.device)
    a = _lib.GraphFilterFwdArgs()
    a.batch, a.agents, a.in_dim, a.out_dim, a.taps = B, N, F, G, int(taps)
    a.add_self_loop, a.has_skip, a.relu, a.num_actions = int(add_self_loop), int(has_skip), int(relu), 0
    a.x, a.x_sb, a.x_sn, a.x_sf = x.data_ptr(), F * N, 1, N
    a.gso, a.wt = gso.data_ptr(), wt.data_ptr()
    a.y, a.y_sb, a.y_sn, a.y_sg = y.data_ptr(), G * N, 1, N
    with torch.cuda.device(x.device):
        _lib.check(lib.mapf_graph_filter_fwd(a, _stream()), "graph_filter_fwd")
    return y


def _head_fwd(x, act_w, act_b):
    lib = _lib.load()
    x, act_w, act_b = _f32c(x), _f32c(act_w), _f32c(act_b)
    rows, D = int(x.shape[0]), int(x.shape[1])
    A = int(act_w.shape[0])
    logits = torch.empty((rows, A), dtype=torch.float32, device=x.device)
    actions = torch.empty((rows,), dtype=torch.int32, device=x.device)
    a = _lib.HeadFwdArgs(rows, D, A, x.data_ptr(), act_w.data_ptr(), act_b.data_ptr(), logits.data_ptr(),
                         actions.data_ptr())
    with torch.cuda.device(x.device):
        _lib.check(lib.mapf_head_fwd(a, _stream()), "head_fwd")
    return logits, actions


def _linear(x, w, b, relu):
    """act(x W^T + b) on the tensor cores (3xTF32, fp32-level accuracy): nn.Linear (+ ReLU) of the encoder / action MLPs
    (framework_gnn.py:90-96) for x [rows, K], W [N <= 128, K], K a multiple of 4."""
    lib = _lib.load()
    x, w = _f32c(x), _f32c(w)
    rows, K = int(x.shape[0]), int(x.shape[1])
    Nout = int(w.shape[0])
    if int(w.shape[1]) != K or Nout > 128 or (K & 3):
        raise ValueError(f"mapf::linear supports W [N <= 128, K % 4 == 0] matching x; got x {tuple(x.shape)}, W {tuple(w.shape)}")
    out = torch.empty((rows, Nout), dtype=torch.float32, device=x.device)
    pk = torch.empty((lib.mapf_tc_packed_b_size(Nout, K),), dtype=torch.float32, device=x.device)
    bias = None if b is None else _f32c(b)
    with torch.cuda.device(x.device):
        _lib.check(lib.mapf_tc_pack_b(w.data_ptr(), K, None, 0, Nout, pk.data_ptr(), _stream()), "tc_pack_b")
        _lib.check(lib.mapf_tc_gemm(rows, Nout, K, x.data_ptr(), K, pk.data_ptr(), out.data_ptr(), Nout,
                                    None if bias is None else bias.data_ptr(), int(relu), _stream()), "tc_gemm")
    return out


def _env_step(pos, goal, actions, cells, time, done, fov, adj, board, d2_limit, do_move):
    lib = _lib.load()
    E, N = int(pos.shape[0]), int(pos.shape[1])
    a = _lib.EnvStepArgs()
    a.episodes, a.agents, a.board, a.cell_stride = E, N, int(board), int(cells.shape[1])
    a.d2_limit, a.do_move = int(d2_limit), int(do_move)
    a.pos, a.goal, a.actions = pos.data_ptr(), goal.data_ptr(), None if actions is None else actions.data_ptr()
    a.cells, a.time, a.done, a.fov, a.adj = (cells.data_ptr(), time.data_ptr(), done.data_ptr(), fov.data_ptr(),
                                             adj.data_ptr())
    with torch.cuda.device(pos.device):
        _lib.check(lib.mapf_env_step(a, _stream()), "env_step")


def _train_tensors(conv_w, conv_b, bn_gamma, bn_beta, bn_mean, bn_var, fc_w, fc_b, gf_w, gf_w2, act_w, act_b):
    return {"conv_w": list(conv_w), "conv_b": list(conv_b), "bn_gamma": list(bn_gamma), "bn_beta": list(bn_beta),
            "bn_mean": list(bn_mean), "bn_var": list(bn_var), "fc_w": fc_w, "fc_b": fc_b, "gf_w": gf_w, "gf_w2": gf_w2,
            "act_w": act_w, "act_b": act_b}


def train_workspace_floats(dims, batch, agents) -> int:
    """Floats of caller-owned scratch a train-mode forward / backward pair needs (mapf_train_workspace_size)."""
    size = _lib.load().mapf_train_workspace_size(net_params(dims), int(batch), int(agents))
    if size < 0:
        raise RuntimeError("mapf_b200 train: unsupported network / batch dimensions")
    return int(size)


def _train_fwd(dims, message, states, gso, conv_w, conv_b, bn_gamma, bn_beta, bn_mean, bn_var, fc_w, fc_b, gf_w, gf_w2,
               act_w, act_b, bn_batches, workspace, flags):
    """Train-mode Network.forward (train.py:87): per-agent-slot BatchNorm statistics (framework_gnn.py:160-163);
    bn_mean / bn_var / bn_batches receive the N sequential running-stat updates; workspace keeps what backward needs."""
    lib = _lib.load()
    B, N = int(states.shape[0]), int(states.shape[1])
    p = net_params(dims, _train_tensors(conv_w, conv_b, bn_gamma, bn_beta, bn_mean, bn_var, fc_w, fc_b, gf_w, gf_w2,
                                        act_w, act_b))
    logits = torch.empty((B, N, p.num_actions), dtype=torch.float32, device=states.device)
    a = _lib.TrainFwdArgs()
    a.batch, a.agents, a.message, a.momentum = B, N, int(message), 0.1
    a.states, a.workspace, a.logits = states.data_ptr(), workspace.data_ptr(), logits.data_ptr()
    if p.taps > 0:
        a.gso = gso.data_ptr()
    for l, t in enumerate(bn_batches):
        a.bn_batches[l] = t.data_ptr()
    a.flags = int(flags)
    with torch.cuda.device(states.device):
        _lib.check(lib.mapf_train_forward(p, a, _stream()), "train_forward")
    return logits


def _train_bwd(dims, message, states, gso, conv_w, conv_b, bn_gamma, bn_beta, bn_mean, bn_var, fc_w, fc_b, gf_w, gf_w2,
               act_w, act_b, workspace, dlogits, grads):
    """Backward of train_fwd from d loss / d logits; `grads` in the order conv_w[], conv_b[], bn_gamma[], bn_beta[],
    fc_w, fc_b, (gf_w), (gf_w2), act_w, act_b is overwritten."""
    lib = _lib.load()
    B, N = int(states.shape[0]), int(states.shape[1])
    p = net_params(dims, _train_tensors(conv_w, conv_b, bn_gamma, bn_beta, bn_mean, bn_var, fc_w, fc_b, gf_w, gf_w2,
                                        act_w, act_b))
    a = _lib.TrainBwdArgs()
    a.batch, a.agents, a.message = B, N, int(message)
    a.states, a.workspace, a.dlogits = states.data_ptr(), workspace.data_ptr(), dlogits.data_ptr()
    if p.taps > 0:
        a.gso = gso.data_ptr()
    L = p.num_conv
    names = [("conv_w", i) for i in range(L)] + [("conv_b", i) for i in range(L)] + \
            [("bn_gamma", i) for i in range(L)] + [("bn_beta", i) for i in range(L)] + [("fc_w", None), ("fc_b", None)]
    if gf_w is not None:
        names.append(("gf_w", None))
    if gf_w2 is not None:
        names.append(("gf_w2", None))
    names += [("act_w", None), ("act_b", None)]
    if len(names) != len(grads):
        raise ValueError(f"train_bwd expects {len(names)} gradient buffers, got {len(grads)}")
    for (name, i), t in zip(names, grads):
        if i is None:
            setattr(a.grads, name, t.data_ptr())
        else:
            getattr(a.grads, name)[i] = t.data_ptr()
    a.flags = train_flags()
    with torch.cuda.device(states.device):
        _lib.check(lib.mapf_train_backward(p, a, _stream()), "train_backward")


def _ce_loss(logits, targets):
    """nn.CrossEntropyLoss (mean over B*N rows, train.py:90-93) -> (loss[1], d loss / d logits)."""
    lib = _lib.load()
    logits = _f32c(logits)
    A = int(logits.shape[-1])
    rows = logits.numel() // A
    targets = targets.to(dtype=torch.int32).reshape(rows).contiguous()
    loss = torch.empty((1,), dtype=torch.float32, device=logits.device)
    dlogits = torch.empty_like(logits)
    a = _lib.CeLossArgs(rows, A, logits.data_ptr(), targets.data_ptr(), loss.data_ptr(), dlogits.data_ptr())
    with torch.cuda.device(logits.device):
        _lib.check(lib.mapf_ce_loss(a, _stream()), "ce_loss")
    return loss, dlogits


def _clip_adam(params, grads, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, weight_decay, max_norm, grad_norm,
               scratch):
    """clip_grad_norm_(max_norm) + Adam with L2 weight decay (train.py:64,97,100) over flat fp32 buffers; `step`
    (int32[1]) and `lr` (float32[1]) live on the device so that a captured graph can be replayed unchanged."""
    lib = _lib.load()
    a = _lib.ClipAdamArgs()
    a.n = params.numel()
    a.params, a.grads = params.data_ptr(), grads.data_ptr()
    a.exp_avg, a.exp_avg_sq = exp_avg.data_ptr(), exp_avg_sq.data_ptr()
    a.lr, a.beta1, a.beta2, a.eps = 0.0, float(beta1), float(beta2), float(eps)
    a.weight_decay, a.max_norm, a.step = float(weight_decay), float(max_norm), 0
    a.grad_norm, a.scratch = grad_norm.data_ptr(), scratch.data_ptr()
    a.step_dev, a.lr_dev = step.data_ptr(), lr.data_ptr()
    with torch.cuda.device(params.device):
        _lib.check(lib.mapf_clip_adam(a, _stream()), "clip_adam")


def _ce_loss_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])
    ctx.mark_non_differentiable(output[1])


def _ce_loss_backward(ctx, g_loss, g_dlogits):
    (dlogits,) = ctx.saved_tensors
    return dlogits * g_loss.reshape(()), None


_LIB.impl("pack_weights", _pack_weights, "CUDA")
_LIB.impl("train_fwd", _train_fwd, "CUDA")
_LIB.impl("train_bwd", _train_bwd, "CUDA")
_LIB.impl("ce_loss", _ce_loss, "CUDA")
_LIB.impl("clip_adam", _clip_adam, "CUDA")
torch.library.register_autograd("mapf::ce_loss", _ce_loss_backward, setup_context=_ce_loss_setup, lib=_LIB)
_LIB.impl("policy_fwd", _policy_fwd, "CUDA")
_LIB.impl("encoder_fwd", _encoder_fwd, "CUDA")
_LIB.impl("pack_graph_weights", _pack_graph_weights, "CUDA")
_LIB.impl("graph_filter_fwd", _graph_filter_fwd, "CUDA")
_LIB.impl("head_fwd", _head_fwd, "CUDA")
_LIB.impl("linear", _linear, "CUDA")
_LIB.impl("env_step", _env_step, "CUDA")
