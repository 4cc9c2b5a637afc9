"""Training step on the GPU: train-mode forward/backward of the policy networks as a custom autograd
function (so ``train.py:82-100`` of the reference runs unchanged on top of it) and ``Trainer``, the fused
step (forward -> cross-entropy -> backward -> [NCCL all-reduce] -> clip + Adam) used by the benchmark.

Every number is produced by libmapf_b200.so (``mapf_train_forward / mapf_train_backward / mapf_ce_loss /
mapf_clip_adam``); torch only owns the buffers, the stream and -- when world_size > 1 -- the NCCL
all-reduce of the flat gradient.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .ops import net_params

_GRAD_ORDER = ("conv_w", "conv_b", "bn_gamma", "bn_beta", "fc_w", "fc_b", "gf_w", "gf_w2", "act_w", "act_b")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _param_list(groups):
    """Trainable tensors in a fixed order + the (name, index) slot of each."""
    out = []
    for name in _GRAD_ORDER:
        v = groups[name]
        if isinstance(v, list):
            out += [((name, i), t) for i, t in enumerate(v)]
        elif v is not None:
            out.append(((name, None), v))
    return out


def _raw_params(net):
    g = net.param_groups()
    detached = {k: ([t.detach() for t in v] if isinstance(v, list) else (None if v is None else v.detach()))
                for k, v in g.items()}
    return net_params(net.dims(), detached), detached


class _Workspace:
    """Caller-owned scratch for one (B, N) shape, reused across steps."""

    def __init__(self):
        self.key, self.buf = None, None

    def get(self, lib, p, B, N, device):
        key = (B, N, device)
        if self.key != key:
            size = lib.mapf_train_workspace_size(p, B, N)
            if size < 0:
                raise RuntimeError("mapf_b200 train: unsupported network / batch dimensions")
            self.buf = torch.empty((size,), dtype=torch.float32, device=device)
            self.key = key
        return self.buf


def _forward(net, states, gso, ws):
    lib = _lib.load()
    states = states.contiguous()
    B, N = int(states.shape[0]), int(states.shape[1])
    p, keep = _raw_params(net)
    work = ws.get(lib, p, B, N, states.device)
    logits = torch.empty((B, N, p.num_actions), dtype=torch.float32, device=states.device)
    a = _lib.TrainFwdArgs()
    a.batch, a.agents, a.message, a.momentum = B, N, int(net.message), 0.1
    a.states, a.workspace, a.logits = states.data_ptr(), work.data_ptr(), logits.data_ptr()
    if p.taps > 0:
        gso = gso.to(torch.float32).contiguous()
        a.gso = gso.data_ptr()
    bns = [m for m in net.convLayers if isinstance(m, torch.nn.BatchNorm2d)]
    for l, bn in enumerate(bns):
        a.bn_batches[l] = bn.num_batches_tracked.data_ptr()
    from . import ops
    a.flags = ops.train_flags()
    with torch.cuda.device(states.device):
        _lib.check(lib.mapf_train_forward(p, a, _stream()), "train_forward")
    net._weights_epoch = getattr(net, "_weights_epoch", 0) + 1   # running stats changed behind torch's back
    return logits, (states, gso, work, keep)


def _backward(net, saved, dlogits, grads):
    lib = _lib.load()
    states, gso, work, keep = saved
    B, N = int(states.shape[0]), int(states.shape[1])
    p = net_params(net.dims(), keep)
    a = _lib.TrainBwdArgs()
    a.batch, a.agents, a.message = B, N, int(net.message)
    a.states, a.workspace, a.dlogits = states.data_ptr(), work.data_ptr(), dlogits.data_ptr()
    if p.taps > 0:
        a.gso = gso.data_ptr()
    for (name, i), t in grads:
        if i is None:
            setattr(a.grads, name, t.data_ptr())
        else:
            getattr(a.grads, name)[i] = t.data_ptr()
    with torch.cuda.device(states.device):
        _lib.check(lib.mapf_train_backward(p, a, _stream()), "train_backward")


class _PolicyTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, states, gso, *params):
        if not hasattr(net, "_train_ws"):
            net._train_ws = _Workspace()
        logits, saved = _forward(net, states.detach(), None if gso is None else gso.detach(), net._train_ws)
        ctx.net, ctx.saved = net, saved
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        net = ctx.net
        slots = _param_list(net.param_groups())
        grads = [(slot, torch.empty_like(t)) for slot, t in slots]
        _backward(net, ctx.saved, dlogits.contiguous(), grads)
        return (None, None, None, *[g for _, g in grads])


def policy_train_forward(net, states, gso):
    """Train-mode ``Network.forward`` (per-agent-slot BatchNorm statistics, running stats updated) with
    autograd support: ``loss.backward()`` fills ``.grad`` of every parameter from the backward kernels."""
    params = [t for _, t in _param_list(net.param_groups())]
    return _PolicyTrainFn.apply(net, states, gso, *params)


def graph_filter_backward(x, gso, W, W2, self_loop, gy):
    """Backward of the standalone GCNLayer / MessagePassingLayer in the reference layouts."""
    lib = _lib.load()
    x, gso, gy = x.detach().contiguous(), gso.detach().to(torch.float32).contiguous(), gy.contiguous()
    Wd = W.detach().contiguous()
    W2d = None if W2 is None else W2.detach().contiguous()
    B, F, N = (int(v) for v in x.shape)
    K, G = int(W.shape[1]), int(W.shape[2])
    size = lib.mapf_graph_filter_bwd_workspace_size(B, N, F, G, K, int(W2 is not None))
    work = torch.empty((size,), dtype=torch.float32, device=x.device)
    dx, dW = torch.empty_like(x), torch.empty_like(Wd)
    dW2 = None if W2 is None else torch.empty_like(W2d)
    a = _lib.GraphFilterBwdArgs()
    a.batch, a.agents, a.in_dim, a.out_dim, a.taps, a.add_self_loop = B, N, F, G, K, int(self_loop)
    a.x, a.gso, a.w, a.w2, a.dy = x.data_ptr(), gso.data_ptr(), Wd.data_ptr(), _lib.ptr(W2d), gy.data_ptr()
    a.workspace, a.dx, a.dw, a.dw2 = work.data_ptr(), dx.data_ptr(), dW.data_ptr(), _lib.ptr(dW2)
    with torch.cuda.device(x.device):
        _lib.check(lib.mapf_graph_filter_bwd(a, _stream()), "graph_filter_bwd")
    return dx, dW, dW2


def cross_entropy(logits, targets):
    """nn.CrossEntropyLoss (mean over B*N rows, train.py:90-93) -> (loss[1], dlogits)."""
    lib = _lib.load()
    logits = logits.contiguous()
    A = int(logits.shape[-1])
    rows = logits.numel() // A
    targets = targets.to(device=logits.device, dtype=torch.int32).reshape(rows).contiguous()
    loss = torch.empty((1,), dtype=torch.float32, device=logits.device)
    dlogits = torch.empty_like(logits)
    a = _lib.CeLossArgs(rows, A, logits.data_ptr(), targets.data_ptr(), loss.data_ptr(), dlogits.data_ptr())
    with torch.cuda.device(logits.device):
        _lib.check(lib.mapf_ce_loss(a, _stream()), "ce_loss")
    return loss, dlogits


class Trainer:
    """The reference's optimisation step (train.py:64-66,82-100) as one launch sequence:
    Adam(lr 3e-4, weight_decay 1e-4 as L2), CrossEntropyLoss(mean), clip_grad_norm_(1.0).

    Parameters are re-pointed at views of ONE flat fp32 buffer (same for gradients and the Adam moments) so that
    the clip + Adam update is a single kernel pair over <= 64 216 floats and the multi-GPU gradient all-reduce
    is one NCCL call over that buffer (batch rows sharded, parameters replicated, per-shard BatchNorm
    statistics like DDP of the reference).
    """

    def __init__(self, net, lr=3e-4, weight_decay=1e-4, max_norm=1.0, betas=(0.9, 0.999), eps=1e-8,
                 process_group=None, world_size=None, use_graph=False):
        self.net, self.lib = net, _lib.load()
        self.lr, self.weight_decay, self.max_norm, self.betas, self.eps = lr, weight_decay, max_norm, betas, eps
        self.slots = _param_list(net.param_groups())
        dev = self.slots[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("Trainer needs the network on a CUDA device (no CPU fallback)")
        sizes = [t.numel() for _, t in self.slots]
        offs, total = [], 0
        for n in sizes:
            offs.append(total)
            total += (n + 3) // 4 * 4
        self.flat = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.grad_views = []
        with torch.no_grad():
            for (slot, t), off, n in zip(self.slots, offs, sizes):
                view = self.flat[off:off + n].view_as(t)
                view.copy_(t)
                t.data = view
                gview = self.flat_grad[off:off + n].view_as(t)
                t.grad = gview
                self.grad_views.append((slot, gview))
        self.grad_norm = torch.zeros((1,), dtype=torch.float32, device=dev)
        self.scratch = torch.zeros((1,), dtype=torch.float64, device=dev)
        self.step_count = 0
        self.step_dev = torch.zeros((1,), dtype=torch.int32, device=dev)     # device-resident Adam step count
        self.lr_dev = torch.full((1,), lr, dtype=torch.float32, device=dev)
        self.use_graph = use_graph
        self._graph, self._graph_key = None, None
        self.ws = _Workspace()
        self.group = process_group
        if world_size is None:
            import torch.distributed as dist
            world_size = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.world = world_size

    def set_lr(self, lr):
        """CosineAnnealingLR is stepped per epoch by the caller (train.py:65,151)."""
        self.lr = lr
        self.lr_dev.fill_(lr)

    def _step_body(self, states, gso, targets):
        net = self.net
        logits, saved = _forward(net, states, gso, self.ws)
        loss, dlogits = cross_entropy(logits, targets)
        _backward(net, saved, dlogits, self.grad_views)
        if self.world > 1:
            from .parallel import average_gradients
            average_gradients(self.flat_grad, self.group)
        a = _lib.ClipAdamArgs()
        a.n = self.flat.numel()
        a.params, a.grads = self.flat.data_ptr(), self.flat_grad.data_ptr()
        a.exp_avg, a.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        a.lr, a.beta1, a.beta2, a.eps = self.lr, self.betas[0], self.betas[1], self.eps
        a.weight_decay, a.max_norm, a.step = self.weight_decay, self.max_norm, 0
        a.grad_norm, a.scratch = self.grad_norm.data_ptr(), self.scratch.data_ptr()
        a.step_dev, a.lr_dev = self.step_dev.data_ptr(), self.lr_dev.data_ptr()
        with torch.cuda.device(self.flat.device):
            _lib.check(self.lib.mapf_clip_adam(a, _stream()), "clip_adam")
        return loss

    @torch.no_grad()
    def step(self, states, gso, targets):
        """One optimisation step; returns the (local) mean cross-entropy as a device scalar.  With ``use_graph`` the
        launch sequence (~50 kernels) is captured once per input shape into a CUDA graph and replayed: the small-batch
        regime of train.py (batch 128) is launch-bound otherwise."""
        net = self.net
        if not net.training:
            raise RuntimeError("Trainer.step needs the network in train() mode (train.py:80)")
        targets = targets.to(device=self.flat.device, dtype=torch.int32)
        if not self.use_graph:
            loss = self._step_body(states, gso, targets)
        else:
            key = (tuple(states.shape), None if gso is None else tuple(gso.shape))
            if self._graph_key != key:
                self._capture(key, states, gso, targets)
            self._g_states.copy_(states)
            if gso is not None:
                self._g_gso.copy_(gso)
            self._g_targets.copy_(targets)
            self._graph.replay()
            loss = self._g_loss
        self.step_count += 1
        net._weights_epoch = getattr(net, "_weights_epoch", 0) + 1   # parameters changed behind torch's back
        return loss

    def _capture(self, key, states, gso, targets):
        self._g_states = states.clone()
        self._g_gso = None if gso is None else gso.to(torch.float32).clone()
        self._g_targets = targets.clone()
        # everything the replay touches must already exist: allocate the workspace with an eager forward on a copy
        # of the optimiser state, then restore it (warm-up must not count as training steps)
        snap = (self.flat.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.step_dev.clone(),
                [b.clone() for b in self.net.buffers()])
        side = torch.cuda.Stream(device=self.flat.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._step_body(self._g_states, self._g_gso, self._g_targets)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

        def restore():
            self.flat.copy_(snap[0]), self.exp_avg.copy_(snap[1]), self.exp_avg_sq.copy_(snap[2])
            self.step_dev.copy_(snap[3])
            for b, old in zip(self.net.buffers(), snap[4]):
                b.copy_(old)
        restore()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._g_loss = self._step_body(self._g_states, self._g_gso, self._g_targets)
        restore()       # capture does not execute, but keep the state exactly as before for safety
        self._graph_key = key
