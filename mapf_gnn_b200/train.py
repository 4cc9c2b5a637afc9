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

from . import _lib, ops
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
    """Caller-owned scratch for one (B, N) shape, reused across steps (the fused ``Trainer`` owns exactly one)."""

    def __init__(self):
        self.key, self.buf = None, None

    def get(self, dims, B, N, device):
        key = (B, N, device)
        if self.key != key:
            self.buf = torch.empty((ops.train_workspace_floats(dims, B, N),), dtype=torch.float32, device=device)
            self.key = key
        return self.buf


class _WorkspacePool:
    """Scratch buffers of the autograd path: every train-mode forward checks out its OWN workspace and its backward
    hands it back, so several forwards may be alive at once (gradient accumulation over micro-batches, two losses from
    two forwards) like with the reference's ordinary autograd graph.  A forward that never gets a backward simply drops
    its buffer."""

    def __init__(self, keep=4):
        self.free, self.keep = {}, keep

    def checkout(self, dims, B, N, device):
        lst = self.free.get((B, N, device))
        if lst:
            return lst.pop()
        return torch.empty((ops.train_workspace_floats(dims, B, N),), dtype=torch.float32, device=device)

    def release(self, B, N, device, buf):
        lst = self.free.setdefault((B, N, device), [])
        if len(lst) < self.keep:
            lst.append(buf)


def _op_params(keep):
    return (keep["conv_w"], keep["conv_b"], keep["bn_gamma"], keep["bn_beta"], keep["bn_mean"], keep["bn_var"],
            keep["fc_w"], keep["fc_b"], keep["gf_w"], keep["gf_w2"], keep["act_w"], keep["act_b"])


def _forward(net, states, gso, work):
    """``mapf::train_fwd`` on the network's current parameters; returns logits and what backward needs."""
    states = states.contiguous()
    p, keep = _raw_params(net)
    if p.taps > 0:
        gso = gso.to(torch.float32).contiguous()
    else:
        gso = None
    bns = [m.num_batches_tracked for m in net.convLayers if isinstance(m, torch.nn.BatchNorm2d)]
    logits = torch.ops.mapf.train_fwd(net.dims(), bool(net.message), states, gso, *_op_params(keep), bns, work,
                                      ops.train_flags())
    net._weights_epoch = getattr(net, "_weights_epoch", 0) + 1   # running stats changed behind torch's back
    return logits, (states, gso, work, keep)


def _backward(net, saved, dlogits, grads):
    states, gso, work, keep = saved
    torch.ops.mapf.train_bwd(net.dims(), bool(net.message), states, gso, *_op_params(keep), work, dlogits,
                             [t for _, t in grads])


class _PolicyTrainFn(torch.autograd.Function):
    """Autograd formula of the train-mode network: forward = ``mapf::train_fwd``, backward = ``mapf::train_bwd``."""

    @staticmethod
    def forward(ctx, net, states, gso, *params):
        if not hasattr(net, "_train_pool"):
            net._train_pool = _WorkspacePool()
        B, N = int(states.shape[0]), int(states.shape[1])
        work = net._train_pool.checkout(net.dims(), B, N, states.device)
        logits, saved = _forward(net, states.detach(), None if gso is None else gso.detach(), work)
        ctx.net, ctx.saved, ctx.key = net, saved, (B, N, states.device)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        net = ctx.net
        if ctx.saved is None:
            raise RuntimeError("mapf_b200: backward through the same train-mode forward twice (its workspace was "
                               "returned after the first backward; retain_graph is not supported)")
        slots = _param_list(net.param_groups())
        grads = [(slot, torch.empty_like(t)) for slot, t in slots]
        _backward(net, ctx.saved, dlogits.contiguous(), grads)
        net._train_pool.release(*ctx.key, ctx.saved[2])
        ctx.saved = None
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
    """nn.CrossEntropyLoss (mean over B*N rows, train.py:90-93) -> (loss[1], dlogits): ``mapf::ce_loss`` (differentiable
    with respect to the logits through its registered autograd formula)."""
    return torch.ops.mapf.ce_loss(logits, targets.to(device=logits.device))


class Trainer:
    """The reference's optimisation step (train.py:64-66,82-100) as one launch sequence:
    Adam(lr 3e-4, weight_decay 1e-4 as L2), CrossEntropyLoss(mean), clip_grad_norm_(1.0).

    Parameters are re-pointed at views of ONE flat fp32 buffer (same for gradients and the Adam moments) so that
    the clip + Adam update is a single kernel pair over <= 64 216 floats and the multi-GPU gradient all-reduce
    is one NCCL call over that buffer (batch rows sharded, parameters replicated, per-shard BatchNorm
    statistics like DDP of the reference).
    """

    def __init__(self, net, lr=3e-4, weight_decay=1e-4, max_norm=1.0, betas=(0.9, 0.999), eps=1e-8,
                 process_group=None, world_size=None, use_graph=False, peer_allreduce=None):
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
        self._graphs, self._graph_key = None, None
        self.ws = _Workspace()
        self.group = process_group
        if world_size is None:
            import torch.distributed as dist
            world_size = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.world = world_size
        # world > 1: the gradient all-reduce runs over NVLink peer memory inside the gradient-norm kernel
        # (mapf_peer_reduce_clip_adam: no NCCL call, capturable in the step's CUDA graph).  peer_allreduce=False keeps the
        # NCCL all-reduce of the flat gradient (also the fallback when CUDA IPC is unavailable, e.g. ranks on several nodes).
        self.peer = None
        self._parity = 0
        self.peer_grad_views = None
        if self.world > 1 and peer_allreduce is not False:
            import torch.distributed as dist
            try:
                if dist.get_backend(process_group) != "nccl":
                    raise RuntimeError("peer all-reduce needs CUDA ranks")
                from .parallel import PeerGradients
                self.peer = PeerGradients(total, dev, process_group)
                self.peer_grad_views = []
                for par in (0, 1):
                    views = [(slot, self.peer.grad[par][off:off + n].view_as(t))
                             for (slot, t), off, n in zip(self.slots, offs, sizes)]
                    self.peer_grad_views.append(views)
            except Exception as exc:          # noqa: BLE001
                if peer_allreduce is True:
                    raise
                import warnings
                warnings.warn(f"mapf_b200: peer-memory all-reduce unavailable ({exc}); using the NCCL all-reduce")
                self.peer = None

    def close(self):
        """Release the peer-mapped gradient buffers (world > 1); the trainer must not be stepped afterwards."""
        if self.peer is not None:
            self.peer.close()
            self.peer = None

    def set_lr(self, lr):
        """CosineAnnealingLR is stepped per epoch by the caller (train.py:65,151)."""
        self.lr = lr
        self.lr_dev.fill_(lr)

    def _step_body(self, states, gso, targets, parity=0):
        net = self.net
        B, N = int(states.shape[0]), int(states.shape[1])
        logits, saved = _forward(net, states, gso, self.ws.get(net.dims(), B, N, states.device))
        loss, dlogits = cross_entropy(logits, targets)
        if self.peer is not None:
            # the backward kernels write this step's gradient straight into the peer-mapped buffer of this parity; the
            # fused kernel sums all ranks' buffers over NVLink, leaves the mean in flat_grad and feeds clip + Adam
            _backward(net, saved, dlogits, self.peer_grad_views[parity])
            a = _lib.ClipAdamArgs()
            a.n = self.flat.numel()
            a.params, a.grads = self.flat.data_ptr(), self.flat_grad.data_ptr()
            a.exp_avg, a.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
            a.lr, a.beta1, a.beta2, a.eps = self.lr, self.betas[0], self.betas[1], self.eps
            a.weight_decay, a.max_norm, a.step = self.weight_decay, self.max_norm, 0
            a.grad_norm, a.scratch = self.grad_norm.data_ptr(), self.scratch.data_ptr()
            a.step_dev, a.lr_dev = self.step_dev.data_ptr(), self.lr_dev.data_ptr()
            with torch.cuda.device(self.flat.device):
                _lib.check(self.lib.mapf_peer_reduce_clip_adam(self.peer.args(parity), a, _stream()),
                           "peer_reduce_clip_adam")
            return loss
        _backward(net, saved, dlogits, self.grad_views)
        if self.world > 1:
            from .parallel import average_gradients
            average_gradients(self.flat_grad, self.group)
        torch.ops.mapf.clip_adam(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_dev, self.lr_dev,
                                 self.betas[0], self.betas[1], self.eps, self.weight_decay, self.max_norm,
                                 self.grad_norm, self.scratch)
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
        parity = self._parity
        if self.peer is not None:
            self._parity ^= 1
        if not self.use_graph:
            loss = self._step_body(states, gso, targets, parity)
        else:
            key = (tuple(states.shape), None if gso is None else tuple(gso.shape))
            if self._graph_key != key:
                self._capture(key, states, gso, targets)
            # staging copies into the graph's static inputs -- skipped for tensors that ARE those inputs (input_buffers():
            # three eager copies cost ~20 us of launch latency in front of a 0.2 ms step)
            if states.data_ptr() != self._g_states.data_ptr():
                self._g_states.copy_(states)
            if gso is not None and gso.data_ptr() != self._g_gso.data_ptr():
                self._g_gso.copy_(gso)
            if targets.data_ptr() != self._g_targets.data_ptr():
                self._g_targets.copy_(targets)
            self._graphs[parity].replay()
            loss = self._g_loss[parity]
        self.step_count += 1
        net._weights_epoch = getattr(net, "_weights_epoch", 0) + 1   # parameters changed behind torch's back
        return loss

    def input_buffers(self, states, gso, targets):
        """The static input tensors ``(states, gso, targets)`` of the captured step for this batch shape, initialised with
        the given batch (``use_graph`` only).  Write the next batch into them -- e.g. as the destination of the loader's
        host-to-device copies -- and pass them to ``step``, which then launches the graph without staging copies.
        ``targets`` is int32 (train.py:90 casts the loader's float action ids with ``.long()``; the values are 0..4)."""
        if not self.use_graph:
            raise RuntimeError("input_buffers: the trainer does not replay a CUDA graph (use_graph=False)")
        targets = targets.to(device=self.flat.device, dtype=torch.int32)
        key = (tuple(states.shape), None if gso is None else tuple(gso.shape))
        if self._graph_key != key:
            self._capture(key, states, gso, targets)
        else:
            self._g_states.copy_(states)
            if gso is not None:
                self._g_gso.copy_(gso)
            self._g_targets.copy_(targets)
        return self._g_states, self._g_gso, self._g_targets

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
            for par in (0, 1):          # executed for real on every rank: the peer epochs advance in lockstep
                self._step_body(self._g_states, self._g_gso, self._g_targets, par)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

        def restore():
            self.flat.copy_(snap[0]), self.exp_avg.copy_(snap[1]), self.exp_avg_sq.copy_(snap[2])
            self.step_dev.copy_(snap[3])
            for b, old in zip(self.net.buffers(), snap[4]):
                b.copy_(old)
        restore()
        self._graphs, self._g_loss = [], []
        for par in ((0, 1) if self.peer is not None else (0,)):     # one graph per gradient-buffer parity
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                self._g_loss.append(self._step_body(self._g_states, self._g_gso, self._g_targets, par))
            self._graphs.append(gph)
        if self.peer is None:
            self._graphs.append(self._graphs[0]), self._g_loss.append(self._g_loss[0])
        restore()       # capture does not execute, but keep the state exactly as before for safety
        self._graph_key = key


def fit(config, net=None, data_loader=None, results_dir=None, tests_episodes=25, device="cuda", seed=None, use_graph=True,
        log=print):
    """The epoch loop of train.py:56-165 on the kernels of this package: per epoch one pass over the loader
    (``Trainer.step`` = forward + CrossEntropyLoss + backward + clip_grad_norm_(1.0) + Adam(3e-4, weight_decay 1e-4),
    train.py:82-100), then ``tests_episodes`` validation rollouts advancing TOGETHER (``evaluate(validation=True)`` =
    the per-episode loop of train.py:105-141: ``max_steps`` steps, goals drawn independently of the obstacles), the best
    model by success rate saved as ``best_model.pt``, CosineAnnealingLR(T_max = epochs) stepped per epoch (train.py:65,151)
    and ``loss.npy / success_rate.npy / flow_time.npy / model.pt`` written at the end (train.py:153-165).

    config: the reference's YAML dict (``epochs``, network keys, ``train`` loader block, ``board_size``, ``max_steps`` ...).
    Returns ``{"loss", "success_rate", "flow_time", "best_success_rate", "net"}``."""
    import math
    import os

    import numpy as np

    from .dataset import GNNDataLoader
    from .evaluate import evaluate
    from .rollout import build_network
    if seed is not None:
        torch.manual_seed(seed)
        np.random.seed(seed)
    loader = data_loader if data_loader is not None else GNNDataLoader(config, device=device, seed=seed)
    net = (net if net is not None else build_network(config)).to(device)
    trainer = Trainer(net, lr=3e-4, weight_decay=1e-4, use_graph=use_graph)
    epochs = int(config["epochs"])
    if results_dir is not None:
        os.makedirs(results_dir, exist_ok=True)
    losses, rates, flows, best = [], [], [], 0.0
    try:
        for epoch in range(epochs):
            trainer.set_lr(3e-4 * (1.0 + math.cos(math.pi * epoch / epochs)) / 2.0)      # CosineAnnealingLR, eta_min = 0
            net.train()
            total = torch.zeros((), dtype=torch.float32, device=device)
            batches = 0
            full = None
            for states, trajectories, gso in loader.train_loader:
                full = states.shape[0] if full is None else full
                trainer.use_graph = use_graph and states.shape[0] == full      # the ragged last batch runs eagerly
                total += trainer.step(states, gso, trajectories).reshape(())
                batches += 1
            losses.append(float(total.item()) / max(1, batches))
            net.eval()
            summary, _ = evaluate(net, config, episodes=tests_episodes, seed=(seed or 0) * 100003 + epoch, device=device,
                                  validation=True)
            rates.append(summary["avg_success_rate"])
            flows.append(summary["avg_flow_time"])
            log(f"Epoch {epoch}  Loss: {losses[-1]:.4f}  Success rate: {rates[-1]:.3f}  Flow time: {flows[-1]:.2f}  "
                f"Learning rate: {trainer.lr:.6f}")
            if rates[-1] > best:
                best = rates[-1]
                if results_dir is not None:
                    torch.save(net.state_dict(), os.path.join(results_dir, "best_model.pt"))
    finally:
        trainer.close()
    if results_dir is not None:
        np.save(os.path.join(results_dir, "success_rate.npy"), np.array(rates))
        np.save(os.path.join(results_dir, "flow_time.npy"), np.array(flows))
        np.save(os.path.join(results_dir, "loss.npy"), np.array(losses))
        torch.save(net.state_dict(), os.path.join(results_dir, "model.pt"))
    return {"loss": losses, "success_rate": rates, "flow_time": flows, "best_success_rate": best, "net": net}
