"""Host-side multi-GPU plumbing: one process per GPU, episodes / batch rows sharded in
contiguous blocks, parameters replicated.  Rollout needs no data-path collective; training needs one
all-reduce of the flat gradient.  These helpers only touch ``torch.distributed`` (NCCL on the GPU box,
gloo in the CPU tests); all compute stays in the CUDA library."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(total: int, rank: int, world: int):
    """Contiguous block [lo, hi) of `total` units owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def aggregate_throughput(units: float, seconds: float, group=None, device="cpu"):
    """Whole-job throughput = (units of all ranks) / (max time over ranks); returns (value, max_seconds, units)."""
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(u[0]) / float(t[0]), float(t[0]), float(u[0])


def average_gradients(flat_grad: torch.Tensor, group=None):
    """In-place mean of the flat gradient over ranks (each rank's loss is a mean over its own equal-sized shard, so
    the global-batch mean gradient is the rank average).  NCCL supports AVG natively; gloo does not."""
    if not (dist.is_available() and dist.is_initialized()):
        return flat_grad
    world = dist.get_world_size(group)
    if world == 1:
        return flat_grad
    if flat_grad.is_cuda:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad.div_(world)
    return flat_grad


def gather_metrics(success: torch.Tensor, flow_time: torch.Tensor, group=None):
    """Per-episode metrics of every rank concatenated in rank order (end-of-rollout gather, a few bytes/episode)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return success, flow_time
    world = dist.get_world_size(group)
    s = [torch.empty_like(success) for _ in range(world)]
    f = [torch.empty_like(flow_time) for _ in range(world)]
    dist.all_gather(s, success, group=group)
    dist.all_gather(f, flow_time, group=group)
    return torch.cat(s), torch.cat(f)


class PeerGradients:
    """Peer-mapped gradient buffers for the one-shot NVLink all-reduce (``mapf_peer_reduce_clip_adam``): per rank one CUDA
    IPC allocation ``[2 parities x n floats | 16 int32 flags]``, handles exchanged once with ``all_gather_object``, every
    peer's block mapped into this process.  ``grad[parity]`` are torch views of the local block (the backward kernels
    write the gradient there directly); ``args(parity)`` fills the C struct.  Single node only (P2P over NVLink)."""

    def __init__(self, n_floats: int, device, group=None):
        import ctypes as C

        from . import _lib
        self.lib = _lib.load()
        self.n = int(n_floats)
        assert self.n % 4 == 0
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise RuntimeError("peer all-reduce supports up to 8 ranks (one NVSwitch box)")
        self.device = torch.device(device)
        self.bytes = 2 * self.n * 4 + 64
        handle = C.create_string_buffer(64)
        ptr = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mapf_peer_alloc(self.bytes, C.byref(ptr), handle), "peer_alloc")
        self.local_ptr = ptr.value
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.local_ptr)
                continue
            p = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(self.lib.mapf_peer_open(h, C.byref(p)), "peer_open")
            self.ptrs.append(p.value)
        self.grad = [self._view(self.local_ptr + par * self.n * 4, self.n) for par in (0, 1)]
        dist.barrier(group=group)        # every rank has mapped every block before anyone signals

    def _view(self, ptr, n):
        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}
        t = torch.as_tensor(raw, device=self.device)
        t._mapf_keepalive = self          # the allocation must outlive the view
        return t

    def args(self, parity: int):
        from . import _lib
        a = _lib.PeerReduceArgs()
        a.world, a.rank = self.world, self.rank
        for r in range(self.world):
            a.peer_grads[r] = self.ptrs[r] + parity * self.n * 4
            a.peer_flags[r] = self.ptrs[r] + 2 * self.n * 4
        return a

    def close(self):
        if self.ptrs is None:
            return
        torch.cuda.synchronize(self.device)
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                self.lib.mapf_peer_close(p)
        self.lib.mapf_peer_free(self.local_ptr)
        self.ptrs = None
