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
