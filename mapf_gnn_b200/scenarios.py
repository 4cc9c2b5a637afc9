"""Host-side synthetic scenario generation (numpy): the evaluate.py:207-208 order -- unique
obstacles, then unique obstacle-free goals -- plus unique obstacle-free starts (explicit
starts, because ``GraphEnv.reset``'s row/column draw :77-90 fails on dense maps).  Vectorised over
episodes so that 65 536-episode batches are generated in seconds.  Not on the timed path."""
from __future__ import annotations

import numpy as np


def make_scenarios(rng: np.random.Generator, episodes, num_agents, board_size, num_obstacles, chunk=4096,
                   goals_ignore_obstacles=False):
    """Returns starts int32[E,N,2], goals int32[E,N,2], obstacles int32[E,M,2] as (x, y).

    ``goals_ignore_obstacles=True`` is the train-time validation order (train.py:111-112): goals are drawn first and
    independently of the obstacles, so a goal may lie under an obstacle (that agent can then never finish)."""
    H = int(board_size)
    cells = H * H
    N, M = int(num_agents), int(num_obstacles)
    if M + N > cells:
        raise ValueError(f"Not enough cells ({cells}) for {M} obstacles and {N} agents")
    starts = np.empty((episodes, N, 2), dtype=np.int32)
    goals = np.empty((episodes, N, 2), dtype=np.int32)
    obst = np.empty((episodes, M, 2), dtype=np.int32)
    for lo in range(0, episodes, chunk):
        hi = min(episodes, lo + chunk)
        # two independent random orders of the cells: the first places obstacles then goals; starts take
        # the first N obstacle-free cells of the second order
        order = np.argsort(rng.random((hi - lo, cells), dtype=np.float32), axis=1)
        o, g = order[:, :M], order[:, M:M + N]
        if goals_ignore_obstacles:
            g = np.argsort(rng.random((hi - lo, cells), dtype=np.float32), axis=1)[:, :N]
        is_obst = np.zeros((hi - lo, cells), dtype=bool)
        np.put_along_axis(is_obst, o, True, axis=1)
        key = rng.random((hi - lo, cells), dtype=np.float32) + is_obst * 2.0
        s = np.argsort(key, axis=1)[:, :N]
        obst[lo:hi, :, 0], obst[lo:hi, :, 1] = o % H, o // H
        goals[lo:hi, :, 0], goals[lo:hi, :, 1] = g % H, g // H
        starts[lo:hi, :, 0], starts[lo:hi, :, 1] = s % H, s // H
    return starts, goals, obst


def make_scenarios_device(episodes, num_agents, board_size, num_obstacles, seed=0, device="cuda"):
    """Same contract as ``make_scenarios`` but generated on the GPU (one kernel, no host work): int32 device tensors
    starts [E,N,2], goals [E,N,2], obstacles [E,M,2].  Raises ValueError like ``create_goals`` when the board is too
    small (env_graph_gridv1.py:494-495)."""
    import torch
    from . import _lib
    lib = _lib.load()
    E, N, H, M = int(episodes), int(num_agents), int(board_size), int(num_obstacles)
    if M + N > H * H:
        raise ValueError(f"Not enough valid positions ({H * H - M}) for {N} agents")
    dev = torch.device(device)
    starts = torch.empty((E, N, 2), dtype=torch.int32, device=dev)
    goals = torch.empty((E, N, 2), dtype=torch.int32, device=dev)
    obst = torch.empty((E, M, 2), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mapf_make_scenarios(E, N, H, M, int(seed) & (2 ** 64 - 1), starts.data_ptr(), goals.data_ptr(),
                                           obst.data_ptr() if M else None, _lib.current_stream()), "make_scenarios")
    return starts, goals, obst
