"""Batched evaluator: the per-episode loop of the reference's ``evaluate.py:201-271`` --
obstacles, then goals, then ``max_steps + 10`` policy/env steps with early stop on done, then
``computeMetrics`` -- run for all episodes at once on the GPU.  Per-episode results follow the reference's
``EpisodeStats`` fields (``evaluate.py:101-123``); the summary follows ``print_results`` (:333-376).

Scenario generation is host-side numpy by default (``scenarios.make_scenarios``: unique obstacle-free goals and
starts; the reference's ``reset`` draws starts per row/column with replacement, :77-90, which fails on dense maps)
or one kernel on the GPU with ``device_scenarios=True`` -- same distribution family, not the same random stream.
"""
from __future__ import annotations

import numpy as np
import torch

from .env import BatchedGraphEnv
from .rollout import Rollout
from .scenarios import make_scenarios, make_scenarios_device


@torch.no_grad()
def evaluate(net, config, episodes=None, seed=0, extra_steps=10, device="cuda", device_scenarios=False):
    """Returns (summary dict, per-episode dict of numpy arrays)."""
    episodes = int(episodes if episodes is not None else config.get("tests_episodes", 100))
    N, H = int(config["num_agents"]), int(config["board_size"][0])
    if device_scenarios:
        starts, goals, obst = make_scenarios_device(episodes, N, H, int(config.get("obstacles", 0)), seed, device)
    else:
        starts, goals, obst = make_scenarios(np.random.default_rng(seed), episodes, N, H, int(config.get("obstacles", 0)))
    env = BatchedGraphEnv(config, goals, starts, obst if obst.shape[1] else None,
                          sensing_range=config.get("sensing_range", 6), device=device)
    net = net.to(device).eval()
    ro = Rollout(net, env)
    max_steps = int(config["max_steps"]) + int(extra_steps)          # evaluate.py:221
    ro.run(max_steps)                                                 # done episodes freeze (early `break`, :251)
    success, flow = env.computeMetrics()
    success, flow = success.cpu().numpy(), flow.cpu().numpy()
    steps = env.time.cpu().numpy()                                    # steps actually taken (time stops at done)
    done = env.done.cpu().numpy().astype(bool)
    per_episode = {"success_rate": success, "flow_time": flow, "steps_taken": steps, "completed": done,
                   "agents_at_goal": np.rint(success * N).astype(np.int64)}
    summary = {
        "episodes": episodes,
        "avg_success_rate": float(success.mean()),
        "complete_success": int(done.sum()),
        "complete_success_rate": float(done.mean()),
        "avg_flow_time": float(flow.mean()),
        "avg_steps": float(steps.mean()),
        "total_agents": episodes * N,
        "agents_at_goal": int(per_episode["agents_at_goal"].sum()),
    }
    return summary, per_episode
