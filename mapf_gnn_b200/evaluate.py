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
def evaluate(net, config, episodes=None, seed=0, extra_steps=10, device="cuda", device_scenarios=False,
             validation=False):
    """Returns (summary dict, per-episode dict of numpy arrays).  ``validation=True`` is the loop train.py runs after
    each epoch (train.py:105-141): ``max_steps`` steps (no +10) and goals drawn independently of the obstacles."""
    if validation:
        extra_steps, device_scenarios = 0, False
    episodes = int(episodes if episodes is not None else config.get("tests_episodes", 100))
    N, H = int(config["num_agents"]), int(config["board_size"][0])
    if device_scenarios:
        starts, goals, obst = make_scenarios_device(episodes, N, H, int(config.get("obstacles", 0)), seed, device)
    else:
        starts, goals, obst = make_scenarios(np.random.default_rng(seed), episodes, N, H, int(config.get("obstacles", 0)),
                                             goals_ignore_obstacles=validation)
    env = BatchedGraphEnv(config, goals, starts, obst if obst.shape[1] else None,
                          sensing_range=config.get("sensing_range", 6), device=device)
    net = net.to(device).eval()
    ro = Rollout(net, env)
    max_steps = int(config["max_steps"]) + int(extra_steps)          # evaluate.py:221
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    ro.run(max_steps)                                                 # done episodes freeze (early `break`, :251)
    t1.record()
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
        "std_success_rate": float(success.std()),
        "batch_time_s": t0.elapsed_time(t1) * 1e-3,      # all episodes advance together: one clock for the batch
    }
    return summary, per_episode


def results_dict(config, summary, per_episode):
    """The JSON document ``Evaluator.save_results`` writes (evaluate.py:401-422, EpisodeStats.to_dict :113-123).
    Per-episode times do not exist for a batched rollout: ``episode_time_s`` / ``inference_time_ms`` carry the batch
    wall time divided by the number of episodes (resp. by episodes x steps); ``collisions`` is 0 as in the reference,
    which never increments it."""
    from datetime import datetime
    E = int(summary["episodes"])
    per_ep_s = summary.get("batch_time_s", 0.0) / max(E, 1)
    steps = per_episode["steps_taken"]
    episodes = [{
        "success_rate": float(per_episode["success_rate"][e]),
        "flow_time": int(per_episode["flow_time"][e]),
        "steps_taken": int(steps[e]),
        "inference_time_ms": float(per_ep_s * 1e3 / max(int(steps[e]), 1)),
        "episode_time_s": float(per_ep_s),
        "collisions": 0,
        "agents_at_goal": int(per_episode["agents_at_goal"][e]),
        "total_agents": int(config["num_agents"]),
    } for e in range(E)]
    return {
        "timestamp": datetime.now().isoformat(),
        "config": {"model": config.get("exp_name"), "net_type": config.get("net_type", "gnn"),
                   "board_size": list(config["board_size"]), "num_agents": int(config["num_agents"]),
                   "obstacles": int(config.get("obstacles", 0)), "episodes": E},
        "summary": {"complete_success": int(summary["complete_success"]),
                    "avg_success_rate": float(summary["avg_success_rate"]),
                    "std_success_rate": float(summary["std_success_rate"]),
                    "avg_flow_time": float(summary["avg_flow_time"]), "avg_steps": float(summary["avg_steps"]),
                    "avg_inference_ms": float(np.mean([ep["inference_time_ms"] for ep in episodes])) if episodes else 0.0},
        "episodes": episodes,
    }
