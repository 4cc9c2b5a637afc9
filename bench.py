#!/usr/bin/env python
"""Benchmark of the batched decentralized-policy step (BASELINE.json metric: agent-steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2a] [--impl mine|reference]

A "step" is one pass of the hot path over one batch of E resident episodes: policy forward
(CNN encoder -> graph filter -> head -> argmax) followed by the environment step (move, collision
rules, closest-4 adjacency, FOV).  ``value`` = agent-steps/s with all state resident in HBM, timed with
CUDA events per step (L2 flushed between steps, outside the event pairs), max over ranks, summed over
ranks' episodes (weak scaling: every GPU owns E episodes, no data-path collective).  ``e2e`` is the same
step driven from HOST buffers: positions go up from pinned memory, actions / positions / done come back,
every step.  ``--impl reference`` times the CPU port of the reference (oracle/) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic forward FLOPs (2*MAC) per agent-step and bytes per agent-step of the FOV+adjacency kernel
WORKLOADS = {
    "c1": dict(model="gnn_k3", E=4096, N=5, H=18, M=5, r=6, flops=347.7e3, enc_flops=296.0e3, env_bytes=301,
               desc="GNN K=3 (gnn_k3), 5 agents, 18x18, 5 obstacles"),
    "c2a": dict(model="gnn_k2", E=4096, N=5, H=28, M=8, r=6, flops=330.7e3, enc_flops=296.0e3, env_bytes=393,
                desc="GNN K=2 (gnn_k2), 4096 episodes, 5 agents, 28x28"),
    "c2b": dict(model="baseline", E=4096, N=5, H=28, M=8, r=6, flops=426.2e3, enc_flops=425.6e3, env_bytes=393,
                desc="baseline CNN policy, 4096 episodes, 5 agents, 28x28"),
    "c3": dict(model="gnn_msg_k3", E=4096, N=20, H=32, M=40, r=6, flops=367.9e3, enc_flops=296.0e3, env_bytes=347,
               desc="message-passing GNN K=3 (gnn_msg_k3), 20 agents, 32x32"),
    "c5": dict(model="gnn_k4_seed4", E=8192, N=64, H=64, M=200, r=6, flops=387.4e3, enc_flops=296.0e3, env_bytes=536,
               desc="GNN K=4 (random init seed 4), 64 agents, 64x64, 8192 episodes per GPU"),
}
MODEL_CFG = {
    "gnn_k3": dict(net_type="gnn", msg_type="gcn", graph_filters=[3], channels=[16, 16, 16]),
    "gnn_k2": dict(net_type="gnn", msg_type="gcn", graph_filters=[2], channels=[16, 16, 16]),
    "gnn_msg_k3": dict(net_type="gnn", msg_type="message", graph_filters=[3], channels=[16, 16, 16]),
    "baseline": dict(net_type="baseline", channels=[32, 16, 16], graph_filters=[3]),
    "gnn_k4_seed4": dict(net_type="gnn", msg_type="gcn", graph_filters=[4], channels=[16, 16, 16]),
}
EPISODE_LEN = 32   # max_steps of every config: environments are re-reset every 32 steps (outside the timers)


def model_config(w):
    return dict(map_shape=[5, 5], num_actions=5, last_convs=[400], node_dims=[128], encoder_dims=[64],
                encoder_layers=1, action_layers=1, min_time=1, max_time=32, num_agents=w["N"],
                board_size=[w["H"], w["H"]], device="cuda", **MODEL_CFG[w["model"]])


def load_weights(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", f"weights_{name}.npz"))
    return {k: z[k] for k in z.files}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": float(p["hbm_gbs"]), "sm_max_mhz": float(p.get("sm_max_mhz", 1965.0)),
                "tensor_tflops": float(p.get("bf16_tflops", 1590.0)), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "tensor_tflops": 1590.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 25 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "25"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])), mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference (R2 protocol of SURVEY 8d: python loop over reference-style
# single-episode envs + one batched torch-CPU forward using every host thread)
# ----------------------------------------------------------------------------------------------
def cpu_reference_steps(w, episodes, steps, warmup, seed=0):
    import torch
    from mapf_gnn_b200.scenarios import make_scenarios
    from oracle import env_oracle as eo
    from oracle import model_oracle as mo
    rng = np.random.default_rng(seed)
    N, H = w["N"], w["H"]
    starts, goals, obst = make_scenarios(rng, episodes, N, H, w["M"])
    params = {k: torch.from_numpy(v) for k, v in load_weights(w["model"]).items()}
    envs = [eo.EnvOracle(N, H, goals[e], starts[e], obst[e], sensing_range=w["r"]) for e in range(episodes)]
    obs = [e.observe() for e in envs]
    total = 0.0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        states = torch.tensor(np.stack([o["fov"] for o in obs])).float()
        gso = torch.tensor(np.stack([o["adj_matrix"] for o in obs])).float()
        with torch.no_grad():
            logits = mo.forward(params, states, gso)
        acts = mo.greedy_actions(logits)
        obs = [envs[e].step(acts[e])[0] for e in range(episodes)]
        dt = time.perf_counter() - t0
        if it >= warmup:
            total += dt
    return episodes * N * steps / total, total


def run_reference(args, w):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)   # torchrun pins OMP_NUM_THREADS=1; the CPU arm may use every core
    episodes = args.ref_episodes
    value, total = cpu_reference_steps(w, episodes, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "agent-steps/sec (batched rollout)", "value": value, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['desc']}", "episodes_per_step": episodes, "agents": w["N"],
                   "board": w["H"]},
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{episodes} episodes x {args.steps} steps of the same workload "
                                   f"(python loop over single-episode numpy envs + batched torch-CPU forward)"},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_mine(args, w):
    import torch
    import torch.distributed as dist
    import mapf_gnn_b200 as m

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    E, N, H = args.episodes or w["E"], w["N"], w["H"]
    rng = np.random.default_rng(1000 + rank)
    starts, goals, obst = m.make_scenarios(rng, E, N, H, w["M"])
    cfg = model_config(w)
    net = m.build_network(cfg)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in load_weights(w["model"]).items()})
    net = net.to(dev).eval()
    env = m.BatchedGraphEnv(cfg, goals, starts, obst, sensing_range=w["r"], device=dev)
    ro = m.Rollout(net, env)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, args.warmup
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    active = torch.zeros(K, dtype=torch.int64, device=dev)
    for i in range(W):
        ro.run(1)
    # the timed step is ONE CUDA-graph replay of the step's launch sequence (same kernels, same order): the host then
    # issues one launch per step and cannot open gaps between the kernels when 8 ranks share the box's cores
    step_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(step_graph):
        ro.run(1)
    env.reset()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    for i in range(K):
        if i % EPISODE_LEN == 0 and i > 0:
            env.reset()
        active[i] = E - env.done.sum()
        flush.zero_()
        ev0[i].record()
        step_graph.replay()
        ev1[i].record()
    barrier()
    wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    units = float(active.sum().item()) * N
    t = torch.tensor([dev_ms, units, float(E) * N * K], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dev_ms_max = float(tmax[0])
    else:
        dev_ms_max = dev_ms
    units_all, units_nominal = float(t[1]), float(t[2])
    value = units_all / (dev_ms_max * 1e-3)

    # ---- dominant kernel timed alone, live, on the launching stream: the conv stack of the encoder, on the tensor cores
    # (encoder_tc_kernel) for 16-channel networks, else inside the fused FFMA2 encoder kernel (encoder_fwd_kernel) ----
    dims = net.dims()
    packed = net.packed_weights()
    from mapf_gnn_b200 import _lib as _mlib
    tc_encoder = ro.worka is not None and _mlib.load().mapf_encoder_tc_supported(ro.params) == 1
    if tc_encoder:
        _ea = _mlib.EncoderFwdArgs(E * N, env.fov.data_ptr(), packed.data_ptr(), ro.worka.data_ptr())
        _st = _mlib.current_stream()

        def dominant():
            _mlib.check(_mlib.load().mapf_encoder_tc_conv_fwd(ro.params, _ea, _st), "encoder_tc_conv_fwd")
    else:
        def dominant():
            torch.ops.mapf.encoder_fwd(dims, env.fov, packed)
    enc_ms = []
    for i in range(3 + 10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dominant()
        b.record()
        b.synchronize()
        if i >= 3:
            enc_ms.append(a.elapsed_time(b))
    enc_ms = float(np.mean(enc_ms))
    env_ms = []
    for i in range(3 + 10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        env.observe()
        b.record()
        b.synchronize()
        if i >= 3:
            env_ms.append(a.elapsed_time(b))
    env_ms = float(np.mean(env_ms))
    pk = peaks()
    traffic = None
    try:       # ncu-measured DRAM bytes per launch of the dominant kernel (committed summary; null when not captured)
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
            rec = json.load(fh).get(args.workload, {})
            want = "encoder_tc_kernel" if tc_encoder else "encoder_fwd_kernel"
            rec = rec.get(want, rec if rec.get("kernel") == want else {})
            traffic = rec.get("dram_bytes_per_launch") if E == w["E"] else None
    except OSError:
        pass
    fp32_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    # algorithmic FLOPs of what the timed kernel computes: the whole encoder (296.0 kFLOP per agent) for the
    # fused kernel, the conv stack alone (296.0 k minus the 2 x 400 x 64 FLOP of the Linear) for the tensor-core kernel
    dom_flops = w["enc_flops"] - (2.0 * dims[6] * dims[7] if tc_encoder else 0.0)
    enc_tflops = E * N * dom_flops / (enc_ms * 1e-3) / 1e12
    env_gbs = E * N * w["env_bytes"] / (env_ms * 1e-3) / 1e9

    # ---- end to end through the public API with HOST buffers (pinned), every step ----
    # Per step: the state (positions) is uploaded from pinned host memory, the step runs, and actions, new positions
    # and done flags are read back into pinned host memory and waited for; the next step's upload uses those host
    # values (host round trip).  No host arithmetic inside the loop: done flags land in a [steps, E] pinned log.
    # The episodes are split into independent groups, each with its own stream and pinned blocks, so that the host
    # round trip of one group overlaps the kernels of the other (PipelinedHostRollout); every group still does the
    # full upload / step / download / wait protocol every step.
    ke = max(8, min(K, 64))
    G = max(1, min(args.e2e_groups, E))
    pl = m.PipelinedHostRollout(net, cfg, goals, starts, obst, groups=G, sensing_range=w["r"], device=dev)
    done_log = np.zeros((ke + 1, E), dtype=np.uint8)
    pl.start()
    for i in range(4):
        for g in range(G):
            pl.advance(g, relaunch=i < 3)
    pl.reset()
    barrier()
    t0 = time.perf_counter()
    pl.start()
    for i in range(ke):
        for g in range(G):
            _, pos_out, done_out = pl.advance(g, relaunch=i + 1 < ke)   # wait for group g's step, launch its next one
            done_log[i + 1, pl.slices[g]] = done_out.numpy()           # (returned positions are, in place, the next input)
    e2e_s = time.perf_counter() - t0
    e2e_units = float((E - done_log[:ke].astype(np.int64).sum(axis=1)).sum()) * N   # episodes alive before each step
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    u = torch.tensor([e2e_units], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    e2e_value = float(u[0]) / float(t[0])
    # the sampler ran through the device-timed steps, the per-kernel timings and the end-to-end loop
    clocks = sampler.stop() if rank == 0 else None

    train = None if args.no_train else run_train(args, w, env, dev, rank, world, barrier)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            v, tot = cpu_reference_steps(w, args.cpu_episodes, args.cpu_steps, 1)
            cpu = {"value": v, "unit": "agent-steps/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{args.cpu_episodes} episodes x {args.cpu_steps} steps of the same workload, "
                             f"{tot:.1f} s (python loop over single-episode numpy envs + batched torch-CPU forward)"}
        # encoder conv stack, encoder FC, graph filter + head (or baseline head), env step
        launches_per_step = 4 if tc_encoder else 3
        if tc_encoder:
            roof = {"bound": "tensor", "kernel": "encoder_tc_kernel", "achieved": enc_tflops,
                    "peak": pk["tensor_tflops"], "unit": "TFLOP/s", "frac": enc_tflops / pk["tensor_tflops"],
                    "traffic": traffic, "ms_per_launch": enc_ms, "share_of_step": enc_ms / (dev_ms_max / K),
                    "peak_source": f"dense 16-bit tensor peak, cuBLAS burst ({pk['source']})",
                    "algorithmic_flops_per_launch": E * N * dom_flops,
                    "vs_fp32_ffma_peak": enc_tflops / fp32_peak,
                    "note": "fp32-accurate conv stack = 3 fp16 MMAs per product (hi/lo split); the kernel is bound by "
                            "the per-layer MMA issue/commit latency, not by tensor throughput (DESIGN.md 3.2)"}
        else:
            roof = {"bound": "fp32_ffma", "kernel": "encoder_fwd_kernel", "achieved": enc_tflops,
                    "peak": fp32_peak, "unit": "TFLOP/s", "frac": enc_tflops / fp32_peak, "traffic": traffic,
                    "ms_per_launch": enc_ms, "share_of_step": enc_ms / (dev_ms_max / K),
                    "peak_source": f"148 SM x 128 FFMA x 2 x {pk['sm_max_mhz']:.0f} MHz ({pk['source']})"}
        line = {
            "metric": "agent-steps/sec (batched rollout)", "value": value, "unit": "agent-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}", "episodes_per_gpu": E, "agents": N, "board": H,
                       "obstacles": w["M"], "weights": f"tests/golden/weights_{w['model']}.npz",
                       "l2": "flushed between steps (256 MiB memset outside the event pairs)",
                       "timing": "cuda events per step on the launching stream, max over ranks; a step = one CUDA-graph "
                                 "replay of the rollout step's kernel sequence",
                       "counted": "agent-steps of episodes not yet done (done episodes are frozen); "
                                  f"nominal E*N*K = {units_nominal:.0f}",
                       "wall_s_incl_flush": wall},
            "roofline": roof,
            "roofline_env": {"bound": "hbm", "kernel": "env_step_kernel", "achieved": env_gbs, "peak": pk["hbm_gbs"],
                             "unit": "GB/s", "frac": env_gbs / pk["hbm_gbs"], "ms_per_launch": env_ms,
                             "peak_source": pk["source"]},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": pl.h2d_bytes,
                    "d2h_bytes_per_step": pl.d2h_bytes, "steps": ke, "groups": G,
                    "api": "PipelinedHostRollout: per group and step pinned upload, CUDA-graph replay, pinned download, "
                           "event wait; the groups' host round trips overlap each other's kernels"},
            "gpu_launches": K * launches_per_step,
            "clocks": clocks,
            "train": train,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_train_steps(model, states, gso, targets, steps):
    """train.py:82-100 on the host through the oracle port: train-mode forward, CE, backward, clip, Adam."""
    import torch
    from oracle import model_oracle as mo
    params = {k: torch.from_numpy(v) for k, v in load_weights(model).items()}
    keys = mo.trainable_keys(params)
    m = {k: torch.zeros_like(params[k]) for k in keys}
    v = {k: torch.zeros_like(params[k]) for k in keys}
    total = 0.0
    for it in range(steps + 1):
        t0 = time.perf_counter()
        loss, grads, bn = mo.loss_and_grads(params, states, gso, targets)
        newp, m, v, _ = mo.clip_and_adam(params, grads, m, v, step=it + 1)
        params = dict(params, **newp, **{k: t.detach() for k, t in bn.items()})
        if it > 0:
            total += time.perf_counter() - t0
    return states.shape[0] * steps / total


def run_train(args, w, env, dev, rank, world, barrier):
    """Second BASELINE metric: train samples/sec of the full imitation step (train.py:82-100) on synthetic
    CBS-shaped batches: observations recorded from the resident episodes, A + I as the loader feeds it
    (data_loader.py:74), uniform random expert actions; the message-passing network train.py trains."""
    import torch
    import mapf_gnn_b200 as m
    import mapf_gnn_b200.train as tr
    model = "baseline" if w["model"] == "baseline" else ("gnn_msg_k3" if w["N"] == 5 else w["model"])
    wl = dict(w, model=model)
    cfg = model_config(wl)
    N = w["N"]
    env.reset()
    results = []
    eye = torch.eye(N, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(3 + rank)
    for B in args.train_batches:
        net = m.build_network(cfg)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in load_weights(model).items()})
        net = net.to(dev).train()
        trainer = tr.Trainer(net, use_graph=(world == 1 and not args.no_train_graph))
        reps = (B + env.episodes - 1) // env.episodes
        states = env.fov.repeat(reps, 1, 1, 1, 1)[:B].contiguous()
        gso = (env.adj_matrix + eye).repeat(reps, 1, 1)[:B].contiguous()
        targets = torch.randint(0, 5, (B, N), device=dev, generator=gen, dtype=torch.int32)
        for _ in range(3):
            trainer.step(states, gso, targets)
        barrier()
        steps = 10
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            loss = trainer.step(states, gso, targets)
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        entry = {"batch_per_gpu": B, "global_batch": B * world, "ms_per_step": ms / steps, "cuda_graph": trainer.use_graph,
                 "samples_per_s": B * world * steps / (ms * 1e-3), "loss": float(loss.item())}
        if rank == 0 and world == 1 and not args.no_cpu_baseline and B <= 1024:
            entry["cpu_samples_per_s"] = cpu_train_steps(model, states.cpu(), gso.cpu(), targets.cpu().long(), 3)
        results.append(entry)
    return {"metric": "train samples/sec (forward + CE + backward + clip + Adam" + (" + NCCL all-reduce)" if world > 1
                                                                                    else ")"),
            "model": model, "agents": N, "results": results}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="mine", choices=["mine", "reference"])
    ap.add_argument("--workload", default="c2a", choices=sorted(WORKLOADS))
    ap.add_argument("--episodes", type=int, default=0, help="episodes per GPU (default: the workload's)")
    ap.add_argument("--ref-episodes", type=int, default=512, help="episodes per step of the --impl reference arm")
    ap.add_argument("--cpu-episodes", type=int, default=4096)
    ap.add_argument("--cpu-steps", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-groups", type=int, default=4,
                    help="independent episode groups whose host round trips overlap in the end-to-end measurement")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-train-graph", action="store_true", help="do not replay the training step as a CUDA graph")
    ap.add_argument("--train-batches", type=int, nargs="*", default=[128, 1024, 8192])
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_mine(args, w)


if __name__ == "__main__":
    main()
