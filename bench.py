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


def config_of(name, w, episodes):
    """The `config` object of the JSON line: the same for both arms (the driver compares them)."""
    return {"workload": f"{name}: {w['desc']}", "episodes_per_gpu": episodes, "agents": w["N"], "board": w["H"],
            "obstacles": w["M"], "sensing_range": w["r"], "weights": f"tests/golden/weights_{w['model']}.npz"}


# ----------------------------------------------------------------------------------------------
# CPU arm.  Nothing below imports mapf_gnn_b200 (the reference arm must not map the product's library).
# With the unmodified reference staged under oracle/_ref (oracle/stage_ref.py) the reference's own GraphEnv and Network
# classes are timed (kind "reference"); without it the numpy / torch-CPU port in oracle/ (kind "port").
#   R2 (the stronger baseline, the one the >= 100x target is quoted against): python loop over E
#       single-episode GraphEnv.step (benchmarks/benchmark_performance.py:84-95) + ONE batched model forward per step
#       on every host thread.
#   R1: the literal evaluate.py:223-241 loop -- one GraphEnv and one B = 1 model call per episode-step.
# ----------------------------------------------------------------------------------------------
def _cpu_setup(w, episodes, seed):
    import torch
    from oracle import env_oracle as eo
    from oracle import ref_loader
    rng = np.random.default_rng(seed)
    N, H = w["N"], w["H"]
    starts, goals, obst = eo.make_scenarios(rng, episodes, N, H, w["M"])
    weights = {k: torch.from_numpy(v) for k, v in load_weights(w["model"]).items()}
    if ref_loader.available():
        cfg = dict(model_config(w), device="cpu")
        ns = ref_loader.load()
        net = ref_loader.build_network(cfg)
        net.load_state_dict(weights)
        net.eval()
        envs = [ns.GraphEnv(cfg, goal=goals[e].astype(np.int64), starting_positions=starts[e].astype(np.int64).copy(),
                            obstacles=obst[e].astype(np.int64), sensing_range=w["r"]) for e in range(episodes)]
        obs = [env.reset() for env in envs]
        emb = envs[0].getEmbedding()
        is_gnn = cfg["net_type"] == "gnn"

        def forward(states, gso):
            return net(states, gso) if is_gnn else net(states)

        def step(e, act):
            return envs[e].step(act, emb)[0]
        return "reference", obs, forward, step
    from oracle import model_oracle as mo
    envs = [eo.EnvOracle(N, H, goals[e], starts[e], obst[e], sensing_range=w["r"]) for e in range(episodes)]
    obs = [env.observe() for env in envs]
    return "port", obs, (lambda states, gso: mo.forward(weights, states, gso)), (lambda e, act: envs[e].step(act)[0])


def cpu_rollout_r2(w, episodes, steps, warmup, seed=0):
    import torch
    kind, obs, forward, step = _cpu_setup(w, episodes, seed)
    total = 0.0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        states = torch.tensor(np.stack([o["fov"] for o in obs])).float()
        gso = torch.tensor(np.stack([o["adj_matrix"] for o in obs])).float()
        with torch.no_grad():
            logits = forward(states, gso)
        acts = np.argmax(logits.numpy(), axis=2)
        obs = [step(e, acts[e]) for e in range(episodes)]
        dt = time.perf_counter() - t0
        if it >= warmup:
            total += dt
    return {"value": episodes * w["N"] * steps / total, "seconds": total, "kind": kind}


def cpu_rollout_r1(w, episodes, steps, seed=0):
    import torch
    kind, obs, forward, step = _cpu_setup(w, episodes, seed)
    t0 = time.perf_counter()
    for e in range(episodes):
        o = obs[e]
        for _ in range(steps):
            fov = torch.tensor(o["fov"]).float().unsqueeze(0)
            gso = torch.tensor(o["adj_matrix"]).float().unsqueeze(0)
            with torch.no_grad():
                action = forward(fov, gso).squeeze(0).numpy()
            o = step(e, np.argmax(action, axis=1))
    total = time.perf_counter() - t0
    return {"value": episodes * w["N"] * steps / total, "seconds": total, "kind": kind}


_PROTOCOL = {"reference": "the unmodified reference classes staged under oracle/_ref (GraphEnv.step per episode in a python "
                          "loop + one batched Network.forward per step; SURVEY 8d protocol R2)",
             "port": "oracle/ port of the reference (python loop over single-episode numpy envs + batched torch-CPU "
                     "forward; oracle/_ref not staged on this box)"}


def run_reference(args, w):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)   # torchrun pins OMP_NUM_THREADS=1; the CPU arm may use every core
    episodes = args.ref_episodes
    r2 = cpu_rollout_r2(w, episodes, args.steps, args.warmup)
    r1 = cpu_rollout_r1(w, max(1, min(32, episodes)), min(args.steps, 8))
    cores = torch.get_num_threads()
    value = r2["value"]
    line = {
        "impl": "reference", "metric": "agent-steps/sec (batched rollout)", "value": value, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r2["seconds"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args.workload, w, args.episodes or w["E"]),
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": r2["kind"],
                         "sample": f"{episodes} episodes x {args.steps} steps of the workload per timed run "
                                   f"(throughput is per agent-step, independent of the episode count): "
                                   + _PROTOCOL[r2["kind"]],
                         "r1_literal_evaluate_loop": {"value": r1["value"], "unit": "agent-steps/s",
                                                      "sample": "B = 1 model call per episode-step, evaluate.py:223-241",
                                                      "seconds": r1["seconds"]}},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def _time_alone(fn, flush, reps=10, warm=3):
    """Mean CUDA-event time (ms) of one launch sequence issued alone on the current stream, L2 flushed before each."""
    import torch
    ms = []
    for i in range(warm + reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        if i >= warm:
            ms.append(a.elapsed_time(b))
    return float(np.mean(ms))


def _traffic(name, kernel, full_size):
    """ncu-measured DRAM bytes per launch of `kernel` on workload `name` (committed summary; None when not captured)."""
    if not full_size:
        return None
    for fn in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as fh:
                rec = json.load(fh).get(name, {})
        except OSError:
            continue
        rec = rec.get(kernel, rec if rec.get("kernel") == kernel else {})
        if rec.get("dram_bytes_per_launch") is not None:
            return rec["dram_bytes_per_launch"]
    return None


def measure_workload(m, name, w, args, dev, rank, world, barrier, flush, K, W):
    """Device-timed rollout steps of one workload + its kernels timed alone.  Returns the numbers and the live objects."""
    import torch
    import torch.distributed as dist
    from mapf_gnn_b200 import _lib as _mlib
    lib = _mlib.load()
    E, N, H = args.episodes or w["E"], w["N"], w["H"]
    rng = np.random.default_rng(1000 + rank)
    starts, goals, obst = m.make_scenarios(rng, E, N, H, w["M"])
    cfg = model_config(w)
    net = m.build_network(cfg)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in load_weights(w["model"]).items()})
    net = net.to(dev).eval()
    env = m.BatchedGraphEnv(cfg, goals, starts, obst, sensing_range=w["r"], device=dev)
    ro = m.Rollout(net, env)

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
    ms_step = dev_ms_max / K
    pk = peaks()
    full = E == w["E"]
    rows = E * N
    dims = net.dims()
    packed = net.packed_weights()
    st = _mlib.current_stream()

    # ---- dominant kernel timed alone, live, on the launching stream: the conv stack of the encoder, on the tensor cores
    # (encoder_tc_kernel) where the dimensions fit, else inside the fused FFMA2 encoder kernel (encoder_fwd_kernel) ----
    tc_encoder = ro.worka is not None and lib.mapf_encoder_tc_supported(ro.params) == 1
    if tc_encoder:
        ea = _mlib.EncoderFwdArgs(rows, env.fov.data_ptr(), packed.data_ptr(), ro.worka.data_ptr())
        enc_ms = _time_alone(lambda: _mlib.check(lib.mapf_encoder_tc_conv_fwd(ro.params, ea, st), "enc_tc"), flush)
    else:
        enc_ms = _time_alone(lambda: torch.ops.mapf.encoder_fwd(dims, env.fov, packed), flush)
    fp32_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    # algorithmic FLOPs of what the timed kernel computes: the whole encoder (296.0 kFLOP per agent) for the
    # fused kernels, the conv stack alone (296.0 k minus the 2 x 400 x 64 FLOP of the Linear) when the FC is a second kernel
    fc_separate = tc_encoder and not ENC_TC_FUSED_FC
    dom_flops = w["enc_flops"] - (2.0 * dims[6] * dims[7] if fc_separate else 0.0)
    enc_tflops = rows * dom_flops / (enc_ms * 1e-3) / 1e12
    if tc_encoder:
        roof = {"bound": "tensor", "kernel": "encoder_tc_kernel", "achieved": enc_tflops, "peak": pk["tensor_tflops"],
                "unit": "TFLOP/s", "frac": enc_tflops / pk["tensor_tflops"],
                "traffic": _traffic(name, "encoder_tc_kernel", full), "ms_per_launch": enc_ms,
                "share_of_step": enc_ms / ms_step,
                "peak_source": f"dense 16-bit tensor peak, cuBLAS burst ({pk['source']})",
                "algorithmic_flops_per_launch": rows * dom_flops, "vs_fp32_ffma_peak": enc_tflops / fp32_peak,
                "note": "fp32-accurate conv stack = 3 fp16 MMAs per product (hi/lo split) on 25 of 32 lanes; the kernel "
                        "is bound by the per-layer MMA issue/commit latency, not by tensor throughput (DESIGN.md 3.2)"}
    else:
        roof = {"bound": "fp32_ffma", "kernel": "encoder_fwd_kernel", "achieved": enc_tflops, "peak": fp32_peak,
                "unit": "TFLOP/s", "frac": enc_tflops / fp32_peak, "traffic": _traffic(name, "encoder_fwd_kernel", full),
                "ms_per_launch": enc_ms, "share_of_step": enc_ms / ms_step,
                "peak_source": f"148 SM x 128 FFMA x 2 x {pk['sm_max_mhz']:.0f} MHz ({pk['source']})"}

    # ---- environment step kernel alone (FOV + adjacency: HBM-bound; SURVEY 8d bytes per agent-step) ----
    env_ms = _time_alone(lambda: env.observe(), flush)
    env_gbs = rows * w["env_bytes"] / (env_ms * 1e-3) / 1e9
    roof_env = {"bound": "hbm", "kernel": "env_step_kernel", "achieved": env_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": env_gbs / pk["hbm_gbs"], "ms_per_launch": env_ms, "share_of_step": env_ms / ms_step,
                "traffic": _traffic(name, "env_step_kernel", full),
                "algorithmic_bytes_per_launch": rows * w["env_bytes"], "peak_source": pk["source"]}

    # ---- graph-filter stage alone (S, hops, mix, ReLU, head, argmax) ----
    roof_graph = None
    if dims[8] > 0:
        F, K_, G = dims[7], dims[8], dims[9]
        skip = 1 if net.message else 0
        KT = (K_ + skip) * F
        gf = net.GFL[0]
        Wt = gf.W.detach().contiguous()
        W2 = gf.W2.detach().contiguous() if net.message else None
        Wp = torch.empty((lib.mapf_tc_packed_b_size(G, KT),), device=dev)
        _mlib.check(lib.mapf_tc_pack_b(Wt.data_ptr(), K_ * F, None if W2 is None else W2.data_ptr(), F if skip else 0, G,
                                       Wp.data_ptr(), st), "tc_pack_b")
        aw = net.actionMLP[0].weight.detach().contiguous()
        ab = net.actionMLP[0].bias.detach().contiguous()
        lg = torch.empty((rows, 5), device=dev)
        ac = torch.empty((rows,), dtype=torch.int32, device=dev)
        x = ro.work
        # SURVEY 8d per agent: shift 2(K-1)FN + mix 2KFG (+ skip 2FG) + head 2*G*5 FLOP; fused bytes 4F + 4N in, 24 out
        g_flops = 2.0 * (K_ - 1) * F * N + 2.0 * KT * G + 2.0 * G * 5
        fused = N <= 32 and lib.mapf_tc_graph_supported(N, F, G, K_) == 1
        if fused:
            fn = lambda: _mlib.check(lib.mapf_tc_graph_head(E, N, F, G, K_, 0 if skip else 1, skip, x.data_ptr(),
                                                           env.adj_matrix.data_ptr(), Wp.data_ptr(), aw.data_ptr(),
                                                           ab.data_ptr(), 5, lg.data_ptr(), ac.data_ptr(), st), "tc_graph")
            g_ms = _time_alone(fn, flush)
            g_bytes = rows * (4 * F + 4 * N + 24)
            tfl = rows * g_flops / (g_ms * 1e-3) / 1e12
            roof_graph = {"bound": "tensor", "kernel": "tc_graph_kernel", "achieved": tfl, "peak": pk["tensor_tflops"],
                          "unit": "TFLOP/s", "frac": tfl / pk["tensor_tflops"], "ms_per_launch": g_ms,
                          "share_of_step": g_ms / ms_step, "algorithmic_flops_per_launch": rows * g_flops,
                          "hbm_gbs": g_bytes / (g_ms * 1e-3) / 1e9, "hbm_frac": g_bytes / (g_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                          "algorithmic_bytes_per_launch": g_bytes, "traffic": _traffic(name, "tc_graph_kernel", full),
                          "note": "fused S + hops + 3xTF32 mix + head: the taps never touch HBM, so the stage is bound by "
                                  "its per-CTA latency chain, not by HBM (SURVEY 8d: fused it turns compute-bound)"}
        else:
            z = ro.workz
            ga = _mlib.GraphFilterFwdArgs()
            ga.batch, ga.agents, ga.in_dim, ga.out_dim, ga.taps = E, N, F, G, K_
            ga.add_self_loop, ga.has_skip = 0 if skip else 1, skip
            ga.x, ga.x_sb, ga.x_sn, ga.x_sf = x.data_ptr(), N * F, F, 1
            ga.gso, ga.z = env.adj_matrix.data_ptr(), z.data_ptr()
            taps_ms = _time_alone(lambda: _mlib.check(lib.mapf_graph_taps(ga, st), "graph_taps"), flush)
            mix_ms = _time_alone(lambda: _mlib.check(lib.mapf_tc_mix_head(rows, G, KT, z.data_ptr(), KT, Wp.data_ptr(),
                                                                         aw.data_ptr(), ab.data_ptr(), 5, lg.data_ptr(),
                                                                         ac.data_ptr(), None, st), "tc_mix_head"), flush)
            t_bytes = rows * (4 * F + 4 * N + 4 * KT)        # SURVEY 8d recurrence kernel: read X, A; write Z
            gbs = t_bytes / (taps_ms * 1e-3) / 1e9
            hops_tc = N == 64 and F == 64 and not skip
            roof_graph = {"bound": "hbm", "kernel": "graph_hops_tc_kernel (hops on tcgen05, M = 64 per episode)" if hops_tc
                          else "graph_taps_kernel", "achieved": gbs, "peak": pk["hbm_gbs"],
                          "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "ms_per_launch": taps_ms,
                          "share_of_step": taps_ms / ms_step, "algorithmic_bytes_per_launch": t_bytes,
                          "traffic": _traffic(name, "graph_taps_kernel", full),
                          "mix_head": {"kernel": "tc_gemm_kernel<EPI_HEAD>", "ms_per_launch": mix_ms,
                                       "tflops": rows * (2.0 * KT * G + 2.0 * G * 5) / (mix_ms * 1e-3) / 1e12}}
    launches = (2 if fc_separate else 1) + (0 if dims[8] == 0 else (1 if (N <= 32) else 2)) + (1 if dims[8] == 0 else 0) + 1
    res = {"value": value, "ms_per_step": ms_step, "units_nominal": units_nominal, "wall": wall,
           "roofline": roof, "roofline_env": roof_env, "roofline_graph": roof_graph, "launches_per_step": launches}
    live = {"net": net, "env": env, "ro": ro, "cfg": cfg, "goals": goals, "starts": starts, "obst": obst, "E": E}
    return res, live


ENC_TC_FUSED_FC = False   # set by run_mine from the library (True once the conv kernel carries the FC in its tail)


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
    from mapf_gnn_b200 import _lib as _mlib
    global ENC_TC_FUSED_FC
    ENC_TC_FUSED_FC = bool(getattr(_mlib, "ENC_TC_FUSED_FC", False))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, args.warmup
    N = w["N"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res, live = measure_workload(m, args.workload, w, args, dev, rank, world, barrier, flush, K, W)
    net, env, cfg, E = live["net"], live["env"], live["cfg"], live["E"]
    goals, starts, obst = live["goals"], live["starts"], live["obst"]

    # ---- end to end through the public API with HOST buffers (pinned), every step ----
    # Per step: the state (positions) is uploaded from pinned host memory, the step runs, and actions, new positions
    # and done flags are read back into pinned host memory and waited for; the next step's upload uses those host
    # values (host round trip).  No host arithmetic inside the loop: done flags land in a [steps, E] pinned log.
    # The episodes are split into independent groups, each with its own stream and pinned blocks, so that the host
    # round trip of one group overlaps the kernels of the other (PipelinedHostRollout); every group still does the
    # full upload / step / download / wait protocol every step.
    ke = max(8, min(K, 64))
    G = max(1, min(args.e2e_groups, E))
    pl = m.PipelinedHostRollout(net, cfg, goals, starts, obst, groups=G, sensing_range=w["r"], device=dev,
                                zero_copy=not args.e2e_copies)
    done_log = np.zeros((ke + 1, E), dtype=np.uint8)
    pl.start()
    for i in range(4):
        for g in range(G):
            pl.advance(g, relaunch=i < 3)
    pl.reset()
    barrier()
    t0 = time.perf_counter()
    pl.start()
    hs, sls = pl.groups, pl.slices
    for i in range(ke):
        row, again = done_log[i + 1], i + 1 < ke
        for g in range(G):
            h = hs[g]
            _, pos_out, done_out = h.wait(True)      # group g's step is in pinned host memory (numpy views of it)
            if again:
                h.launch()                           # its next step: the returned positions are, in place, the next input
            row[sls[g]] = done_out
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
    h2d, d2h = pl.h2d_bytes, pl.d2h_bytes
    del pl

    train = None if args.no_train else run_train(args, w, env, dev, rank, world, barrier)
    del live, env, net

    # ---- the other BASELINE configs, compact: same protocol (flushed L2, one graph replay per step), fewer steps ----
    others = {}
    if not args.no_other_workloads and args.workload == "c2a" and not args.episodes:
        for name in args.other_workloads:
            ow = WORKLOADS[name]
            r, lv = measure_workload(m, name, ow, args, dev, rank, world, barrier, flush, args.other_steps, max(3, W))
            del lv
            torch.cuda.empty_cache()
            others[name] = {"config": config_of(name, ow, ow["E"]), "value": r["value"], "unit": "agent-steps/s",
                            "ms_per_step": r["ms_per_step"], "steps": args.other_steps, "n_gpus": world,
                            "roofline": {k: r["roofline"][k] for k in ("bound", "kernel", "achieved", "peak", "unit", "frac",
                                                                        "ms_per_launch", "share_of_step")},
                            "roofline_env": {k: r["roofline_env"][k] for k in ("achieved", "peak", "unit", "frac",
                                                                                "ms_per_launch")},
                            "roofline_graph": None if r["roofline_graph"] is None else
                            {k: v for k, v in r["roofline_graph"].items() if k != "note"}}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            r2 = cpu_rollout_r2(w, args.cpu_episodes, args.cpu_steps, 1)
            cpu = {"value": r2["value"], "unit": "agent-steps/s", "cores": torch.get_num_threads(), "kind": r2["kind"],
                   "sample": f"{args.cpu_episodes} episodes x {args.cpu_steps} steps of the same workload, "
                             f"{r2['seconds']:.1f} s: " + _PROTOCOL[r2["kind"]]}
        line = {
            "metric": "agent-steps/sec (batched rollout)", "value": res["value"], "unit": "agent-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args.workload, w, E),
            "measurement": {"l2": "flushed between steps (256 MiB memset outside the event pairs)",
                            "timing": "cuda events per step on the launching stream, max over ranks; a step = one "
                                      "CUDA-graph replay of the rollout step's kernel sequence",
                            "counted": "agent-steps of episodes not yet done (done episodes are frozen); "
                                       f"nominal E*N*K = {res['units_nominal']:.0f}",
                            "wall_s_incl_flush": res["wall"]},
            "roofline": res["roofline"],
            "roofline_env": res["roofline_env"],
            "roofline_graph": res["roofline_graph"],
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": ke, "groups": G,
                    "api": ("PipelinedHostRollout: the state a host caller holds (agent POSITIONS and done flags only: boards, "
                            "goals and weights are resident, observations are produced on the device) lives in pinned host "
                            "memory; per group and step " +
                            ("the positions are uploaded (cudaMemcpyAsync), the step's CUDA graph replays, actions + new "
                             "positions + done flags are downloaded" if args.e2e_copies else
                             "the step's CUDA graph replays and its kernels READ the positions / done flags from and WRITE "
                             "actions + new positions + done flags to the pinned blocks themselves (mapped host memory, "
                             "no copy nodes; h2d / d2h bytes = what those loads / stores move over the bus)") +
                            ", then the host waits for the step's event and reads the results; the groups' host round "
                            "trips overlap each other's kernels.  This is not the reference caller protocol of shipping "
                            "fov / gso tensors per step (evaluate.py:225-226)")},
            "gpu_launches": K * res["launches_per_step"],
            "clocks": clocks,
            "train": train,
            "workloads": others,
            "expert": run_expert(args) if world == 1 and not args.no_cpu_baseline else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_train_steps(model, states, gso, targets, steps):
    """train.py:82-100 on the host: train-mode forward, CrossEntropyLoss, backward, clip_grad_norm_(1.0), Adam(3e-4,
    weight_decay 1e-4) -- with the unmodified reference Network when it is staged (oracle/_ref), else the oracle port."""
    import torch
    from oracle import ref_loader
    weights = {k: torch.from_numpy(v) for k, v in load_weights(model).items()}
    if ref_loader.available():
        cfg = dict(model_config(dict(model=model, N=int(states.shape[1]), H=18)), device="cpu")
        net = ref_loader.build_network(cfg)
        net.load_state_dict(weights)
        net.train()
        opt = torch.optim.Adam(net.parameters(), lr=3e-4, weight_decay=1e-4)
        crit = torch.nn.CrossEntropyLoss()
        is_gnn = cfg["net_type"] == "gnn"
        total = 0.0
        for it in range(steps + 1):
            t0 = time.perf_counter()
            opt.zero_grad()
            out = net(states, gso) if is_gnn else net(states)
            loss = crit(out.reshape(-1, 5), targets.long().reshape(-1))
            loss.backward()
            torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)
            opt.step()
            if it > 0:
                total += time.perf_counter() - t0
        return states.shape[0] * steps / total, "reference"
    from oracle import model_oracle as mo
    params = weights
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
    return states.shape[0] * steps / total, "port"


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
        trainer = tr.Trainer(net, use_graph=False, peer_allreduce=False if args.nccl_allreduce else None)
        # the step replays as ONE CUDA graph; with several ranks that needs the peer-memory all-reduce (no NCCL call inside)
        trainer.use_graph = (not args.no_train_graph) and (world == 1 or trainer.peer is not None)
        reps = (B + env.episodes - 1) // env.episodes
        states = env.fov.repeat(reps, 1, 1, 1, 1)[:B].contiguous()
        gso = (env.adj_matrix + eye).repeat(reps, 1, 1)[:B].contiguous()
        targets = torch.randint(0, 5, (B, N), device=dev, generator=gen, dtype=torch.int32)
        for _ in range(3):
            trainer.step(states, gso, targets)
        if trainer.use_graph:
            # the batch lives in the graph's static input tensors (where a loader's host-to-device copies would land):
            # the timed step is the graph launch alone, no device-to-device staging copies in front of it
            states, gso, targets = trainer.input_buffers(states, gso, targets)
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
                 "inputs": "resident in the trainer's static input tensors" if trainer.use_graph else "resident, passed per step",
                 "allreduce": None if world == 1 else ("peer-memory one-shot, fused with the gradient norm"
                                                       if trainer.peer is not None else "nccl"),
                 "samples_per_s": B * world * steps / (ms * 1e-3), "loss": float(loss.item())}
        # ~3x forward FLOPs per agent -> 5.46 MFLOP (msg_k3) / 5.22 MFLOP (gcn_k3) per N = 5 sample;
        # other shapes: 3 x the workload's forward figure per agent
        per_sample = {"gnn_msg_k3": 5.46e6, "gnn_k3": 5.22e6}.get(model, 3.0 * w["flops"] * N) if N == 5 else 3.0 * w["flops"] * N
        pk = peaks()
        tfl = B * per_sample / (ms / steps * 1e-3) / 1e12
        fp32_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
        entry["roofline"] = {"bound": "mixed: tensor-core contractions (3 MMAs per product, fp32 accuracy), FFMA2 conv weight "
                                      "gradients, HBM-bound BatchNorm kernels", "achieved": tfl, "unit": "TFLOP/s",
                             "algorithmic_flops_per_step": B * per_sample,
                             "frac_of_fp32_ffma_peak": tfl / fp32_peak, "fp32_ffma_peak": fp32_peak,
                             "frac_of_tensor_peak": tfl / pk["tensor_tflops"], "tensor_peak": pk["tensor_tflops"]}
        if rank == 0 and world == 1 and not args.no_cpu_baseline and B <= 1024:
            entry["cpu_samples_per_s"], entry["cpu_kind"] = cpu_train_steps(model, states.cpu(), gso.cpu(),
                                                                            targets.cpu().long(), 3)
        results.append(entry)
        trainer.close()
    return {"metric": "train samples/sec (forward + CE + backward + clip + Adam" + (" + gradient all-reduce)" if world > 1
                                                                                    else ")"),
            "model": model, "agents": N, "results": results}


def run_expert(args):
    """SURVEY 8f row 4, compact: CBS plans per second of the native expert (one host thread, and the batch call on all
    host threads) on main_data.py's instance shape (28x28, 5 agents, 8 obstacles; np.random.seed(0) through gen_input),
    beside the unmodified reference's `CBS(Environment(...)).search()` on a bounded sample of the same instances.  Host
    code on both sides: the search is branchy and sequential; no GPU kernel is involved."""
    import mapf_gnn_b200.cbs as cbs
    np.random.seed(0)
    inputs = [cbs.gen_input((28, 28), 8, 5) for _ in range(256)]
    insts = [(i["map"]["dimensions"], [a["start"] for a in i["agents"]], [a["goal"] for a in i["agents"]],
              i["map"]["obstacles"]) for i in inputs]
    cbs.solve_batch(insts[:8], threads=1)
    t0 = time.perf_counter()
    one = cbs.solve_batch(insts, threads=1)
    t1 = time.perf_counter()
    many = cbs.solve_batch(insts, threads=0)
    t2 = time.perf_counter()
    assert all((a is None) == (b is None) and (a is None or a[1] == b[1]) for a, b in zip(one, many))
    per = []
    for it in insts[:64]:                      # per-instance times: the totals above are dominated by the few capped searches
        ta = time.perf_counter()
        cbs.solve_batch([it], threads=1)
        per.append(time.perf_counter() - ta)
    out = {"metric": "CBS instances/s (28x28, 5 agents, 8 obstacles)", "instances": len(insts),
           "solved": sum(p is not None for p in many), "native_1_thread": len(insts) / (t1 - t0),
           "native_all_threads": len(insts) / (t2 - t1), "host_threads": os.cpu_count(),
           "native_median_ms_per_instance": float(np.median(per)) * 1e3,
           "note": "instances that hit the 5000-node cap (one here) take ~1000x the median and bound the batch call"}
    try:
        from oracle import ref_loader
        root = ref_loader.root()
        if root is not None and os.path.exists(os.path.join(root, "cbs", "cbs.py")):
            if root not in sys.path:
                sys.path.insert(0, root)
            import importlib
            ref = importlib.import_module("cbs.cbs")
            sample = 16
            t0 = time.perf_counter()
            same = True
            ref_per = []
            for inp, mine in zip(inputs[:sample], many[:sample]):
                ta = time.perf_counter()
                plan = ref.CBS(ref.Environment(inp["map"]["dimensions"], inp["agents"], inp["map"]["obstacles"]), verbose=False).search()
                ref_per.append(time.perf_counter() - ta)
                got = None if not plan else [[(s["x"], s["y"]) for s in plan[a["name"]]] for a in inp["agents"]]
                same = same and (got is None) == (mine is None) and (got is None or got == [[tuple(c) for c in p] for p in mine[0]])
            out["reference_1_thread"] = sample / (time.perf_counter() - t0)
            out["reference_median_ms_per_instance"] = float(np.median(ref_per)) * 1e3
            out["reference_sample"] = f"first {sample} of the same instances, unmodified cbs/cbs.py staged under oracle/_ref"
            out["plans_identical_on_sample"] = bool(same)
    except Exception as exc:            # the expert line is an extra: never fail the bench on it
        out["reference_error"] = repr(exc)[:200]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="mine", choices=["mine", "reference"])
    ap.add_argument("--workload", default="c2a", choices=sorted(WORKLOADS))
    ap.add_argument("--episodes", type=int, default=0, help="episodes per GPU (default: the workload's)")
    ap.add_argument("--ref-episodes", type=int, default=512, help="episodes per step of the --impl reference arm")
    ap.add_argument("--cpu-episodes", type=int, default=1024)
    ap.add_argument("--cpu-steps", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-groups", type=int, default=2,
                    help="independent episode groups whose host round trips overlap in the end-to-end measurement")
    ap.add_argument("--e2e-copies", action="store_true",
                    help="end-to-end loop: cudaMemcpyAsync nodes around the kernels instead of the kernels reading / "
                         "writing the pinned host blocks themselves")
    ap.add_argument("--no-other-workloads", action="store_true",
                    help="skip the compact lines of the other BASELINE configs (key `workloads`)")
    ap.add_argument("--other-workloads", nargs="*", default=["c1", "c2b", "c3", "c5"], choices=sorted(WORKLOADS))
    ap.add_argument("--other-steps", type=int, default=32)
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-train-graph", action="store_true", help="do not replay the training step as a CUDA graph")
    ap.add_argument("--nccl-allreduce", action="store_true",
                    help="training: NCCL all-reduce of the flat gradient instead of the fused peer-memory kernel")
    ap.add_argument("--train-batches", type=int, nargs="*", default=[128, 1024, 8192])
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_mine(args, w)


if __name__ == "__main__":
    main()
