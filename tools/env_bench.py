"""Environment step kernel alone: time per launch and achieved HBM bandwidth for every work decomposition.
    python tools/env_bench.py [c2a c5 ...]"""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import mapf_gnn_b200 as m

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name in (sys.argv[1:] or ["c2a", "c1", "c3", "c5"]):
    w = bench.WORKLOADS[name]
    E, N, H = w["E"], w["N"], w["H"]
    rng = np.random.default_rng(0)
    starts, goals, obst = m.make_scenarios(rng, E, N, H, w["M"])
    env = m.BatchedGraphEnv(bench.model_config(w), goals, starts, obst, sensing_range=w["r"])
    acts = torch.randint(0, 5, (E, N), device="cuda", dtype=torch.int32)
    lanes = 8
    while lanes < N:
        lanes *= 2
    for epb in [0] + [e for e in (1, 2, 4, 8, 16) if e <= 128 // lanes]:
        env.episodes_per_cta = epb
        res = {}
        for mode, fl in (("flushed", True), ("warm", False)):
            ts = []
            for i in range(13):
                if fl:
                    flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                env.step(acts, freeze_done=False)
                b.record()
                b.synchronize()
                if i >= 3:
                    ts.append(a.elapsed_time(b))
            res[mode] = float(np.median(ts)) * 1e3
        gb = E * N * w["env_bytes"] / 1e9
        print(f"{name} epb={epb}: flushed {res['flushed']:.1f} us ({gb / res['flushed'] * 1e6:.0f} GB/s)  "
              f"warm {res['warm']:.1f} us ({gb / res['warm'] * 1e6:.0f} GB/s)", flush=True)
