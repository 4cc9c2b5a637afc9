"""GPU cost of one host-stepped group step (upload node + kernels + download node), replayed back to back without
host waits, against the group size: what the pipelined end-to-end loop can reach at best with G groups."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import mapf_gnn_b200 as m
from mapf_gnn_b200 import _lib
from mapf_gnn_b200.rollout import Rollout, HostSteppedRollout
from mapf_gnn_b200.env import BatchedGraphEnv

w = bench.WORKLOADS["c2a"]
cfg = bench.model_config(w)
net = m.build_network(cfg)
net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(w["model"]).items()})
net = net.cuda().eval()
lib = _lib.load()
for E in [int(a) for a in sys.argv[1:]] or [544, 1184, 1776, 2368, 4096]:
    starts, goals, obst = m.make_scenarios(np.random.default_rng(1000), E, w["N"], w["H"], w["M"])
    env = BatchedGraphEnv(cfg, goals, starts, obst, sensing_range=w["r"], device="cuda")
    h = HostSteppedRollout(Rollout(net, env))
    torch.cuda.synchronize()
    for rep in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(h.stream):
            a.record()
            for i in range(40):
                lib.mapf_graph_launch(h._exec[i & 1], h._stream, h._event)
            b.record()
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(40):
        lib.mapf_graph_launch(h._exec[i & 1], h._stream, h._event)
        lib.mapf_event_wait(h._event)
    rt = (time.perf_counter() - t0) / 40 * 1e6
    print(f"E={E}: {a.elapsed_time(b) / 40 * 1e3:.1f} us per step back to back, {rt:.1f} us per launch+wait round trip", flush=True)
