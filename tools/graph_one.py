"""Graph-filter stage alone at a workload's shape (for ncu).  python tools/graph_one.py [c2a|c3|c5]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import mapf_gnn_b200 as m
name = sys.argv[1] if len(sys.argv) > 1 else "c2a"
w = bench.WORKLOADS[name]
E, N = w["E"], w["N"]
rng = np.random.default_rng(0)
starts, goals, obst = m.make_scenarios(rng, E, N, w["H"], w["M"])
cfg = bench.model_config(w)
net = m.build_network(cfg)
net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(w["model"]).items()})
net = net.cuda().eval()
env = m.BatchedGraphEnv(cfg, goals, starts, obst, sensing_range=w["r"])
ro = m.Rollout(net, env)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(6):
    flush.zero_()
    ro.run(1)
torch.cuda.synchronize()
print("ok")
