"""Development: device timeline (CUPTI through torch.profiler) of the bench's timed rollout step (one CUDA-graph replay per step,
L2 flushed in front).    python tools/rollout_timeline.py [workload]"""
import sys, os, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import bench
import mapf_gnn_b200 as m

name = sys.argv[1] if len(sys.argv) > 1 else "c2a"
w = bench.WORKLOADS[name]
E, N, H = w["E"], w["N"], w["H"]
dev = torch.device("cuda:0")
rng = np.random.default_rng(1000)
starts, goals, obst = m.make_scenarios(rng, E, N, H, w["M"])
cfg = bench.model_config(w)
net = m.build_network(cfg)
net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(w["model"]).items()})
net = net.to(dev).eval()
env = m.BatchedGraphEnv(cfg, goals, starts, obst, sensing_range=w["r"], device=dev)
ro = m.Rollout(net, env)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(4):
    ro.run(1)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    ro.run(1)
env.reset()
for _ in range(3):
    flush.zero_(); g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4):
        flush.zero_()
        g.replay()
    torch.cuda.synchronize()
tf = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(tf)
ev = [e for e in json.load(open(tf))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
for e in ev:
    nm = e["name"].replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:60]
    print(f"{e['ts'] - t0:9.1f} +{e['dur']:7.1f}  {nm}  grid {e.get('args', {}).get('grid')}")
