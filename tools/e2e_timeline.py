"""Development: device timeline (CUPTI through torch.profiler) of the bench's end-to-end loop (PipelinedHostRollout, zero copy).
    python tools/e2e_timeline.py [groups]"""
import sys, os, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import bench
import mapf_gnn_b200 as m

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
w = bench.WORKLOADS["c2a"]
E, N, H = w["E"], w["N"], w["H"]
dev = torch.device("cuda:0")
rng = np.random.default_rng(1000)
starts, goals, obst = m.make_scenarios(rng, E, N, H, w["M"])
cfg = bench.model_config(w)
net = m.build_network(cfg)
net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(w["model"]).items()})
net = net.to(dev).eval()
pl = m.PipelinedHostRollout(net, cfg, goals, starts, obst, groups=G, sensing_range=w["r"], device=dev, zero_copy=True)
pl.start()
for i in range(6):
    for g in range(G):
        pl.advance(g, relaunch=i < 5)
pl.reset()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    pl.start()
    hs = pl.groups
    for i in range(6):
        for g in range(G):
            hs[g].wait(True)
            if i < 5:
                hs[g].launch()
    torch.cuda.synchronize()
tf = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(tf)
ev = [e for e in json.load(open(tf))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
for e in ev:
    nm = e["name"].replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:40]
    print(f"{e['ts'] - t0:9.1f} +{e['dur']:7.1f}  s{e.get('args', {}).get('stream')}  {nm}  grid {e.get('args', {}).get('grid')}")
