"""Development: device timeline (kernel start / duration from CUPTI through torch.profiler) of one training step.
    python tools/train_timeline.py [B] [graph]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
import mapf_gnn_b200 as m
import mapf_gnn_b200.train as tr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
use_graph = len(sys.argv) > 2
w = dict(bench.WORKLOADS["c2a"], model="gnn_msg_k3")
net = m.build_network(bench.model_config(w))
net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights("gnn_msg_k3").items()})
net = net.cuda().train()
t = tr.Trainer(net, use_graph=use_graph)
g = torch.Generator(device="cuda").manual_seed(0)
states = torch.randint(0, 4, (B, 5, 2, 5, 5), device="cuda", generator=g).float()
gso = (torch.rand(B, 5, 5, device="cuda", generator=g) < 0.4).float() + torch.eye(5, device="cuda")
targets = torch.randint(0, 5, (B, 5), device="cuda", generator=g)
step = lambda: t.step(states, gso, targets)
for _ in range(6):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4):
        step()
    torch.cuda.synchronize()
import json, tempfile
tf = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(tf)
tr_ = json.load(open(tf))
ev = [e for e in tr_["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ev.sort(key=lambda e: e["ts"])
idx = [i for i, e in enumerate(ev) if "clip_adam" in e["name"]]
if len(idx) >= 2:
    ev = ev[idx[-2] + 1: idx[-1] + 1]
t0 = ev[0]["ts"]
print(f"B = {B}: {len(ev)} device activities, span {ev[-1]['ts'] + ev[-1]['dur'] - t0:.1f} us, sum of durations {sum(e['dur'] for e in ev):.1f} us")
for e in ev:
    nm = e["name"].replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:60]
    print(f"{e['ts'] - t0:8.1f} +{e['dur']:6.1f}  s{e.get('args', {}).get('stream', '?'):<3} {nm}")
if os.environ.get("TIMELINE_RAW"):
    for e in ev[5:9]:
        print(json.dumps(e)[:600])
