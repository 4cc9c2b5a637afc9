"""Filter-tap kernels at the C5 shape: tensor-core hops vs FFMA taps (CUDA events, L2 flushed / warm)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mapf_gnn_b200 import _lib
lib = _lib.load()
B, N, F, K = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 64, 64, 4
x = torch.randn(B * N, F, device="cuda")
adj = (torch.rand(B, N, N, device="cuda") < 0.05).float()
z = torch.zeros(B * N, K * F, device="cuda")
a = _lib.GraphFilterFwdArgs()
a.batch, a.agents, a.in_dim, a.out_dim, a.taps = B, N, F, 128, K
a.add_self_loop, a.has_skip = 1, 0
a.x, a.x_sb, a.x_sn, a.x_sf = x.data_ptr(), N * F, F, 1
a.gso, a.z = adj.data_ptr(), z.data_ptr()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = _lib.current_stream()
outs = {}
for name, fn in (("tensor-core hops", lib.mapf_graph_taps), ("ffma taps", lib.mapf_graph_taps_ffma)):
    ts = []
    for i in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rc = fn(a, st); e1.record(); e1.synchronize()
        assert rc == 0, rc
        if i >= 4: ts.append(e0.elapsed_time(e1))
    torch.cuda.synchronize()
    outs[name] = z.clone()
    print(f"{name}: {np.median(ts) * 1e3:.1f} us flushed", flush=True)
d = (outs["tensor-core hops"].double() - outs["ffma taps"].double()).abs().max().item() / outs["ffma taps"].abs().max().item()
print(f"max rel diff tensor-core vs ffma: {d:.2e}", flush=True)
