"""Timing of the tensor-core conv stack over start-up skews (development)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import load_network
from mapf_gnn_b200 import _lib
from mapf_gnn_b200.ops import net_params
lib = _lib.load()
net = load_network("gnn_k3", num_agents=1).eval()
dims, packed = net.dims(), net.packed_weights()
p = net_params(dims)
s = _lib.current_stream()

def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / iters * 1e3

for rows in (20480, 81920):
    st = torch.randint(0, 4, (rows, 2, 5, 5), device="cuda").float()
    Ap = torch.empty((lib.mapf_tc_presplit_a_size(rows, 400),), device="cuda")
    ea = _lib.EncoderFwdArgs(rows, st.data_ptr(), packed.data_ptr(), Ap.data_ptr())
    out = []
    for skew in sys.argv[1:] or ["0", "350", "700", "1000", "1400"]:
        os.environ["MAPF_B200_ETC_SKEW"] = skew
        out.append(f"skew {skew}: {timeit(lambda: lib.mapf_encoder_tc_conv_fwd(p, ea, s)):.1f} us")
    print(f"rows {rows}: " + " | ".join(out))
