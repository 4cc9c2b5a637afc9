import sys, os
sys.path.insert(0, "/root/repo")
import torch, ctypes as C
from mapf_gnn_b200 import _lib
lib = _lib.load()
st = _lib.current_stream()
M, N = 20480, 64
def t(K, iters=50):
    Ap = torch.zeros((lib.mapf_tc_presplit_a_size(M, K),), device="cuda")
    Bp = torch.zeros((lib.mapf_tc_packed_b_size(N, K),), device="cuda")
    x = torch.empty((M, N), device="cuda"); b = torch.zeros(N, device="cuda")
    f = lambda: lib.mapf_tc_fc_presplit(M, N, K, Ap.data_ptr(), Bp.data_ptr(), x.data_ptr(), N, b.data_ptr(), 1, st)
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): f()
    e.record(); e.synchronize()
    return a.elapsed_time(e) / iters * 1e3
for K in (8, 32, 104, 200, 400, 800):
    print(K, round(t(K), 2), "us")
