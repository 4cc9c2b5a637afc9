"""Timing sweep of the tcgen05 GEMM (development)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mapf_gnn_b200 import _lib
lib = _lib.load()
st = _lib.current_stream()


def run(M, N, K, iters=30):
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    Bp = torch.empty((lib.mapf_tc_packed_b_size(N, K),), device="cuda")
    lib.mapf_tc_pack_b(B.data_ptr(), K, None, 0, N, Bp.data_ptr(), st)
    f = lambda: lib.mapf_tc_gemm(M, N, K, A.data_ptr(), K, Bp.data_ptr(), out.data_ptr(), N, bias.data_ptr(), 1, st)
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        f()
    b.record()
    b.synchronize()
    us = a.elapsed_time(b) / iters * 1e3
    print(f"M={M:7d} N={N:4d} K={K:4d}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:8.2f} TFLOP/s(fp32-equivalent)")


for M in (128, 128 * 148, 20480, 128 * 296, 524288):
    for (N, K) in ((64, 400), (128, 128), (128, 256), (64, 8), (64, 48)):
        run(M, N, K)
