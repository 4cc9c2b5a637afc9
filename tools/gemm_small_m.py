"""Development: duration of the training step's tensor-core GEMM shapes at small M (B = 128 -> 640 rows: five CTAs), warm,
as a CUDA graph of back-to-back launches.     python tools/gemm_small_m.py [M ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mapf_gnn_b200 import _lib

lib = _lib.load()


def timed(fn, reps=50):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s); g.replay(); g.replay(); b.record(s); b.synchronize()
    return a.elapsed_time(b) / (2 * reps) * 1e3


def main():
    Ms = [int(x) for x in sys.argv[1:]] or [640, 5120]
    for M in Ms:
        out = []
        for name, N, K in (("fc 64x400", 64, 400), ("mix 128x256", 128, 256), ("dz 256x128", 256, 128), ("da 400x64", 400, 64)):
            A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda") * 0.1
            C = torch.empty(M, N, device="cuda")
            if N <= 128:
                Bp = torch.empty((lib.mapf_tc_packed_b_size(N, K),), device="cuda")
                _lib.check(lib.mapf_tc_pack_b(B.data_ptr(), K, None, 0, N, Bp.data_ptr(), _lib.current_stream()), "pack")
                fn = lambda: lib.mapf_tc_gemm(M, N, K, A.data_ptr(), K, Bp.data_ptr(), C.data_ptr(), N, None, 1, _lib.current_stream())
            else:
                Bp = torch.empty((((N + 127) // 128) * lib.mapf_tc_packed_b_size(128, K),), device="cuda")
                _lib.check(lib.mapf_tc_pack_b_tiles(B.data_ptr(), K, N, Bp.data_ptr(), _lib.current_stream()), "pack")
                fn = lambda: lib.mapf_tc_gemm_tiles(M, N, K, A.data_ptr(), K, Bp.data_ptr(), C.data_ptr(), N, None, 0, _lib.current_stream())
            torch.cuda.synchronize()
            out.append(f"{name}: {timed(fn):.1f} us")
        print(f"M={M}: " + " | ".join(out), flush=True)


main()
