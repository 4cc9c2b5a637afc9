"""tcgen05 GEMMs of the graph mix / training step at large row counts (CUDA events, flushed): mix + head (C5 shape),
plain store GEMM (training FC shape), column-tiled GEMM (training da shape)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mapf_gnn_b200 import _lib
lib = _lib.load()
st = _lib.current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(name, fn, bytes_):
    ts = []
    for i in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rc = fn(); e1.record(); e1.synchronize()
        assert rc == 0, rc
        if i >= 3: ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts))
    print(f"{name}: {t * 1e3:.1f} us flushed, {bytes_ / t / 1e6:.0f} GB/s algorithmic", flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "mix"):
    M, K, N, A = 8192 * 64, 256, 128, 5
    Z = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") * 0.1
    Wp = torch.empty((lib.mapf_tc_packed_b_size(N, K),), device="cuda")
    lib.mapf_tc_pack_b(W.data_ptr(), K, None, 0, N, Wp.data_ptr(), st)
    aw, ab = torch.randn(A, N, device="cuda"), torch.randn(A, device="cuda")
    logits = torch.empty(M, A, device="cuda"); acts = torch.empty(M, dtype=torch.int32, device="cuda")
    timeit("mix+head 524288 x 256 -> 128", lambda: lib.mapf_tc_mix_head(M, N, K, Z.data_ptr(), K, Wp.data_ptr(), aw.data_ptr(), ab.data_ptr(), A,
                                                                     logits.data_ptr(), acts.data_ptr(), None, st), M * K * 4)
    del Z
if which in ("all", "fc"):
    M, K, N = 40960, 400, 64
    X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") * 0.1
    Wp = torch.empty((lib.mapf_tc_packed_b_size(N, K),), device="cuda")
    lib.mapf_tc_pack_b(W.data_ptr(), K, None, 0, N, Wp.data_ptr(), st)
    C_ = torch.empty(M, N, device="cuda")
    timeit("store GEMM 40960 x 400 -> 64", lambda: lib.mapf_tc_gemm(M, N, K, X.data_ptr(), K, Wp.data_ptr(), C_.data_ptr(), N, None, 1, st), M * (K + N) * 4)
if which in ("all", "da"):
    M, K, N = 40960, 64, 400
    X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") * 0.1
    tiles = (N + 127) // 128
    Wp = torch.empty((tiles * lib.mapf_tc_packed_b_size(128, K),), device="cuda")
    lib.mapf_tc_pack_b_tiles(W.data_ptr(), K, N, Wp.data_ptr(), st)
    C_ = torch.empty(M, N, device="cuda")
    timeit("tiled GEMM 40960 x 64 -> 400", lambda: lib.mapf_tc_gemm_tiles(M, N, K, X.data_ptr(), K, Wp.data_ptr(), C_.data_ptr(), N, None, 0, st), M * (K + N) * 4)
