import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mapf_gnn_b200 import _lib
lib = _lib.load(); st = _lib.current_stream()
lib.mapf_tc_set_debug.argtypes = [C.c_void_p]
dbg = torch.zeros(160, dtype=torch.int64, device="cuda")
for (M, N, K, head) in ((20480, 64, 400, 0), (20480, 128, 128, 0), (20480, 128, 128, 1)):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    aw = torch.randn(5, N, device="cuda"); ab = torch.randn(5, device="cuda")
    lg = torch.empty(M, 5, device="cuda"); ac = torch.empty(M, dtype=torch.int32, device="cuda")
    Bp = torch.empty((lib.mapf_tc_packed_b_size(N, K),), device="cuda")
    lib.mapf_tc_pack_b(B.data_ptr(), K, None, 0, N, Bp.data_ptr(), st)
    lib.mapf_tc_set_debug(dbg.data_ptr())
    for _ in range(3):
        dbg.zero_()
        if head:
            lib.mapf_tc_mix_head(M, N, K, A.data_ptr(), K, Bp.data_ptr(), aw.data_ptr(), ab.data_ptr(), 5, lg.data_ptr(), ac.data_ptr(), None, st)
        else:
            lib.mapf_tc_gemm(M, N, K, A.data_ptr(), K, Bp.data_ptr(), out.data_ptr(), N, bias.data_ptr(), 1, st)
        torch.cuda.synchronize()
    d = dbg.cpu().tolist()
    print(M, N, K, "head" if head else "store", "setup %d | producers done %d | mma done %d | epilogue end %d | exit %d" % tuple(d[i] - d[0] for i in (1, 2, 3, 4, 5)))
