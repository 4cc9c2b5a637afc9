import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mapf_gnn_b200 import _lib
lib = _lib.load(); st = _lib.current_stream()
M, N, K = 20480, 64, 400
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); bias = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda")
Bp = torch.empty((lib.mapf_tc_packed_b_size(N, K),), device="cuda")
lib.mapf_tc_pack_b(B.data_ptr(), K, None, 0, N, Bp.data_ptr(), st)
for _ in range(6):
    lib.mapf_tc_gemm(M, N, K, A.data_ptr(), K, Bp.data_ptr(), out.data_ptr(), N, bias.data_ptr(), 1, st)
torch.cuda.synchronize()
print("ok")
