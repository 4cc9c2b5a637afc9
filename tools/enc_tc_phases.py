"""Per-phase clocks of the tensor-core conv kernel (block 0).  Needs a debug build:
MAPF_B200_NVCC_EXTRA=-DMAPF_ENC_DEBUG python tools/enc_tc_phases.py"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import load_network
from mapf_gnn_b200 import _lib
from mapf_gnn_b200.ops import net_params
lib = _lib.load()
lib.mapf_enc_tc_set_debug.argtypes = [C.c_void_p]
net = load_network("gnn_k3", num_agents=1).eval()
dims, packed = net.dims(), net.packed_weights()
p = net_params(dims)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20480
st = torch.randint(0, 4, (rows, 2, 5, 5), device="cuda").float()
Ap = torch.empty((lib.mapf_tc_presplit_a_size(rows, 400),), device="cuda")
ea = _lib.EncoderFwdArgs(rows, st.data_ptr(), packed.data_ptr(), Ap.data_ptr())
dbg = torch.zeros(16 * 20 * 10, dtype=torch.int64, device="cuda")
lib.mapf_enc_tc_set_debug(dbg.data_ptr())
for _ in range(3):
    dbg.zero_(); lib.mapf_encoder_tc_conv_fwd(p, ea, _lib.current_stream()); torch.cuda.synchronize()
d = dbg.cpu().numpy().reshape(16, 20, 10)
t0 = d[0, :, 0].min()
names = ["top", "synced", "mma done", "ld0", "sums", "stored"]
for it in range(16):
    for wg in range(5):
        w = d[it, wg * 4:(wg + 1) * 4]
        if w[0, 0] == 0:
            continue
        rel = w - t0
        print(f"it {it} wg {wg}: " + " | ".join(f"{names[i]} {int(rel[:, i].min())}..{int(rel[:, i].max())}" for i in range(6))
              + f" || wait-sync {int((w[:, 1] - w[:, 0]).mean())} mma {int((w[:, 2] - w[:, 1]).mean())} ld0 {int((w[:, 3] - w[:, 2]).mean())}"
              f" sums {int((w[:, 4] - w[:, 3]).mean())} store {int((w[:, 5] - w[:, 4]).mean())}"
              f" || issuer: fence {int((w[:, 6] - w[:, 1]).mean())} issue {int((w[:, 7] - w[:, 6]).mean())} commit {int((w[:, 8] - w[:, 7]).mean())} done-after-commit {int((w[:, 2] - w[:, 8]).mean())}")
