"""One launch series of the tensor-core conv stack (for ncu).  python tools/enc_tc_one.py [rows] [model]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import load_network
from mapf_gnn_b200 import _lib
from mapf_gnn_b200.ops import net_params
lib = _lib.load()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20480
net = load_network(sys.argv[2] if len(sys.argv) > 2 else "gnn_k2", num_agents=1).eval()
dims, packed = net.dims(), net.packed_weights()
p = net_params(dims); s = _lib.current_stream()
st = torch.randint(0, 4, (rows, 2, 5, 5), device="cuda").float()
Ap = torch.zeros((lib.mapf_tc_presplit_a_size(rows, 400),), device="cuda")
ea = _lib.EncoderFwdArgs(rows, st.data_ptr(), packed.data_ptr(), Ap.data_ptr())
for _ in range(6):
    lib.mapf_encoder_tc_conv_fwd(p, ea, s)
torch.cuda.synchronize()
print("ok")
