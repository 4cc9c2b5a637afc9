import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import mapf_gnn_b200 as m
from mapf_gnn_b200 import _lib
lib = _lib.load()
lib.mapf_enc_set_debug.argtypes = [C.c_void_p]
w = bench.WORKLOADS["c2a"]; cfg = bench.model_config(w)
net = m.build_network(cfg); net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(w["model"]).items()}); net = net.cuda().eval()
states = torch.randint(0, 4, (4096, 5, 2, 5, 5), device="cuda").float()
dims, packed = net.dims(), net.packed_weights()
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
lib.mapf_enc_set_debug(dbg.data_ptr())
for _ in range(3):
    dbg.zero_(); torch.ops.mapf.encoder_fwd(dims, states, packed); torch.cuda.synchronize()
d = dbg.cpu().tolist(); t0 = d[0]
names = ["staged", "conv1", "conv2", "conv3", "fc"]
i = 1
for tile in range(5):
    row = d[i:i + 5]
    prev = d[i - 1]
    print("tile", tile, {n: row[k] - (prev if k == 0 else row[k - 1]) for k, n in enumerate(names)})
    i += 5
print("total", d[i - 1] - t0)
