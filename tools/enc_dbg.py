"""Per-phase / per-warp clock breakdown of the persistent encoder kernel (block 0).  Needs a debug build:
MAPF_B200_NVCC_EXTRA=-DMAPF_ENC_DEBUG python tools/enc_dbg.py   (rebuilds the library with the hooks; rebuild
without the variable afterwards)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import mapf_gnn_b200 as m
from mapf_gnn_b200 import _lib
lib = _lib.load()
lib.mapf_enc_set_debug.argtypes = [C.c_void_p]
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2a"]; cfg = bench.model_config(w)
net = m.build_network(cfg); net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(w["model"]).items()}); net = net.cuda().eval()
states = torch.randint(0, 4, (w["E"], w["N"], 2, 5, 5), device="cuda").float()
dims, packed = net.dims(), net.packed_weights()
dbg = torch.zeros(1024, dtype=torch.int64, device="cuda")
lib.mapf_enc_set_debug(dbg.data_ptr())
if os.environ.get("ENC_DBG_SPLIT"):
    from mapf_gnn_b200.ops import net_params
    p = net_params(dims); rows = w["E"] * w["N"]
    Ap = torch.empty((lib.mapf_tc_presplit_a_size(rows, dims[6]),), device="cuda")
    es = _lib.EncoderFwdArgs(rows, states.data_ptr(), packed.data_ptr(), Ap.data_ptr())
    run = lambda: lib.mapf_encoder_conv_split_fwd(p, es, _lib.current_stream())
else:
    run = lambda: torch.ops.mapf.encoder_fwd(dims, states, packed)
for _ in range(3):
    dbg.zero_(); run(); torch.cuda.synchronize()
d = dbg.cpu().numpy(); t0 = int(d[1023])
a = d[:1024 - 0][:8 * 8 * 8 * 2].reshape(8, 8, 8, 2)     # tile, phase, warp, (arrive, leave)
names = ["stage", "conv1", "conv2", "conv3", "fcA0", "fcB0", "fcA1", "fcB1"]
prev = t0
for tile in range(8):
    if a[tile, 0, 0, 1] == 0:
        break
    out = []
    for ph in range(8):
        if a[tile, ph, 0, 1] == 0:
            continue
        arr, leave = a[tile, ph, :, 0], a[tile, ph, :, 1]
        out.append(f"{names[ph]}: {int(leave.max()) - prev} (arrive spread {int(arr.max() - arr.min())}, mean wait {int((leave - arr).mean())})")
        prev = int(leave.max())
    print("tile", tile, " | ".join(out))
print("total", prev - t0, "first sync after start", int(a[0, 0, 0, 1]) - t0)
h = d[900:924].reshape(8, 3)
print("half-tile last layer, per warp: loop", (h[:, 1] - h[:, 0]).tolist(), "epilogue", (h[:, 2] - h[:, 1]).tolist(), "start spread", int(h[:, 0].max() - h[:, 0].min()))
c = d[950:974].reshape(8, 3)
print("full-tile conv (last 16->16 layer run), per warp: loop", (c[:, 1] - c[:, 0]).tolist(), "epilogue", (c[:, 2] - c[:, 1]).tolist(), "start spread", int(c[:, 0].max() - c[:, 0].min()), "end spread", int(c[:, 2].max() - c[:, 2].min()))
