"""Development check of the tensor-core conv stack (enc_tc.cu): per-depth parity of the conv output against torch on
the CPU, the full encoder (conv + pixel-major FC) against the oracle, and timings at a benchmark size."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.nn.functional as F

from helpers import load_network, oracle_params
from mapf_gnn_b200 import _lib
from mapf_gnn_b200.ops import net_params

lib = _lib.load()
net = load_network("gnn_k3", num_agents=1).eval()
params = oracle_params("gnn_k3")
g = net.param_groups()
d = lambda ts: [t.detach() for t in ts]


def untile(Ap, rows, K):
    nblk, nch = (rows + 127) // 128, (K + 15) // 16
    words = nblk * nch * 2048
    h = Ap[:words].view(torch.float16).cpu().numpy().astype(np.float64)
    a = h.reshape(nblk, nch, 2, 16, 2, 8, 8).transpose(2, 0, 3, 5, 1, 4, 6).reshape(2, nblk * 128, nch * 16)
    inv = Ap[words:words + rows].cpu().numpy().astype(np.float64)[:, None]
    return ((a[0] + a[1])[:rows, :K] * inv), (a[0][:rows, :K] * inv), (a[1][:rows, :K] * inv)


def conv_ref(states, depth):
    h = states
    for l in range(depth):
        li = 3 * l
        h = F.conv2d(h, params[f"convLayers.{li}.weight"], params[f"convLayers.{li}.bias"], padding=1)
        bn = f"convLayers.{li + 1}"
        h = F.batch_norm(h, params[f"{bn}.running_mean"], params[f"{bn}.running_var"], params[f"{bn}.weight"],
                         params[f"{bn}.bias"], False, 0.1, 1e-5)
        h = F.relu(h)
    return h.permute(0, 2, 3, 1).reshape(states.shape[0], -1).numpy()     # pixel-major (y, x, c)


def run_conv(depth, states):
    rows = states.shape[0]
    dims = [depth, 2] + [16] * depth + [0] * (4 - depth) + [400, 64, 0, 0, 5]
    packed = torch.ops.mapf.pack_weights(dims, d(g["conv_w"][:depth]), d(g["conv_b"][:depth]), d(g["bn_gamma"][:depth]),
                                         d(g["bn_beta"][:depth]), d(g["bn_mean"][:depth]), d(g["bn_var"][:depth]),
                                         g["fc_w"].detach(), g["fc_b"].detach(), None, None,
                                         torch.zeros(5, 64, device="cuda"), torch.zeros(5, device="cuda"))
    p = net_params(dims)
    assert lib.mapf_encoder_tc_supported(p) == 1
    Ap = torch.full((lib.mapf_tc_presplit_a_size(rows, 400),), float("nan"), device="cuda")
    st = states.cuda()
    ea = _lib.EncoderFwdArgs(rows, st.data_ptr(), packed.data_ptr(), Ap.data_ptr())
    _lib.check(lib.mapf_encoder_tc_conv_fwd(p, ea, _lib.current_stream()), "tc_conv")
    torch.cuda.synchronize()
    return Ap, p, packed, st


rng = np.random.default_rng(0)
for rows in (4, 37, 1000):
    states = torch.tensor(rng.integers(0, 4, size=(rows, 2, 5, 5))).float()
    states[::5] *= 0.61
    for depth in (1, 2, 3):
        with torch.no_grad():
            ref = conv_ref(states, depth)
        Ap, *_ = run_conv(depth, states)
        got, hi, lo = untile(Ap, rows, 400)
        err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30)
        print(f"rows {rows} depth {depth}: rel err {err:.3e}  finite {np.isfinite(got).all()}  max|ref| {np.abs(ref).max():.3f}")
        if not (err < 1e-4):
            bad = np.argwhere(~(np.abs(got - ref) <= 1e-4 * np.abs(ref).max()))
            print("   first mismatches (row, k = pix*16+c):", bad[:8].tolist())
            for r, k in bad[:4]:
                print(f"   row {r} pix {k // 16} c {k % 16}: got {got[r, k]:.6f} (hi {hi[r, k]:.6f} lo {lo[r, k]:.3e}) ref {ref[r, k]:.6f}")
            print("   per-pixel max err of row 0:", np.abs(got[0] - ref[0]).reshape(25, 16).max(1).round(4).tolist())
            print("   per-channel max err of row 0:", np.abs(got[0] - ref[0]).reshape(25, 16).max(0).round(4).tolist())

# ---- full encoder: conv (tensor cores) + pixel-major FC against the oracle ----
from oracle import model_oracle as mo
rows = 2400
states = torch.tensor(rng.integers(0, 4, size=(rows, 1, 2, 5, 5))).float()
with torch.no_grad():
    ref = mo.encode_agents(params, states)[:, :, 0].numpy()
dims, packed = net.dims(), net.packed_weights()
p = net_params(dims)
st = states.cuda()
Ap = torch.empty((lib.mapf_tc_presplit_a_size(rows, 400),), device="cuda")
Bp = torch.empty((lib.mapf_tc_packed_b_f16_size(64, 400),), device="cuda")
fcw, fcb = g["fc_w"].detach().contiguous(), g["fc_b"].detach().contiguous()
s = _lib.current_stream()
_lib.check(lib.mapf_tc_pack_b_f16_pixel_major(fcw.data_ptr(), 16, 64, Bp.data_ptr(), s), "pack pm")
x = torch.empty((rows, 64), device="cuda")
ea = _lib.EncoderFwdArgs(rows, st.data_ptr(), packed.data_ptr(), Ap.data_ptr())
_lib.check(lib.mapf_encoder_tc_conv_fwd(p, ea, s), "tc_conv")
_lib.check(lib.mapf_tc_fc_presplit_f16(rows, 64, 400, Ap.data_ptr(), Bp.data_ptr(), x.data_ptr(), 64, fcb.data_ptr(), 1, s), "fc")
torch.cuda.synchronize()
got = x.cpu().numpy()
print(f"full encoder rows {rows}: rel err {np.abs(got - ref).max() / np.abs(ref).max():.3e}")


# ---- timings ----
def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters * 1e3


for rows in (20480, 81920):
    st = torch.tensor(rng.integers(0, 4, size=(rows, 2, 5, 5))).float().cuda()
    Ap = torch.empty((lib.mapf_tc_presplit_a_size(rows, 400),), device="cuda")
    x = torch.empty((rows, 64), device="cuda")
    ea = _lib.EncoderFwdArgs(rows, st.data_ptr(), packed.data_ptr(), Ap.data_ptr())
    ex = _lib.EncoderFwdArgs(rows, st.data_ptr(), packed.data_ptr(), x.data_ptr())
    t_conv = timeit(lambda: lib.mapf_encoder_tc_conv_fwd(p, ea, s))
    t_fc = timeit(lambda: lib.mapf_tc_fc_presplit_f16(rows, 64, 400, Ap.data_ptr(), Bp.data_ptr(), x.data_ptr(), 64, fcb.data_ptr(), 1, s))
    t_both = timeit(lambda: (lib.mapf_encoder_tc_conv_fwd(p, ea, s),
                             lib.mapf_tc_fc_presplit_f16(rows, 64, 400, Ap.data_ptr(), Bp.data_ptr(), x.data_ptr(), 64, fcb.data_ptr(), 1, s)))
    t_old = timeit(lambda: lib.mapf_encoder_fwd(p, ex, s))
    print(f"rows {rows}: tc conv {t_conv:.1f} us | presplit FC {t_fc:.1f} us | both {t_both:.1f} us | FFMA2 encoder {t_old:.1f} us")
