"""Per-kernel CUDA-event timings of one rollout step (warm caches), for development."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import mapf_gnn_b200 as m

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2a")
ap.add_argument("--episodes", type=int, default=0)
ap.add_argument("--iters", type=int, default=50)
args = ap.parse_args()
w = bench.WORKLOADS[args.workload]
E, N, H = args.episodes or w["E"], w["N"], w["H"]
rng = np.random.default_rng(0)
starts, goals, obst = m.make_scenarios(rng, E, N, H, w["M"])
cfg = bench.model_config(w)
net = m.build_network(cfg)
net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(w["model"]).items()})
net = net.cuda().eval()
env = m.BatchedGraphEnv(cfg, goals, starts, obst, sensing_range=w["r"])
ro = m.Rollout(net, env)
ro.run(2)
dims, packed = net.dims(), net.packed_weights()
x = torch.ops.mapf.encoder_fwd(dims, env.fov, packed)


def timeit(fn, iters=args.iters):
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


t_enc = timeit(lambda: torch.ops.mapf.encoder_fwd(dims, env.fov, packed))
t_pol = timeit(lambda: torch.ops.mapf.policy_fwd(dims, net.message, env.fov, env.adj_matrix, packed))
t_env = timeit(lambda: env.observe())
t_step = timeit(lambda: ro.run(1))
t_roll = timeit(lambda: ro.run(8), iters=10) / 8
print(f"{args.workload} E={E} N={N}: encoder {t_enc:.1f} us | policy(enc+gf+head) {t_pol:.1f} us -> gf ~{t_pol - t_enc:.1f} us | "
      f"env(observe) {t_env:.1f} us | step {t_step:.1f} us | step inside run(8) {t_roll:.1f} us | "
      f"{E * N / t_roll:.1f} M agent-steps/s")

# ---- pieces of the tensor-core path through the C ABI ----
from mapf_gnn_b200 import _lib
from mapf_gnn_b200.ops import net_params
import ctypes as C
lib = _lib.load()
p = net_params(dims)
rows = E * N
act = torch.empty((rows, dims[6]), device="cuda")
xx = torch.empty((rows, dims[7]), device="cuda")
st = _lib.current_stream()
ea = _lib.EncoderFwdArgs(rows, env.fov.data_ptr(), packed.data_ptr(), act.data_ptr())
t_conv = timeit(lambda: lib.mapf_encoder_conv_fwd(p, ea, st))
fcw = net.compressMLP[0].weight.detach().contiguous()
fcb = net.compressMLP[0].bias.detach().contiguous()
Bp = torch.empty((lib.mapf_tc_packed_b_size(dims[7], dims[6]),), device="cuda")
lib.mapf_tc_pack_b(fcw.data_ptr(), dims[6], None, 0, dims[7], Bp.data_ptr(), st)
t_fc = timeit(lambda: lib.mapf_tc_gemm(rows, dims[7], dims[6], act.data_ptr(), dims[6], Bp.data_ptr(), xx.data_ptr(),
                                       dims[7], fcb.data_ptr(), 1, st))
print(f"   conv-only {t_conv:.1f} us | tc FC gemm [{rows}x{dims[6]}]x[{dims[7]}] {t_fc:.1f} us")
Ap = torch.empty((lib.mapf_tc_presplit_a_size(rows, dims[6]),), device="cuda")
es = _lib.EncoderFwdArgs(rows, env.fov.data_ptr(), packed.data_ptr(), Ap.data_ptr())
x2 = torch.empty((rows, dims[7]), device="cuda")
t_convs = timeit(lambda: lib.mapf_encoder_conv_split_fwd(p, es, st))
t_fcs = timeit(lambda: lib.mapf_tc_fc_presplit(rows, dims[7], dims[6], Ap.data_ptr(), Bp.data_ptr(), x2.data_ptr(), dims[7],
                                                 fcb.data_ptr(), 1, st))
t_both = timeit(lambda: (lib.mapf_encoder_conv_split_fwd(p, es, st),
                         lib.mapf_tc_fc_presplit(rows, dims[7], dims[6], Ap.data_ptr(), Bp.data_ptr(), x2.data_ptr(),
                                                 dims[7], fcb.data_ptr(), 1, st)))
torch.cuda.synchronize()
ref = torch.ops.mapf.encoder_fwd(dims, env.fov, packed)
err = float((x2 - ref).abs().max() / ref.abs().max())
print(f"   conv + pre-split A {t_convs:.1f} us | all-TMA tc FC {t_fcs:.1f} us | back to back {t_both:.1f} us | max rel diff vs fused FFMA encoder {err:.2e}")
if dims[8] > 0:
    KT = (dims[8] + (1 if net.message else 0)) * dims[7]
    z = torch.empty((rows, KT), device="cuda")
    g = _lib.GraphFilterFwdArgs()
    g.batch, g.agents, g.in_dim, g.out_dim, g.taps = E, N, dims[7], dims[9], dims[8]
    g.add_self_loop, g.has_skip = int(not net.message), int(net.message)
    g.x, g.x_sb, g.x_sn, g.x_sf = xx.data_ptr(), N * dims[7], dims[7], 1
    g.gso, g.z = env.adj_matrix.data_ptr(), z.data_ptr()
    t_taps = timeit(lambda: lib.mapf_graph_taps(g, st))
    gf = net.GFL[0]
    W = gf.W.detach().contiguous()
    W2 = gf.W2.detach().contiguous() if net.message else None
    Wp = torch.empty((lib.mapf_tc_packed_b_size(dims[9], KT),), device="cuda")
    lib.mapf_tc_pack_b(W.data_ptr(), dims[8] * dims[7], None if W2 is None else W2.data_ptr(),
                       dims[7] if W2 is not None else 0, dims[9], Wp.data_ptr(), st)
    aw = net.actionMLP[0].weight.detach().contiguous()
    ab = net.actionMLP[0].bias.detach().contiguous()
    lg = torch.empty((rows, 5), device="cuda")
    ac = torch.empty((rows,), dtype=torch.int32, device="cuda")
    t_mix = timeit(lambda: lib.mapf_tc_mix_head(rows, dims[9], KT, z.data_ptr(), KT, Wp.data_ptr(), aw.data_ptr(),
                                                ab.data_ptr(), 5, lg.data_ptr(), ac.data_ptr(), None, st))
    print(f"   taps {t_taps:.1f} us | tc mix+head [{rows}x{KT}]x[{dims[9]}] {t_mix:.1f} us")
