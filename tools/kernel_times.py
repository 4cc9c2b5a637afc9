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
