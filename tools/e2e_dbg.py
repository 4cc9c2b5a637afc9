"""Where does the host-stepped (end-to-end) step go?  Times launch / wait on the host and the per-group step latency."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import mapf_gnn_b200 as m
w = bench.WORKLOADS["c2a"]; cfg = bench.model_config(w)
E, N, H = w["E"], w["N"], w["H"]
starts, goals, obst = m.make_scenarios(np.random.default_rng(0), E, N, H, w["M"])
net = m.build_network(cfg); net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(w["model"]).items()}); net = net.cuda().eval()
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pl = m.PipelinedHostRollout(net, cfg, goals, starts, obst, groups=G, sensing_range=w["r"])
print("group sizes", [s.stop - s.start for s in pl.slices])
for g, h in enumerate(pl.groups):           # latency of one group's step, alone
    for _ in range(5): h.step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): h.launch(); h.wait()
    print(f"group {g} alone: {(time.perf_counter() - t0) / 50 * 1e6:.1f} us per step")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(h.stream):
        ev0.record()
        for _ in range(20): h.graphs[0].replay(); h.graphs[1].replay()
        ev1.record()
    ev1.synchronize(); print(f"   back-to-back graph replays: {ev0.elapsed_time(ev1) / 40 * 1e3:.1f} us per step (device)")
pl.reset()
tl = tw = 0.0
K = 200
torch.cuda.synchronize(); t0 = time.perf_counter()
pl.start()
for i in range(K):
    for g in range(G):
        a = time.perf_counter(); out = pl.groups[g].wait(); b = time.perf_counter()
        if i + 1 < K: pl.groups[g].launch()
        c = time.perf_counter(); tw += b - a; tl += c - b
tot = time.perf_counter() - t0
print(f"pipelined G={G}: {tot / K * 1e6:.1f} us per full step; host in wait {tw / K * 1e6:.1f} us, in launch {tl / K * 1e6:.1f} us")
