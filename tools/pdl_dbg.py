import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import mapf_gnn_b200 as m
from helpers import load_network, model_config
def run(name, N, nopdl, steps=7):
    if nopdl: os.environ["MAPF_B200_NO_PDL"] = "1"
    else: os.environ.pop("MAPF_B200_NO_PDL", None)
    rng = np.random.default_rng(5)
    starts, goals, obst = m.make_scenarios(rng, 64, N, 14, 6)
    env = m.BatchedGraphEnv(model_config(name, N, board_size=[14, 14]), goals, starts, obst, sensing_range=5)
    ro = m.Rollout(load_network(name, num_agents=N).eval(), env)
    outs = []
    for t in range(steps):
        ro.run(1)
        torch.cuda.synchronize()
        outs.append((ro.logits.cpu().numpy().copy(), env.pos.cpu().numpy().copy(), ro.actions.cpu().numpy().copy(), env.fov.cpu().numpy().copy(), env.adj_matrix.cpu().numpy().copy()))
    return outs
def run7(name, N, nopdl):
    if nopdl: os.environ["MAPF_B200_NO_PDL"] = "1"
    else: os.environ.pop("MAPF_B200_NO_PDL", None)
    rng = np.random.default_rng(5)
    starts, goals, obst = m.make_scenarios(rng, 64, N, 14, 6)
    env = m.BatchedGraphEnv(model_config(name, N, board_size=[14, 14]), goals, starts, obst, sensing_range=5)
    ro = m.Rollout(load_network(name, num_agents=N).eval(), env)
    ro.run(7); torch.cuda.synchronize()
    return (ro.logits.cpu().numpy().copy(), env.pos.cpu().numpy().copy(), ro.actions.cpu().numpy().copy(), env.fov.cpu().numpy().copy(), env.adj_matrix.cpu().numpy().copy())
for name, N in (("gnn_k3", 5), ("baseline", 5), ("gnn_msg_k3", 20)):
    a = run(name, N, False); b = run(name, N, True)
    for t in range(7):
        d = [np.array_equal(x, y) for x, y in zip(a[t], b[t])]
        print(name, "stepwise t", t, "logits/pos/act/fov/adj equal:", d)
    c = run7(name, N, False); d7 = run7(name, N, True)
    print(name, "7 at once pdl vs nopdl:", [np.array_equal(x, y) for x, y in zip(c, d7)], " pdl7 vs stepwise:", [np.array_equal(x, y) for x, y in zip(c, a[6])], " nopdl7 vs stepwise:", [np.array_equal(x, y) for x, y in zip(d7, a[6])])
