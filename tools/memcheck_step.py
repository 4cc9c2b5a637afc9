"""Small rollouts + standalone encoder calls over ragged sizes: a quick smoke of the whole tensor-core pipeline (and the
input for compute-sanitizer where that tool is available -- it is closed on this pool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import mapf_gnn_b200 as m
from helpers import load_network, model_config
from mapf_gnn_b200 import _lib
from mapf_gnn_b200.ops import net_params
lib = _lib.load()
for name, E, N in (("gnn_k3", 37, 5), ("gnn_msg_k3", 9, 20), ("gnn_k2", 3, 5)):
    rng = np.random.default_rng(E)
    starts, goals, obst = m.make_scenarios(rng, E, N, 12, 4)
    cfg = model_config(name, N, board_size=[12, 12])
    env = m.BatchedGraphEnv(cfg, goals, starts, obst, sensing_range=4)
    net = load_network(name, num_agents=N).eval()
    ro = m.Rollout(net, env)
    ro.run(3)
    torch.cuda.synchronize()
    print(name, "rollout ok", int(env.done.sum()))
net = load_network("gnn_k3", num_agents=1).eval()
p = net_params(net.dims())
for rows in (1, 5, 131, 600):
    st = torch.randint(0, 4, (rows, 2, 5, 5), device="cuda").float()
    Ap = torch.empty((lib.mapf_tc_presplit_a_size(rows, 400),), device="cuda")
    ea = _lib.EncoderFwdArgs(rows, st.data_ptr(), net.packed_weights().data_ptr(), Ap.data_ptr())
    _lib.check(lib.mapf_encoder_tc_conv_fwd(p, ea, _lib.current_stream()), "tc_conv")
    torch.cuda.synchronize()
print("encoder ok")
