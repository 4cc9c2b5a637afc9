import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import mapf_gnn_b200 as m
import mapf_gnn_b200.train as tr
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
w = dict(bench.WORKLOADS["c2a"], model="gnn_msg_k3")
cfg = bench.model_config(w)
net = m.build_network(cfg)
net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights("gnn_msg_k3").items()})
net = net.cuda().train()
t = tr.Trainer(net)
g = torch.Generator(device="cuda").manual_seed(0)
states = torch.randint(0, 4, (B, 5, 2, 5, 5), device="cuda", generator=g).float()
gso = (torch.rand(B, 5, 5, device="cuda", generator=g) < 0.4).float() + torch.eye(5, device="cuda")
targets = torch.randint(0, 5, (B, 5), device="cuda", generator=g)
for _ in range(4):
    t.step(states, gso, targets)
torch.cuda.synchronize()
print("ok")
