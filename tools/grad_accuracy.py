"""Development: gradient error of this package's training step and of the torch-CPU fp32 oracle against a float64 oracle, at a
large batch, under the development switches.   python tools/grad_accuracy.py [B]"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 3900
MODE = os.environ.get("GRAD_MODE")
name, N = "gnn_msg_k3", 5
rng = np.random.default_rng(12)
states = torch.tensor(rng.integers(0, 4, size=(B, N, 2, 5, 5))).float()
adj = torch.tensor((rng.random((B, N, N)) < 0.3).astype(np.float32))
adj[:, torch.arange(N), torch.arange(N)] = 0
gso = adj + torch.eye(N)
targets = torch.tensor(rng.integers(0, 5, size=(B, N)))
if MODE is None:
    from helpers import oracle_params
    from oracle import model_oracle as mo
    params = oracle_params(name)
    _, g32, _ = mo.loss_and_grads(params, states, gso, targets)
    p64 = {k: (v.double() if v.is_floating_point() else v) for k, v in params.items()}
    _, g64, _ = mo.loss_and_grads(p64, states.double(), gso.double(), targets)
    np.savez("/tmp/g64.npz", **{k: v.numpy() for k, v in g64.items()})
    err = {k: float((g32[k].double() - g64[k]).abs().max() / g64[k].abs().max()) for k in g64}
    print("torch fp32 oracle vs fp64: worst", max(err, key=err.get), f"{max(err.values()):.2e}", {k: f"{v:.1e}" for k, v in err.items() if not (k.startswith("convLayers") and k.endswith("bias") and int(k.split(".")[1]) % 3 == 0)})
    for mode in ("default", "MAPF_B200_TRAIN_FFMA_CONV", "MAPF_B200_TRAIN_BN_PASS", "MAPF_B200_TRAIN_FFMA_BWD"):
        env = dict(os.environ, GRAD_MODE=mode)
        if mode != "default":
            env[mode] = "1"
        print(subprocess.run([sys.executable, __file__, str(B)], env=env, capture_output=True, text=True).stdout.strip())
else:
    from helpers import load_network
    g64 = np.load("/tmp/g64.npz")
    net = load_network(name, num_agents=N).train()
    loss = torch.nn.functional.cross_entropy(net(states.cuda(), gso.cuda()).reshape(-1, 5), targets.cuda().reshape(-1))
    loss.backward()
    err = {k: float(np.abs(p.grad.cpu().numpy().astype(np.float64) - g64[k]).max() / np.abs(g64[k]).max()) for k, p in net.named_parameters()}
    print(f"{MODE}: worst", max(err, key=err.get), f"{max(err.values()):.2e}", {k: f"{v:.1e}" for k, v in err.items() if not (k.startswith("convLayers") and k.endswith("bias") and int(k.split(".")[1]) % 3 == 0)})
