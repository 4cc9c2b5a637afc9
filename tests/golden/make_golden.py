"""Generate the golden fixtures by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py            # needs /root/reference; writes tests/golden/*.npz

The GPU box has no /root/reference, so everything the parity tests need from the reference is
frozen here: env traces (positions / FOV / adjacency / done / metrics under teacher-forced
actions), model logits on the shipped weights, one full training step (loss, gradients, BN
running stats, post-Adam parameters), and the shipped weights themselves re-saved as .npz
(data, not code).  Import shims: the reference imports matplotlib and gymnasium at module
import but only touches them in __init__/render; inert stubs are enough.
"""
from __future__ import annotations

import copy
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("MAPF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def install_shims():
    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Anything()

        def __getattr__(self, name):
            return _Anything()

    def stub(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        def _getattr(attr):
            if attr.startswith("__"):
                raise AttributeError(attr)
            return _Anything()

        mod.__getattr__ = _getattr                          # type: ignore[attr-defined]
        sys.modules[name] = mod
        return mod

    mpl = stub("matplotlib")
    mpl.pyplot = stub("matplotlib.pyplot")
    mpl.cm = stub("matplotlib.cm")
    mpl.colors = stub("matplotlib.colors")
    gym = stub("gymnasium", Env=object)
    gym.spaces = stub("gymnasium.spaces")
    sys.path[:0] = [REF, os.path.join(REF, "models")]


def import_reference():
    install_shims()
    from grid.env_graph_gridv1 import GraphEnv, create_goals, create_obstacles   # noqa
    import models.framework_gnn as fg                                            # noqa
    import framework_gnn_message as fm                                           # noqa
    import framework_baseline as fb                                              # noqa
    return GraphEnv, create_goals, create_obstacles, fg, fm, fb


BASE_CFG = {
    "device": "cpu", "map_shape": [5, 5], "num_actions": 5, "last_convs": [400], "channels": [16, 16, 16],
    "node_dims": [128], "graph_filters": [3], "encoder_dims": [64], "encoder_layers": 1, "action_layers": 1,
    "min_time": 1,
}


def env_cfg(n, h, max_time=32):
    return {"num_agents": n, "board_size": [h, h], "max_time": max_time, "min_time": 1}


# ------------------------------------------------------------------------------------------
# env traces
# ------------------------------------------------------------------------------------------
ENV_CASES = [
    # name, N, H, M, sensing_range, T, seed, flavour
    ("c1_5a_18", 5, 18, 5, 6, 32, 100, "unique"),
    ("c2_5a_28", 5, 28, 8, 6, 16, 101, "unique"),
    ("c3_20a_32", 20, 32, 40, 6, 12, 102, "unique"),
    ("c5_64a_64", 64, 64, 200, 6, 6, 103, "unique"),
    ("r4_5a_28", 5, 28, 8, 4, 16, 104, "unique"),
    ("crowded_6a_4", 6, 4, 2, 6, 24, 105, "unique"),
    ("crowded_12a_5", 12, 5, 3, 3, 24, 106, "colocated"),
    ("noobst_10a_10", 10, 10, 0, 2.5, 16, 107, "colocated"),
    ("single_1a_6", 1, 6, 3, 6, 10, 108, "unique"),
    ("onobst_4a_7", 4, 7, 6, 6, 16, 109, "on_obstacle"),
    ("goalrun_3a_8", 3, 8, 0, 6, 20, 110, "to_goal"),
    ("dense_33a_12", 33, 12, 10, 5, 10, 111, "unique"),
]


def run_env_case(GraphEnv, name, n, h, m, r, T, seed, flavour):
    rng = np.random.RandomState(seed)
    cells = rng.permutation(h * h)
    obst = np.stack([cells[:m] % h, cells[:m] // h], axis=1).astype(np.int64) if m > 0 else None
    free = cells[m:]
    goals = np.stack([free[:n] % h, free[:n] // h], axis=1).astype(np.int64)
    if flavour == "colocated":
        pick = rng.choice(free, size=n, replace=True)
    else:
        pick = rng.choice(free, size=n, replace=False)
    starts = np.stack([pick % h, pick // h], axis=1).astype(np.int64)
    if flavour == "on_obstacle":
        starts[0] = obst[0]
        starts[1] = obst[1]
    env = GraphEnv(env_cfg(n, h), goal=goals.copy(), sensing_range=r,
                   starting_positions=starts.copy(), obstacles=None if obst is None else obst.copy())
    env.starting_positions = starts.copy()        # the first step mutates the aliased array (:75-76,206)
    obs = env.reset()
    pos = [np.stack([env.positionX, env.positionY], axis=1).copy()]
    fov = [obs["fov"].copy()]
    adj = [obs["adj_matrix"].copy()]
    acts, dones = [], []
    for t in range(T):
        if flavour == "to_goal":                  # greedy walk so the episode actually finishes
            a = np.zeros(n, dtype=np.int64)
            for i in range(n):
                dx = goals[i, 0] - env.positionX[i]
                dy = goals[i, 1] - env.positionY[i]
                a[i] = (1 if dx > 0 else 3) if dx != 0 else ((2 if dy > 0 else 4) if dy != 0 else 0)
        else:
            a = rng.randint(0, 5, size=n)
        obs, _, done, _ = env.step(a, np.ones((n, 1)))
        acts.append(a.copy())
        dones.append(done)
        pos.append(np.stack([env.positionX, env.positionY], axis=1).copy())
        fov.append(obs["fov"].copy())
        adj.append(obs["adj_matrix"].copy())
        assert obs["distances"] is obs["adj_matrix"]
    fov = np.stack(fov)
    adj = np.stack(adj)
    assert set(np.unique(fov)).issubset({0.0, 1.0, 2.0, 3.0}) and set(np.unique(adj)).issubset({0.0, 1.0})
    success, flow = env.computeMetrics()
    return {
        f"{name}/meta": np.array([n, h, m, T, 32], dtype=np.int64), f"{name}/range": np.array([r], dtype=np.float64),
        f"{name}/starts": starts.astype(np.int16), f"{name}/goals": goals.astype(np.int16),
        f"{name}/obstacles": (np.zeros((0, 2), np.int16) if obst is None else obst.astype(np.int16)),
        f"{name}/actions": np.stack(acts).astype(np.uint8), f"{name}/done": np.array(dones, dtype=np.bool_),
        f"{name}/pos": np.stack(pos).astype(np.int16), f"{name}/fov": fov.astype(np.uint8),
        f"{name}/adj": adj.astype(np.uint8), f"{name}/board_final": env.board.astype(np.uint8),
        f"{name}/metrics": np.array([success, flow], dtype=np.float64),
    }


# ------------------------------------------------------------------------------------------
# models
# ------------------------------------------------------------------------------------------
def collect_observations(GraphEnv, rng, episodes, n, h, m, r, steps):
    """States / adjacency recorded from reference rollouts under random actions (real FOV values)."""
    S, A = [], []
    for _ in range(episodes):
        cells = rng.permutation(h * h)
        obst = np.stack([cells[:m] % h, cells[:m] // h], axis=1).astype(np.int64)
        goals = np.stack([cells[m:m + n] % h, cells[m:m + n] // h], axis=1).astype(np.int64)
        pick = rng.choice(cells[m:], size=n, replace=False)
        starts = np.stack([pick % h, pick // h], axis=1).astype(np.int64)
        env = GraphEnv(env_cfg(n, h), goal=goals, sensing_range=r, starting_positions=starts.copy(), obstacles=obst)
        env.starting_positions = starts.copy()
        obs = env.reset()
        for t in range(steps):
            obs, _, _, _ = env.step(rng.randint(0, 5, size=n), np.ones((n, 1)))
            S.append(obs["fov"].copy())
            A.append(obs["adj_matrix"].copy())
    return np.stack(S), np.stack(A)


def state_dict_to_npz(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def main():
    GraphEnv, create_goals, create_obstacles, fg, fm, fb = import_reference()
    torch.set_num_threads(4)

    # ---- env ----
    env_out = {}
    for case in ENV_CASES:
        env_out.update(run_env_case(GraphEnv, *case))
    env_out["names"] = np.array([c[0] for c in ENV_CASES])
    np.savez_compressed(os.path.join(HERE, "env_cases.npz"), **env_out)
    print("env_cases.npz", len(ENV_CASES), "cases")

    # ---- shipped weights -> npz ----
    shipped = {
        "gnn_k3": (fg.Network, {"graph_filters": [3]}),
        "gnn_k2": (fg.Network, {"graph_filters": [2]}),
        "gnn_msg_k3": (fm.Network, {"graph_filters": [3]}),
        "baseline": (fb.Network, {"channels": [32, 16, 16]}),
    }
    nets = {}
    for name, (cls, over) in shipped.items():
        cfg = dict(BASE_CFG, num_agents=5, **over)
        net = cls(cfg)
        sd = torch.load(os.path.join(REF, "trained_models", name, "model.pt"), map_location="cpu")
        print(name, net.load_state_dict(sd))
        net.eval()
        nets[name] = (cls, cfg, net)
        np.savez_compressed(os.path.join(HERE, f"weights_{name}.npz"), **state_dict_to_npz(net.state_dict()))

    # ---- inference logits on shipped weights ----
    rng = np.random.RandomState(7)
    model_out = {}
    obs_sets = {
        "gnn_k3": collect_observations(GraphEnv, rng, 8, 5, 18, 5, 6, 6),
        "gnn_k2": collect_observations(GraphEnv, rng, 8, 5, 28, 8, 6, 6),
        "baseline": collect_observations(GraphEnv, rng, 8, 5, 28, 8, 6, 6),
        "gnn_msg_k3": collect_observations(GraphEnv, rng, 3, 20, 32, 40, 6, 6),
    }
    for name, (cls, cfg, net) in nets.items():
        S, A = obs_sets[name]
        n = S.shape[1]
        cfg_n = dict(cfg, num_agents=n)
        net_n = cls(cfg_n)
        net_n.load_state_dict(net.state_dict())
        net_n.eval()
        st = torch.tensor(S).float()
        gs = torch.tensor(A).float()
        with torch.no_grad():
            logits = net_n(st) if name == "baseline" else net_n(st, gs)
        model_out[f"{name}/states"] = S.astype(np.uint8)
        model_out[f"{name}/gso"] = A.astype(np.uint8)
        model_out[f"{name}/logits"] = logits.numpy()
        print(name, "logits", tuple(logits.shape), float(logits.abs().max()))

    # ---- K=4, 64 agents, random init (config 5; no shipped weights) ----
    torch.manual_seed(4)
    cfg5 = dict(BASE_CFG, num_agents=64, graph_filters=[4])
    net5 = fg.Network(cfg5)
    with torch.no_grad():                 # make BN running stats non-trivial but deterministic
        for mod in net5.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.uniform_(-0.2, 0.2)
                mod.running_var.uniform_(0.5, 1.5)
    net5.eval()
    S5, A5 = collect_observations(GraphEnv, rng, 1, 64, 64, 200, 6, 3)
    with torch.no_grad():
        l5 = net5(torch.tensor(S5).float(), torch.tensor(A5).float())
    np.savez_compressed(os.path.join(HERE, "weights_gnn_k4_seed4.npz"), **state_dict_to_npz(net5.state_dict()))
    model_out["gnn_k4_seed4/states"] = S5.astype(np.uint8)
    model_out["gnn_k4_seed4/gso"] = A5.astype(np.uint8)
    model_out["gnn_k4_seed4/logits"] = l5.numpy()
    np.savez_compressed(os.path.join(HERE, "model_logits.npz"), **model_out)

    # ---- one training step (train.py:82-100) from the shipped weights, train-mode BN ----
    train_out = {}
    rng = np.random.RandomState(11)
    S, A = collect_observations(GraphEnv, rng, 4, 5, 18, 5, 6, 6)          # B = 24
    targets = rng.randint(0, 5, size=S.shape[:2])
    train_out["states"] = S.astype(np.uint8)
    train_out["adj"] = A.astype(np.uint8)                                   # loader adds I (data_loader.py:74)
    train_out["targets"] = targets.astype(np.uint8)
    for name in ("gnn_msg_k3", "gnn_k3", "baseline"):
        cls, cfg, net0 = nets[name]
        net = cls(cfg)
        net.load_state_dict(copy.deepcopy(net0.state_dict()))
        net.train()
        opt = torch.optim.Adam(net.parameters(), lr=3e-4, weight_decay=1e-4)
        crit = torch.nn.CrossEntropyLoss()
        st = torch.tensor(S).float()
        gs = torch.tensor(A + np.eye(5)).float()
        tg = torch.tensor(targets).float()
        opt.zero_grad()
        out = net(st) if name == "baseline" else net(st, gs)
        loss = crit(out.reshape(-1, 5), tg.long().reshape(-1))
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
        total_norm = torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)
        opt.step()
        train_out[f"{name}/logits"] = out.detach().numpy()
        train_out[f"{name}/loss"] = np.array([loss.item()], dtype=np.float64)
        train_out[f"{name}/total_norm"] = np.array([float(total_norm)], dtype=np.float64)
        for k, g in grads.items():
            train_out[f"{name}/grad/{k}"] = g.numpy()
        for k, v in net.state_dict().items():
            train_out[f"{name}/after/{k}"] = v.detach().numpy()
        print(name, "train loss", loss.item(), "grad norm", float(total_norm))
    np.savez_compressed(os.path.join(HERE, "train_step.npz"), **train_out)

    # ---- graph layers alone, asymmetric adjacency with an isolated row ----
    from models.networks.gnn import GCNLayer, MessagePassingLayer
    torch.manual_seed(5)
    layer_out = {}
    for nm, cls, k, n in (("gcn_k3_n5", GCNLayer, 3, 5), ("msg_k3_n5", MessagePassingLayer, 3, 5),
                          ("gcn_k4_n20", GCNLayer, 4, 20), ("msg_k2_n7", MessagePassingLayer, 2, 7)):
        lay = cls(n, 64, 128, k)
        x = torch.randn(3, 64, n)
        adj = (torch.rand(3, n, n) < 0.4).float()
        adj[:, torch.arange(n), torch.arange(n)] = 0
        adj[0, 1, :] = 0                      # isolated row
        adj[1, :, 2] = 0                      # nobody listens to agent 2
        lay.addGSO(adj)
        with torch.no_grad():
            y = lay(x)
        layer_out[f"{nm}/x"] = x.numpy()
        layer_out[f"{nm}/adj"] = adj.numpy().astype(np.uint8)
        layer_out[f"{nm}/W"] = lay.W.detach().numpy()
        if hasattr(lay, "W2"):
            layer_out[f"{nm}/W2"] = lay.W2.detach().numpy()
        layer_out[f"{nm}/y"] = y.numpy()
    np.savez_compressed(os.path.join(HERE, "graph_layers.npz"), **layer_out)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
