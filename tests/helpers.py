"""Shared helpers for the parity tests (host-side only; no CUDA here)."""
import numpy as np
import torch

from conftest import load_weights

BASE_CFG = {
    "device": "cuda", "map_shape": [5, 5], "num_actions": 5, "last_convs": [400], "channels": [16, 16, 16],
    "node_dims": [128], "graph_filters": [3], "encoder_dims": [64], "encoder_layers": 1, "action_layers": 1,
    "min_time": 1, "max_time": 32, "num_agents": 5, "board_size": [18, 18],
}

MODEL_CFG = {
    "gnn_k3": dict(net_type="gnn", msg_type="gcn", graph_filters=[3]),
    "gnn_k2": dict(net_type="gnn", msg_type="gcn", graph_filters=[2]),
    "gnn_msg_k3": dict(net_type="gnn", msg_type="message", graph_filters=[3]),
    "baseline": dict(net_type="baseline", channels=[32, 16, 16]),
    "gnn_k4_seed4": dict(net_type="gnn", msg_type="gcn", graph_filters=[4]),
}


def model_config(name, num_agents=5, **over):
    return dict(BASE_CFG, num_agents=num_agents, **MODEL_CFG[name], **over)


def load_network(name, num_agents=5, device="cuda"):
    import mapf_gnn_b200 as m
    net = m.build_network(model_config(name, num_agents))
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in load_weights(name).items()}
    net.load_state_dict(sd)
    return net.to(device)


def oracle_params(name):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in load_weights(name).items()}


def rel_err(a, b, floor=1e-30):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), floor)


def top2_margin(logits):
    s = np.sort(np.asarray(logits, dtype=np.float64), axis=-1)
    return s[..., -1] - s[..., -2]
