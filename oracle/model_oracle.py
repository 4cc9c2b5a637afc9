"""Oracle (test infrastructure) -- torch-CPU fp32 restatement of the reference policy nets.

Follows ``models/framework_gnn.py:143-181`` (GCN variant), ``models/framework_gnn_message.py``
(message-passing variant), ``models/framework_baseline.py:109-129`` and the two graph layers
``models/networks/gnn.py:72-126`` / ``:178-235``; the training step follows ``train.py:82-100``.
Parameters are addressed by the reference ``state_dict`` key layout so the
shipped weights and goldens drop straight in.

The arithmetic is ATen's (torch CPU), as in the reference; this file restates the data flow,
not the library.  Model numerics are unpinned by the reference's own tests; they are pinned by
goldens generated from the unmodified reference (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # nn.BatchNorm2d default, framework_gnn.py:70
BN_MOMENTUM = 0.1


def conv_layer_ids(params):
    """Indices of the Conv2d modules inside ``convLayers`` (0, 3, 6, ... framework_gnn.py:55-79)."""
    ids = sorted({int(k.split(".")[1]) for k in params if k.startswith("convLayers.") and k.endswith(".weight")
                  and params[k].dim() == 4})
    return ids


def encode_agents(params, states, training=False, bn_state=None):
    """A5.  states f32[B,N,2,5,5] -> X f32[B,F,N] (framework_gnn.py:156-166).

    The conv stack is applied once per agent slice, so train-mode BatchNorm statistics are per
    agent index over B*25 values and the running stats take N sequential momentum updates.
    ``bn_state`` (dict of running_mean/var/num_batches_tracked clones) is updated in place when
    training.
    """
    B, N = states.shape[:2]
    cols = []
    for a in range(N):
        h = states[:, a]
        for li in conv_layer_ids(params):
            h = F.conv2d(h, params[f"convLayers.{li}.weight"], params[f"convLayers.{li}.bias"], stride=1, padding=1)
            bn = f"convLayers.{li + 1}"
            if training:
                rm = bn_state[f"{bn}.running_mean"] if bn_state is not None else None
                rv = bn_state[f"{bn}.running_var"] if bn_state is not None else None
                h = F.batch_norm(h, rm, rv, params[f"{bn}.weight"], params[f"{bn}.bias"], True, BN_MOMENTUM, BN_EPS)
                if bn_state is not None:
                    bn_state[f"{bn}.num_batches_tracked"] += 1
            else:
                h = F.batch_norm(h, params[f"{bn}.running_mean"], params[f"{bn}.running_var"],
                                 params[f"{bn}.weight"], params[f"{bn}.bias"], False, BN_MOMENTUM, BN_EPS)
            h = F.relu(h)
        h = h.reshape(B, -1)                                    # (C,H,W) flatten, framework_gnn.py:163
        li = 0
        while f"compressMLP.{li}.weight" in params:
            h = F.relu(F.linear(h, params[f"compressMLP.{li}.weight"], params[f"compressMLP.{li}.bias"]))
            li += 2
        cols.append(h)
    return torch.stack(cols, dim=2)                              # [B,F,N]


def normalise_gso(adj, add_self_loop):
    """S = D^-1/2 (A [+ I]) D^-1/2 with D from row sums, inf -> 0 (gnn.py:87-101 / :185-199)."""
    if add_self_loop:
        adj = adj + torch.eye(adj.shape[-1], dtype=adj.dtype)
    deg = adj.sum(dim=2)
    dinv = torch.pow(deg, -0.5)
    dinv[dinv == torch.inf] = 0
    return (torch.diag_embed(dinv) @ adj) @ torch.diag_embed(dinv)


def graph_filter(x, adj, W, W2=None):
    """A6 (W2 None, self loop added) / A7 (W2 given, no self loop).  x f32[B,F,N] -> [B,G,N].

    x_k = x_{k-1} @ S; Z = concat over k on the feature axis, viewed [B,N,K*F]; W [F,K,G] is
    consumed through a raw reshape to [G,K*F] (gnn.py:104-118); the message variant adds
    x.reshape(B,N,F) @ W2.reshape(G,F)^T, both raw reinterpretations (gnn.py:201-227).
    """
    B, Fin, N = x.shape
    K, G = W.shape[1], W.shape[2]
    S = normalise_gso(adj, add_self_loop=W2 is None)
    taps = [x]
    for _ in range(1, K):
        taps.append(taps[-1] @ S)
    z = torch.stack(taps, dim=1)                                  # [B,K,F,N]
    z = z.permute(0, 3, 1, 2).reshape(B, N, K * Fin)
    y = z @ W.reshape(G, K * Fin).t()                             # [B,N,G]
    if W2 is not None:
        y = y + x.reshape(B, N, Fin) @ W2.reshape(G, Fin).t()
    return y.permute(0, 2, 1)


def action_head(params, y):
    """A8.  y f32[B,D,N] -> logits f32[B,N,5] (framework_gnn.py:173-181)."""
    h = y.permute(0, 2, 1)
    li = 0
    while f"actionMLP.{li}.weight" in params:
        h = F.linear(h, params[f"actionMLP.{li}.weight"], params[f"actionMLP.{li}.bias"])
        if f"actionMLP.{li + 2}.weight" in params:
            h = F.relu(h)
        li += 2
    return h


def forward(params, states, gso=None, training=False, bn_state=None):
    """Whole net.  The variant is read off the parameter names: ``GFL.0.W2`` -> message passing,
    ``GFL.0.W`` -> GCN, neither -> no-communication baseline."""
    x = encode_agents(params, states, training, bn_state)
    li = 0
    while f"GFL.{li}.W" in params:
        x = F.relu(graph_filter(x, gso, params[f"GFL.{li}.W"], params.get(f"GFL.{li}.W2")))
        li += 2
    return action_head(params, x)


def greedy_actions(logits):
    """First-max argmax over the 5 actions (np.argmax tie-break, evaluate.py:238)."""
    return logits.detach().numpy().argmax(axis=-1)


def trainable_keys(params):
    return [k for k in params if not (k.endswith("running_mean") or k.endswith("running_var")
                                      or k.endswith("num_batches_tracked"))]


def loss_and_grads(params, states, gso, targets):
    """train.py:87-96: train-mode forward, mean CE over B*N rows, backward.  Returns
    (loss, {key: grad}, new bn_state)."""
    keys = trainable_keys(params)
    leaf = {k: (params[k].clone().requires_grad_(True) if k in keys else params[k].clone()) for k in params}
    bn_state = {k: leaf[k] for k in leaf if k not in keys}
    logits = forward(leaf, states, gso, training=True, bn_state=bn_state)
    loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), targets.long().reshape(-1))
    loss.backward()
    return loss.detach(), {k: leaf[k].grad for k in keys}, bn_state


def clip_and_adam(params, grads, m, v, step, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4,
                  max_norm=1.0):
    """train.py:97,100: clip_grad_norm_(1.0) then Adam with L2 weight decay (torch.optim.Adam
    defaults).  ``step`` is the 1-based step count.  Returns (new params, m, v, total_norm)."""
    keys = list(grads)
    total = torch.sqrt(sum((grads[k].double() ** 2).sum() for k in keys)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    out_p, out_m, out_v = {}, {}, {}
    for k in keys:
        g = grads[k] * coef + weight_decay * params[k]
        out_m[k] = betas[0] * m[k] + (1 - betas[0]) * g
        out_v[k] = betas[1] * v[k] + (1 - betas[1]) * g * g
        mhat = out_m[k] / (1 - betas[0] ** step)
        denom = (out_v[k].sqrt() / (1 - betas[1] ** step) ** 0.5) + eps
        out_p[k] = params[k] - lr * mhat / denom
    return out_p, out_m, out_v, total
