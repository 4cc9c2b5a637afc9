"""Pin the model oracle (torch-CPU restatement) to logits / gradients produced by the reference."""
import numpy as np
import pytest
import torch

from conftest import load_weights
from oracle import model_oracle as mo

REL = 1e-4   # north-star tolerance for fp32 logits and gradients


GRAD_FLOOR = 1e-3   # x total grad norm.  Conv biases feeding train-mode BatchNorm have mathematically
                    # zero gradient: the reference values are ~1e-7 of rounding noise against a total
                    # norm of ~10, so a per-tensor relative error needs an absolute floor.


def rel_err(a, b, floor=1e-30):
    """max |a-b| / max(max |b|, floor) -- the per-tensor relative error of SURVEY section 8d."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), floor)


def as_params(name):
    return {k: torch.from_numpy(v) for k, v in load_weights(name).items()}


@pytest.mark.parametrize("name", ["gnn_k3", "gnn_k2", "gnn_msg_k3", "baseline", "gnn_k4_seed4"])
def test_logits_match_reference(model_logits, name):
    params = as_params(name)
    states = torch.from_numpy(model_logits[f"{name}/states"]).float()
    gso = torch.from_numpy(model_logits[f"{name}/gso"]).float()
    with torch.no_grad():
        logits = mo.forward(params, states, gso)
    ref = model_logits[f"{name}/logits"]
    assert logits.shape == ref.shape
    assert rel_err(logits.numpy(), ref) < 1e-5
    assert np.array_equal(mo.greedy_actions(logits), ref.argmax(-1))


@pytest.mark.parametrize("name", ["gcn_k3_n5", "msg_k3_n5", "gcn_k4_n20", "msg_k2_n7"])
def test_graph_layers_match_reference(graph_layers, name):
    g = graph_layers
    x = torch.from_numpy(g[f"{name}/x"])
    adj = torch.from_numpy(g[f"{name}/adj"]).float()
    W = torch.from_numpy(g[f"{name}/W"])
    W2 = torch.from_numpy(g[f"{name}/W2"]) if f"{name}/W2" in g.files else None
    y = mo.graph_filter(x, adj, W, W2)
    assert rel_err(y.numpy(), g[f"{name}/y"]) < 1e-6


@pytest.mark.parametrize("name", ["gnn_msg_k3", "gnn_k3", "baseline"])
def test_training_step_matches_reference(train_step, name):
    t = train_step
    params = as_params(name)
    states = torch.from_numpy(t["states"]).float()
    gso = torch.from_numpy(t["adj"]).float() + torch.eye(5)
    targets = torch.from_numpy(t["targets"]).float()
    loss, grads, bn_state = mo.loss_and_grads(params, states, gso, targets)
    assert abs(loss.item() - t[f"{name}/loss"][0]) < 1e-5 * abs(t[f"{name}/loss"][0])
    for k, g in grads.items():
        assert rel_err(g.numpy(), t[f"{name}/grad/{k}"], GRAD_FLOOR * t[f"{name}/total_norm"][0]) < REL, k
    for k, v in bn_state.items():
        ref = t[f"{name}/after/{k}"]
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(ref), k
        else:
            assert rel_err(v.detach().numpy(), ref) < 1e-5, k
    # Adam's first step is lr*g/(|g|+eps): it amplifies the rounding noise of the (mathematically zero)
    # conv-bias gradients, so the optimizer is checked on the reference's own gradients.
    ref_grads = {k: torch.from_numpy(t[f"{name}/grad/{k}"]) for k in grads}
    zeros = {k: torch.zeros_like(params[k]) for k in grads}
    new_p, _, _, total = mo.clip_and_adam(params, ref_grads, zeros, zeros, step=1)
    assert abs(float(total) - t[f"{name}/total_norm"][0]) < 1e-5 * t[f"{name}/total_norm"][0]
    for k in grads:
        assert rel_err(new_p[k].numpy(), t[f"{name}/after/{k}"]) < 1e-5, k
