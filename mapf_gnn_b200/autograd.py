"""Autograd glue: differentiable entry points whose forward and backward are C-ABI kernel calls."""
from __future__ import annotations

import torch

from . import ops  # noqa: F401


class _GraphFilterFn(torch.autograd.Function):
    """GCNLayer / MessagePassingLayer forward (gnn.py:72-126,178-235) in the reference layout."""

    @staticmethod
    def forward(ctx, x, gso, W, W2, self_loop):
        wt = torch.ops.mapf.pack_graph_weights(W.detach(), None if W2 is None else W2.detach())
        y = torch.ops.mapf.graph_filter_fwd(x.detach(), gso.detach(), wt, W.shape[1], bool(self_loop), W2 is not None,
                                            False)
        ctx.save_for_backward(x, gso, W, W2)
        ctx.self_loop = bool(self_loop)
        return y

    @staticmethod
    def backward(ctx, gy):
        from .train import graph_filter_backward
        x, gso, W, W2 = ctx.saved_tensors
        gx, gW, gW2 = graph_filter_backward(x, gso, W, W2, ctx.self_loop, gy.contiguous())
        return gx, None, gW, gW2, None


def graph_filter(x, gso, W, W2, self_loop):
    return _GraphFilterFn.apply(x, gso, W, W2, self_loop)


def policy_train_forward(net, states, gso):
    from .train import policy_train_forward as impl
    return impl(net, states, gso)
