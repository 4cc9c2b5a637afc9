"""Graph layers with the reference's class API (models/networks/gnn.py): ``GCNLayer`` (:12-126) and
``MessagePassingLayer`` (:129-235).  Same constructor, parameter names/shapes (``W [F,K,G]``,
``W2 [F,G]``), init (:54-58,159-165), the stateful ``addGSO`` setter and ``forward(X[B,F,N]) ->
[B,G,N]``; the arithmetic runs in the fused sm_100a graph-filter kernel."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops  # noqa: F401  (registers torch.ops.mapf)


class _GraphFilterBase(nn.Module):
    _self_loop = True
    _skip = False

    def __init__(self, n_nodes, in_features, out_features, filter_number, bias=False, activation=None,
                 name="GCN_Layer"):
        super().__init__()
        if bias:
            raise NotImplementedError("bias=True is never used by the reference networks (framework_gnn.py:107-115)")
        self.n_nodes = n_nodes
        self.in_features = in_features
        self.out_features = out_features
        self.filter_number = filter_number
        self.W = nn.Parameter(torch.empty(in_features, filter_number, out_features))
        if self._skip:
            self.W2 = nn.Parameter(torch.empty(in_features, out_features))
        self.register_parameter("bias", None)
        self.b = None
        self.activation = activation
        self.name = name
        self.adj_matrix = None
        self.init_params()

    def init_params(self):
        stdv = 1.0 / math.sqrt(self.in_features * self.filter_number)
        self.W.data.uniform_(-stdv, stdv)
        if self._skip:
            stdv = 1.0 / math.sqrt(self.in_features)
            self.W2.data.uniform_(-stdv, stdv)

    def extra_repr(self):
        return (f"in_features={self.in_features}, out_features={self.out_features}, "
                f"filter_taps={self.filter_number}, bias=False")

    def addGSO(self, GSO):
        """Set the graph shift operator (adjacency matrix) used by the next forward (gnn.py:68-70)."""
        self.adj_matrix = GSO

    def forward(self, node_feats):
        if self.adj_matrix is None:
            raise RuntimeError("addGSO() must be called before forward()")
        from .autograd import graph_filter
        y = graph_filter(node_feats, self.adj_matrix.to(node_feats.dtype), self.W,
                         self.W2 if self._skip else None, self._self_loop)
        if self.activation is not None:
            y = self.activation(y)
        return y


class GCNLayer(_GraphFilterBase):
    """sum_k S^k X H_k with S = D^-1/2 (A + I) D^-1/2 (gnn.py:72-126)."""
    _self_loop = True
    _skip = False


class MessagePassingLayer(_GraphFilterBase):
    """Same filter without the self loop plus the skip weight W2 (gnn.py:178-235)."""
    _self_loop = False
    _skip = True

    def __init__(self, n_nodes, in_features, out_features, filter_number, bias=False, activation=None,
                 name="GCN_Layer"):
        super().__init__(n_nodes, in_features, out_features, filter_number, bias, activation, name)
