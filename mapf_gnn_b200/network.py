"""Policy networks with the reference's module API and ``state_dict`` layout.

``GNNNetwork``      = models/framework_gnn.py ``Network`` (:21-181)          CNN -> MLP -> GCNLayer -> MLP
``MessageNetwork``  = models/framework_gnn_message.py ``Network`` (:8-162)   same with MessagePassingLayer
``BaselineNetwork`` = models/framework_baseline.py ``Network`` (:109-129)    CNN -> MLP -> Linear, no graph

The torch sub-modules (``convLayers``, ``compressMLP``, ``GFL``, ``actionMLP``) only *hold* the
parameters under the reference's names so that the shipped ``trained_models/*/model.pt`` load with
``load_state_dict``; ``forward`` never calls them.  Eval mode runs the fused inference kernels on
BatchNorm-folded packed weights; train mode runs the training kernels (per-agent-slice BatchNorm
statistics) under a custom autograd function.
"""
from __future__ import annotations

from typing import Any, Dict

import torch
import torch.nn as nn

from . import ops  # noqa: F401
from .layers import GCNLayer, MessagePassingLayer


def weights_init(m):
    """models/networks/utils_weights.py:10-22."""
    classname = m.__class__.__name__
    if classname.find("Conv") != -1:
        torch.nn.init.xavier_normal_(m.weight)
    elif classname.find("BatchNorm") != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0.0)
    elif classname.find("Linear") != -1:
        torch.nn.init.xavier_normal_(m.weight)
        m.bias.data.fill_(0.0)


class _PolicyBase(nn.Module):
    graph_layer = None          # GCNLayer / MessagePassingLayer / None
    message = False

    def __init__(self, config: Dict[str, Any]):
        super().__init__()
        self.config = config
        self.S = None
        self.num_agents = config["num_agents"]
        self.map_shape = config["map_shape"]
        self.num_actions = 5
        if list(self.map_shape) != [5, 5]:
            raise NotImplementedError("the kernels are built for the 5x5 FOV of env_graph_gridv1.py:252")
        channels = [2] + list(config["channels"])
        conv_layers = []
        for l in range(len(channels) - 1):
            conv_layers += [nn.Conv2d(channels[l], channels[l + 1], 3, 1, 1, bias=True),
                            nn.BatchNorm2d(channels[l + 1]), nn.ReLU(inplace=True)]
        self.convLayers = nn.Sequential(*conv_layers)
        self.channels = channels
        if config["encoder_layers"] != 1:
            raise NotImplementedError("encoder_layers != 1 (no shipped config uses it)")
        self.compress_Features_dim = list(config["last_convs"]) + list(config["encoder_dims"])
        if self.compress_Features_dim[0] != 25 * channels[-1]:
            raise ValueError("last_convs must equal 25 * channels[-1]")
        self.compressMLP = nn.Sequential(nn.Linear(self.compress_Features_dim[0], self.compress_Features_dim[1]),
                                         nn.ReLU(inplace=True))
        feat = self.compress_Features_dim[-1]
        if self.graph_layer is not None:
            self.graph_filter = list(config["graph_filters"])
            self.node_dim = list(config["node_dims"])
            if len(self.graph_filter) != 1:
                raise NotImplementedError("more than one graph filtering layer (no shipped config uses it)")
            self.features = [feat] + self.node_dim
            self.GFL = nn.Sequential(self.graph_layer(self.num_agents, self.features[0], self.features[1],
                                                      self.graph_filter[0], activation=None),
                                     nn.ReLU(inplace=True))
            head_in = self.features[-1]
        else:
            head_in = feat
        if config.get("action_layers", 1) != 1:
            raise NotImplementedError("action_layers != 1 (no shipped config uses it)")
        self.actionMLP = nn.Sequential(nn.Linear(head_in, self.num_actions))
        self.apply(weights_init)
        self._packed = None
        self._packed_key = None

    # -- flat description shared with the C ABI ---------------------------------------------------
    def dims(self):
        ch = self.channels + [0] * (5 - len(self.channels))
        taps = self.graph_filter[0] if self.graph_layer is not None else 0
        node = self.node_dim[0] if self.graph_layer is not None else 0
        return [len(self.channels) - 1, *ch, self.compress_Features_dim[0], self.compress_Features_dim[1], taps,
                node, self.num_actions]

    def param_groups(self):
        convs = [m for m in self.convLayers if isinstance(m, nn.Conv2d)]
        bns = [m for m in self.convLayers if isinstance(m, nn.BatchNorm2d)]
        gf = self.GFL[0] if self.graph_layer is not None else None
        return {
            "conv_w": [m.weight for m in convs], "conv_b": [m.bias for m in convs],
            "bn_gamma": [m.weight for m in bns], "bn_beta": [m.bias for m in bns],
            "bn_mean": [m.running_mean for m in bns], "bn_var": [m.running_var for m in bns],
            "fc_w": self.compressMLP[0].weight, "fc_b": self.compressMLP[0].bias,
            "gf_w": None if gf is None else gf.W, "gf_w2": getattr(gf, "W2", None) if gf is not None else None,
            "act_w": self.actionMLP[0].weight, "act_b": self.actionMLP[0].bias,
        }

    def packed_weights(self):
        """BatchNorm-folded, kernel-ordered weights; re-packed when any parameter/buffer changed.  The buffer keeps its
        ADDRESS across re-packs (new weights are copied in place), so launch sequences that captured its pointer in a
        CUDA graph (HostSteppedRollout) read the current weights after the next call of this method instead of freed
        memory."""
        g = self.param_groups()
        flat = [t for v in g.values() for t in (v if isinstance(v, list) else [v]) if t is not None]
        key = (getattr(self, "_weights_epoch", 0),) + tuple((t.data_ptr(), t._version) for t in flat)
        if self._packed is None or key != self._packed_key:
            d = lambda ts: [t.detach() for t in ts]
            fresh = torch.ops.mapf.pack_weights(
                self.dims(), d(g["conv_w"]), d(g["conv_b"]), d(g["bn_gamma"]), d(g["bn_beta"]), d(g["bn_mean"]),
                d(g["bn_var"]), g["fc_w"].detach(), g["fc_b"].detach(),
                None if g["gf_w"] is None else g["gf_w"].detach(),
                None if g["gf_w2"] is None else g["gf_w2"].detach(), g["act_w"].detach(), g["act_b"].detach())
            if self._packed is not None and self._packed.shape == fresh.shape and self._packed.device == fresh.device:
                self._packed.copy_(fresh)
            else:
                self._packed = fresh
            self._packed_key = key
            self._packed_serial = getattr(self, "_packed_serial", 0) + 1
        return self._packed

    def _forward(self, states, gso):
        if self.training:
            from .autograd import policy_train_forward
            return policy_train_forward(self, states, gso)
        logits, _ = torch.ops.mapf.policy_fwd(self.dims(), self.message, states, gso, self.packed_weights())
        return logits

    @torch.no_grad()
    def act(self, states, gso=None):
        """Eval-mode logits and greedy actions (first-max argmax, evaluate.py:238) in one launch sequence."""
        return torch.ops.mapf.policy_fwd(self.dims(), self.message, states, gso, self.packed_weights())


class GNNNetwork(_PolicyBase):
    graph_layer = GCNLayer
    message = False

    def forward(self, states: torch.Tensor, gso: torch.Tensor) -> torch.Tensor:
        return self._forward(states, gso)


class MessageNetwork(_PolicyBase):
    graph_layer = MessagePassingLayer
    message = True

    def forward(self, states: torch.Tensor, gso: torch.Tensor) -> torch.Tensor:
        return self._forward(states, gso)


class BaselineNetwork(_PolicyBase):
    graph_layer = None
    message = False

    def forward(self, states: torch.Tensor) -> torch.Tensor:
        return self._forward(states, None)
