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
        # encoder MLP: `encoder_layers` x (Linear + ReLU) over last_convs + encoder_dims (framework_gnn.py:84-99)
        enc_layers = int(config["encoder_layers"])
        self.compress_Features_dim = list(config["last_convs"]) + list(config["encoder_dims"])
        if self.compress_Features_dim[0] != 25 * channels[-1]:
            raise ValueError("last_convs must equal 25 * channels[-1]")
        if enc_layers < 1 or enc_layers > len(self.compress_Features_dim) - 1:
            raise IndexError("encoder_layers needs that many entries in encoder_dims (framework_gnn.py:90-96)")
        mlp = []
        for l in range(enc_layers):
            mlp += [nn.Linear(self.compress_Features_dim[l], self.compress_Features_dim[l + 1]), nn.ReLU(inplace=True)]
        self.compressMLP = nn.Sequential(*mlp)
        feat = self.compress_Features_dim[-1]
        if self.graph_layer is not None:
            # `len(graph_filters)` x (graph layer + ReLU) over [encoder output] + node_dims (framework_gnn.py:105-123)
            self.graph_filter = list(config["graph_filters"])
            self.node_dim = list(config["node_dims"])
            if len(self.node_dim) < len(self.graph_filter) or not self.graph_filter:
                raise IndexError("node_dims needs one entry per graph filtering layer (framework_gnn.py:107-118)")
            self.features = [feat] + self.node_dim
            gfl = []
            for l in range(len(self.graph_filter)):
                gfl += [self.graph_layer(self.num_agents, self.features[l], self.features[l + 1], self.graph_filter[l],
                                         activation=None), nn.ReLU(inplace=True)]
            self.GFL = nn.Sequential(*gfl)
            head_in = self.features[-1]
        else:
            head_in = feat
        if config.get("action_layers", 1) != 1:
            # the reference itself cannot build this: action_features = [head_in, 5] has no entry for a second layer
            # (framework_gnn.py:38,125-138 raises IndexError)
            raise IndexError("action_layers != 1: the reference's action_features list has a single layer "
                             "(framework_gnn.py:38,130-138)")
        self.actionMLP = nn.Sequential(nn.Linear(head_in, self.num_actions))
        self.apply(weights_init)
        self._packed = None
        self._packed_key = None
        # every shipped config is 3 conv + 1 FC (+ 1 graph layer) + 1 Linear: that shape runs in the fused kernels
        # (inference and training); deeper encoder MLPs / several graph layers run stage by stage (inference only)
        self.fused = enc_layers == 1 and (self.graph_layer is None or len(self.graph_filter) == 1)

    # -- flat description shared with the C ABI ---------------------------------------------------
    def dims(self):
        """Fused shape; for the staged path: the first stage (conv stack + first encoder Linear) as a headless net."""
        ch = self.channels + [0] * (5 - len(self.channels))
        taps = self.graph_filter[0] if self.graph_layer is not None and self.fused else 0
        node = self.node_dim[0] if self.graph_layer is not None and self.fused else 0
        return [len(self.channels) - 1, *ch, self.compress_Features_dim[0], self.compress_Features_dim[1], taps,
                node, self.num_actions]

    def param_groups(self):
        convs = [m for m in self.convLayers if isinstance(m, nn.Conv2d)]
        bns = [m for m in self.convLayers if isinstance(m, nn.BatchNorm2d)]
        gf = self.GFL[0] if self.graph_layer is not None and self.fused else None
        if not self.fused:      # staged path: stage 1 = conv stack + first encoder Linear; the head slot is unused there
            dummy_w = getattr(self, "_stage1_head_w", None)
            if dummy_w is None or dummy_w.device != self.compressMLP[0].weight.device:
                dev = self.compressMLP[0].weight.device
                self._stage1_head_w = dummy_w = torch.zeros((self.num_actions, self.compress_Features_dim[1]), device=dev)
                self._stage1_head_b = torch.zeros((self.num_actions,), device=dev)
            return {
                "conv_w": [m.weight for m in convs], "conv_b": [m.bias for m in convs],
                "bn_gamma": [m.weight for m in bns], "bn_beta": [m.bias for m in bns],
                "bn_mean": [m.running_mean for m in bns], "bn_var": [m.running_var for m in bns],
                "fc_w": self.compressMLP[0].weight, "fc_b": self.compressMLP[0].bias, "gf_w": None, "gf_w2": None,
                "act_w": dummy_w, "act_b": self._stage1_head_b,
            }
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
        if not self.fused:
            return self._staged(states, gso)[0]
        if self.training:
            from .autograd import policy_train_forward
            return policy_train_forward(self, states, gso)
        logits, _ = torch.ops.mapf.policy_fwd(self.dims(), self.message, states, gso, self.packed_weights())
        return logits

    @torch.no_grad()
    def act(self, states, gso=None):
        """Eval-mode logits and greedy actions (first-max argmax, evaluate.py:238) in one launch sequence."""
        if not self.fused:
            return self._staged(states, gso)
        return torch.ops.mapf.policy_fwd(self.dims(), self.message, states, gso, self.packed_weights())

    def _staged(self, states, gso):
        """General configs (encoder_layers > 1, several graph layers; framework_gnn.py:90-123 loops), inference only,
        stage by stage on the same kernels: conv stack + first Linear (mapf::encoder_fwd) -> further Linear + ReLU
        (mapf::linear, tcgen05 GEMM) -> graph layers (mapf::graph_filter_fwd, [B,F,N] layout) + ReLU -> Linear + argmax
        (mapf::head_fwd)."""
        if self.training:
            raise NotImplementedError("training kernels exist for the shipped network shape (1 encoder Linear, <= 1 "
                                      "graph layer); call eval() for deeper configurations")
        B, N = int(states.shape[0]), int(states.shape[1])
        x = torch.ops.mapf.encoder_fwd(self.dims(), states.contiguous(), self.packed_weights())       # [B*N, F1]
        for l in range(2, len(self.compressMLP), 2):
            lin = self.compressMLP[l]
            x = torch.ops.mapf.linear(x, lin.weight.detach(), lin.bias.detach(), True)
        if self.graph_layer is not None:
            gso = gso.to(torch.float32).contiguous()
            xb = x.view(B, N, -1).permute(0, 2, 1).contiguous()                                        # [B,F,N]
            for l in range(0, len(self.GFL), 2):
                layer = self.GFL[l]
                wt = torch.ops.mapf.pack_graph_weights(layer.W.detach(), layer.W2.detach() if self.message else None)
                xb = torch.ops.mapf.graph_filter_fwd(xb, gso, wt, layer.filter_number, not self.message, self.message,
                                                     True)
            x = xb.permute(0, 2, 1).contiguous().view(B * N, -1)
        head = self.actionMLP[0]
        logits, actions = torch.ops.mapf.head_fwd(x, head.weight.detach(), head.bias.detach())
        return logits.view(B, N, -1), actions.view(B, N)


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
