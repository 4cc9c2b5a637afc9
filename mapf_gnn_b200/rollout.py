"""On-device rollout of the evaluate.py:223-241 protocol for E resident episodes, and the model
factory that mirrors evaluate.py:161-199 / train.py:36-41."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .env import BatchedGraphEnv
from .network import BaselineNetwork, GNNNetwork, MessageNetwork
from .ops import net_params


def build_network(config):
    """net_type 'baseline' -> BaselineNetwork; 'gnn' -> MessageNetwork if msg_type == 'message' else
    GNNNetwork (evaluate.py:163-173)."""
    if config.get("net_type", "gnn") == "baseline":
        return BaselineNetwork(config)
    if config.get("msg_type", "gcn") == "message":
        return MessageNetwork(config)
    return GNNNetwork(config)


class Rollout:
    """Binds one network (eval mode) and one ``BatchedGraphEnv``; ``run(steps)`` enqueues
    steps x (policy forward -> argmax -> env step) on the current stream with no host round trip."""

    def __init__(self, net, env: BatchedGraphEnv):
        self.net, self.env = net, env
        self.lib = _lib.load()
        E, N = env.episodes, env.nb_agents
        dims = net.dims()
        dev = env.device
        self.actions = torch.zeros((E, N), dtype=torch.int32, device=dev)
        self.logits = torch.empty((E, N, dims[10]), dtype=torch.float32, device=dev)
        self.work = torch.empty((E * N, dims[7]), dtype=torch.float32, device=dev)
        from . import ops
        self.worka = (torch.empty((E * N, dims[6]), dtype=torch.float32, device=dev) if ops.USE_TENSOR_CORES and ops.USE_TENSOR_CORE_FC else None)
        self.workz = (torch.empty((E * N, (dims[8] + 1) * dims[7]), dtype=torch.float32, device=dev)
                      if dims[8] > 0 and ops.USE_TENSOR_CORES else None)
        self.params = net_params(dims)
        self.is_graph = dims[8] > 0

    def _args(self, steps):
        env, a = self.env, _lib.RolloutArgs()
        a.env = env.step_args(self.actions, 1)
        p = a.policy
        p.batch, p.agents, p.message = env.episodes, env.nb_agents, int(self.net.message)
        p.states, p.gso = env.fov.data_ptr(), env.adj_matrix.data_ptr()
        p.packed, p.workspace_x = self.net.packed_weights().data_ptr(), self.work.data_ptr()
        p.workspace_z = None if self.workz is None else self.workz.data_ptr()
        p.workspace_act = None if self.worka is None else self.worka.data_ptr()
        p.logits, p.actions = self.logits.data_ptr(), self.actions.data_ptr()
        a.steps = int(steps)
        return a

    @torch.no_grad()
    def run(self, steps):
        if self.net.training:
            raise RuntimeError("Rollout needs the network in eval() mode (evaluate.py:198)")
        args = self._args(steps)
        with torch.cuda.device(self.env.device):
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(self.lib.mapf_rollout(self.params, args, stream), "rollout")
        return self.env.done
