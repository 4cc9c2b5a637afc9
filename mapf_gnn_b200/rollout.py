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
        self.actions = env.actions          # lives next to env.pos / env.done (one block for host-driven callers)
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


class HostSteppedRollout:
    """The step protocol of evaluate.py:223-241 driven from HOST buffers, the way an external caller holding numpy
    state would use it: per step the positions go up from pinned memory, the policy + environment step runs, and the
    chosen actions, the new positions and the done flags come back to pinned memory.

    The whole sequence (1 host-to-device copy, 4 kernels, 1 device-to-host copy) is captured once into a CUDA graph
    and replayed per step, so the host pays one graph launch and one stream synchronisation per step.
    Buffers (pinned): ``pos_in`` int32[E,N,2] (written by the caller), ``actions`` int32[E,N], ``pos_out``
    int32[E,N,2], ``done`` uint8[E] (read by the caller after ``step()`` returns)."""

    def __init__(self, rollout: Rollout):
        self.ro, env = rollout, rollout.env
        E, N = env.episodes, env.nb_agents
        nb_pos, nb_act = E * N * 8, E * N * 4
        self.pos_in = torch.empty((E, N, 2), dtype=torch.int32).pin_memory()
        self.out_block = torch.empty((env.host_block.numel(),), dtype=torch.uint8).pin_memory()
        self.pos_out = self.out_block[:nb_pos].view(torch.int32).view(E, N, 2)
        self.actions = self.out_block[nb_pos:nb_pos + nb_act].view(torch.int32).view(E, N)
        self.done = self.out_block[nb_pos + nb_act:nb_pos + nb_act + E]
        self.pos_in.copy_(env.pos.cpu())
        self.h2d_bytes = self.pos_in.numel() * 4
        self.d2h_bytes = self.out_block.numel()
        snap = (env.pos.clone(), env.cells.clone(), env.time.clone(), env.done.clone(), env.fov.clone(),
                env.adj_matrix.clone())
        side = torch.cuda.Stream(device=env.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):                         # warm-up: first-launch attributes, lazy module loading
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for dst, src in zip((env.pos, env.cells, env.time, env.done, env.fov, env.adj_matrix), snap):
            dst.copy_(src)
        self.pos_in.copy_(env.pos.cpu())
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()

    def _body(self):
        env, ro = self.ro.env, self.ro
        env.pos.copy_(self.pos_in, non_blocking=True)
        ro.run(1)
        self.out_block.copy_(env.host_block, non_blocking=True)      # actions + new positions + done in one copy

    def step(self):
        """One host-driven step; returns after the results are visible in the pinned output buffers."""
        self.graph.replay()
        torch.cuda.current_stream().synchronize()
        return self.actions, self.pos_out, self.done
