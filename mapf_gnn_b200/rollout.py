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

    The whole sequence (1 host-to-device copy, 3-4 kernels, 1 device-to-host copy) is captured into CUDA graphs and
    replayed per step, so the host pays one graph launch and one stream synchronisation per step.  Two pinned result
    blocks alternate: the positions returned by step t are, in place, the input of step t+1 (the caller may edit them
    in between), so no host-side copy is needed to close the loop.
    ``pos_in`` int32[E,N,2] is the input of the next ``step()``; ``step()`` returns views ``(actions int32[E,N],
    pos_out int32[E,N,2], done uint8[E])`` of the block it just filled."""

    def __init__(self, rollout: Rollout):
        self.ro, env = rollout, rollout.env
        E, N = env.episodes, env.nb_agents
        nb_pos, nb_act = E * N * 8, E * N * 4
        self.blocks = [torch.empty((env.host_block.numel(),), dtype=torch.uint8).pin_memory() for _ in range(2)]
        view = lambda b: (b[nb_pos:nb_pos + nb_act].view(torch.int32).view(E, N),
                          b[:nb_pos].view(torch.int32).view(E, N, 2), b[nb_pos + nb_act:nb_pos + nb_act + E])
        self.views = [view(b) for b in self.blocks]
        self.h2d_bytes = nb_pos
        self.d2h_bytes = self.blocks[0].numel()
        self.cur = 0                                   # block that the next step() fills; its input is the other one
        snap = (env.pos.clone(), env.cells.clone(), env.time.clone(), env.done.clone(), env.fov.clone(),
                env.adj_matrix.clone())
        for b in self.blocks:
            b.copy_(env.host_block.cpu())
        side = torch.cuda.Stream(device=env.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for k in range(2):                         # warm-up: first-launch attributes, lazy module loading
                self._body(k)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for dst, src in zip((env.pos, env.cells, env.time, env.done, env.fov, env.adj_matrix), snap):
            dst.copy_(src)
        for b in self.blocks:
            b.copy_(env.host_block.cpu())
        self.graphs = []
        for k in range(2):
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                self._body(k)
            self.graphs.append(gph)

    @property
    def pos_in(self):
        """Pinned int32[E,N,2]: positions the next step() uploads (= the positions the previous step returned)."""
        return self.views[1 - self.cur][1]

    def _body(self, k):
        env, ro = self.ro.env, self.ro
        env.pos.copy_(self.views[1 - k][1], non_blocking=True)           # H2D: this step's state
        ro.run(1)
        self.blocks[k].copy_(env.host_block, non_blocking=True)           # D2H: actions + new positions + done

    def step(self):
        """One host-driven step; returns after the results are visible in pinned host memory."""
        k = self.cur
        self.graphs[k].replay()
        torch.cuda.current_stream().synchronize()
        self.cur = 1 - k
        return self.views[k]
