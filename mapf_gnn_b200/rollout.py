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
        self.worka = (torch.empty((self.lib.mapf_tc_presplit_a_size(E * N, dims[6]),), dtype=torch.float32, device=dev)
                      if ops.wants_encoder_workspace(net_params(dims)) else None)
        self.workz = (torch.empty((E * N, (dims[8] + 1) * dims[7]), dtype=torch.float32, device=dev)
                      if dims[8] > 0 and ops.USE_TENSOR_CORES else None)
        self.params = net_params(dims)
        self.is_graph = dims[8] > 0
        # host-resident state (HostSteppedRollout(zero_copy=True)): addresses in mapped pinned host memory
        # (positions in, done flags in, positions out, actions out, done flags out) that the kernels read and write
        # directly instead of env.pos / actions / done
        self.host_state = None

    def _args(self, steps):
        env, a = self.env, _lib.RolloutArgs()
        a.env = env.step_args(self.actions, 1)
        p = a.policy
        p.batch, p.agents, p.message = env.episodes, env.nb_agents, int(self.net.message)
        p.states, p.gso = env.fov.data_ptr(), env.adj_matrix.data_ptr()
        self._packed = self.net.packed_weights()     # keep the buffer alive for as long as launches may read it
        p.packed, p.workspace_x = self._packed.data_ptr(), self.work.data_ptr()
        p.workspace_z = None if self.workz is None else self.workz.data_ptr()
        p.workspace_act = None if self.worka is None else self.worka.data_ptr()
        p.logits, p.actions = self.logits.data_ptr(), self.actions.data_ptr()
        if self.host_state is not None:
            pos_in, done_in, pos_out, act_out, done_out = self.host_state
            a.env.pos, a.env.done, a.env.pos_out, a.env.done_out = pos_in, done_in, pos_out, done_out
            a.env.actions = p.actions = act_out
        from . import ops
        p.flags = ops.policy_flags()
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

    def __init__(self, rollout: Rollout, zero_copy=False):
        self.ro, env = rollout, rollout.env
        E, N = env.episodes, env.nb_agents
        nb_pos, nb_act = E * N * 8, E * N * 4
        # zero_copy: the kernels themselves read the input block (positions, done flags) and write the output block
        # (actions, new positions, done flags) over the bus -- no copy nodes in the step's graph; the two pinned blocks
        # ARE the state (mapf_env_step_args.pos_out / done_out), alternating like with the copies.
        self.zero_copy = bool(zero_copy)
        self.blocks = [torch.empty((env.host_block.numel(),), dtype=torch.uint8).pin_memory() for _ in range(2)]
        view = lambda b: (b[nb_pos:nb_pos + nb_act].view(torch.int32).view(E, N),
                          b[:nb_pos].view(torch.int32).view(E, N, 2), b[nb_pos + nb_act:nb_pos + nb_act + E])
        self.views = [view(b) for b in self.blocks]
        self.views_np = [tuple(t.numpy() for t in v) for v in self.views]     # the same pinned memory as numpy arrays
        self.h2d_bytes = nb_pos + (2 * E if self.zero_copy else 0)      # zero copy: + the done flags, read twice
        self.d2h_bytes = self.blocks[0].numel()
        self._nb = (nb_pos, nb_act)
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
        self.stream = torch.cuda.Stream(device=env.device)     # replays run here, so several rollouts can overlap
        self.event = torch.cuda.Event()
        with torch.cuda.stream(self.stream):
            self.event.record()                                 # materialises the cudaEvent_t handle
        self.stream.wait_stream(torch.cuda.current_stream())
        # raw handles: launch / wait are two C calls each (mapf_graph_launch, mapf_event_wait), not torch wrappers
        self._exec = [C.c_void_p(g.raw_cuda_graph_exec()) for g in self.graphs]
        self._stream = C.c_void_p(self.stream.cuda_stream)
        self._event = C.c_void_p(self.event.cuda_event)
        self._lib = _lib.load()
        self.in_flight = False
        self._packed = rollout._packed               # the captured graphs read this buffer: same address for its lifetime

    @property
    def pos_in(self):
        """Pinned int32[E,N,2]: positions the next step() uploads (= the positions the previous step returned)."""
        return self.views[1 - self.cur][1]

    def _body(self, k):
        env, ro = self.ro.env, self.ro
        if self.zero_copy:
            nb_pos, nb_act = self._nb      # (unified addressing: the pinned host addresses are valid on the device)
            src, dst = self.blocks[1 - k].data_ptr(), self.blocks[k].data_ptr()
            ro.host_state = (src, src + nb_pos + nb_act, dst, dst + nb_pos, dst + nb_pos + nb_act)
            ro.run(1)                                                     # the kernels read / write the pinned blocks
            return
        env.pos.copy_(self.views[1 - k][1], non_blocking=True)           # H2D: this step's state
        ro.run(1)
        self.blocks[k].copy_(env.host_block, non_blocking=True)           # D2H: actions + new positions + done

    def reset(self):
        """Back to the initial scenario: the device environment AND the pinned state this rollout steps from (with
        ``zero_copy`` the pinned blocks are the state; an ``env.reset()`` alone would leave them where they were)."""
        if self.in_flight:
            self.wait()
        self.ro.env.reset()
        torch.cuda.synchronize()
        if self.zero_copy:
            for b in self.blocks:
                b.copy_(self.ro.env.host_block.cpu())
        else:
            self.pos_in.copy_(self.ro.env.pos.cpu())

    def refresh_weights(self):
        """Call after the network's parameters changed (optimizer step, load_state_dict): re-packs them into the buffer
        the captured graphs read.  Raises if the packed buffer had to move (different device / size): re-create the
        rollout then."""
        if self.ro.net.packed_weights() is not self._packed:
            raise RuntimeError("HostSteppedRollout: the packed weight buffer was re-allocated; build a new rollout")

    def launch(self):
        """Enqueue one step (upload, kernels, download) on this rollout's own stream and return immediately.  Work
        the caller issued on other streams (e.g. ``env.reset()``) is NOT waited for here -- ``step()`` does that."""
        if self.in_flight:
            raise RuntimeError("HostSteppedRollout.launch(): the previous step has not been waited for")
        rc = self._lib.mapf_graph_launch(self._exec[self.cur], self._stream, self._event)
        if rc:
            _lib.check(rc, "graph_launch")
        self.in_flight = True

    def wait(self, numpy=False):
        """Block until the step enqueued by ``launch()`` is visible in pinned host memory; returns its views
        (``numpy=True``: the same memory as numpy arrays, without creating tensor wrappers per step)."""
        if not self.in_flight:
            raise RuntimeError("HostSteppedRollout.wait(): nothing in flight")
        rc = self._lib.mapf_event_wait(self._event)
        if rc:
            _lib.check(rc, "event_wait")
        self.in_flight = False
        k = self.cur
        self.cur = 1 - k
        return self.views_np[k] if numpy else self.views[k]

    def step(self):
        """One host-driven step; returns after the results are visible in pinned host memory."""
        self.stream.wait_stream(torch.cuda.current_stream())   # e.g. an env.reset() issued by the caller
        self.launch()
        return self.wait()


class PipelinedHostRollout:
    """Host-driven stepping of G independent episode groups with their host round trips overlapped: every group is a
    ``HostSteppedRollout`` on its own stream (own environment shard, own pinned blocks, same network), so while the
    host consumes the results of one group and uploads its next state, the GPU runs the step of another.  Each group
    still performs the full protocol of evaluate.py:223-241 per step -- state up from pinned memory, policy + env
    step, actions / positions / done back to pinned memory.

    ``start()`` launches the first step of every group; ``advance(g)`` waits for group g's step, returns its views
    ``(actions, pos_out, done)`` and, unless ``relaunch=False``, immediately enqueues its next step (the caller may
    edit ``groups[g].pos_in`` first by passing ``relaunch=False`` and calling ``groups[g].launch()`` itself)."""

    @staticmethod
    def round_aligned_sizes(episodes, agents, groups, sm_count, tile=32):
        """Group sizes whose encoder launches are whole rounds of the persistent kernel (``sm_count`` tiles of ``tile``
        agent images each) for all groups but the last, so that splitting the batch does not add ragged rounds."""
        E, N = int(episodes), int(agents)
        per_round = sm_count * tile
        rounds = max(1, round(E * N / per_round / groups))
        size = per_round * rounds // N
        if size < 1 or size * (groups - 1) >= E:
            return [E * (g + 1) // groups - E * g // groups for g in range(groups)]
        return [size] * (groups - 1) + [E - size * (groups - 1)]

    def __init__(self, net, config, goals, starts, obstacles=None, groups=2, sensing_range=6, device="cuda",
                 group_sizes=None, zero_copy=False):
        import numpy as np
        goals, starts = np.asarray(goals), np.asarray(starts)
        E = goals.shape[0]
        if not 1 <= groups <= E:
            raise ValueError("groups must be in [1, episodes]")
        if group_sizes is None:
            sms = torch.cuda.get_device_properties(torch.device(device)).multi_processor_count
            # images one CTA of the encoder takes per round: 5 (or 3) warpgroups x 4 images on the tensor-core conv stack
            # (enc_tc.cu), one 32-image tile in the FFMA2 kernel
            from . import ops
            tile = 32
            if ops.wants_encoder_workspace(net_params(net.dims())) and \
                    _lib.load().mapf_encoder_tc_supported(net_params(net.dims())) == 1:
                tile = 20 if net.dims()[2] == 16 else 12      # 3 warpgroups when the first layer has 32 channels
            group_sizes = self.round_aligned_sizes(E, goals.shape[1], groups, sms, tile)
        if len(group_sizes) != groups or sum(group_sizes) != E or min(group_sizes) < 1:
            raise ValueError("group_sizes must be positive and sum to the number of episodes")
        bounds = [0]
        for n in group_sizes:
            bounds.append(bounds[-1] + int(n))
        self.slices = [slice(bounds[g], bounds[g + 1]) for g in range(groups)]
        self.envs, self.rollouts, self.groups = [], [], []
        for sl in self.slices:
            ob = None if obstacles is None or np.asarray(obstacles).shape[1] == 0 else np.asarray(obstacles)[sl]
            env = BatchedGraphEnv(config, goals[sl], starts[sl], ob, sensing_range=sensing_range, device=device)
            ro = Rollout(net, env)
            self.envs.append(env), self.rollouts.append(ro), self.groups.append(HostSteppedRollout(ro, zero_copy=zero_copy))
        self.h2d_bytes = sum(h.h2d_bytes for h in self.groups)
        self.d2h_bytes = sum(h.d2h_bytes for h in self.groups)

    def reset(self):
        """Back to the initial scenario in every group (device state and pinned input positions)."""
        for h in self.groups:
            h.reset()
        torch.cuda.synchronize()

    def start(self):
        for h in self.groups:
            h.launch()

    def advance(self, g, relaunch=True, numpy=False):
        h = self.groups[g]
        out = h.wait(numpy)
        if relaunch:
            h.launch()
        return out

    def drain(self):
        return [h.wait() if h.in_flight else None for h in self.groups]
