"""Grid environment on the GPU: ``BatchedGraphEnv`` (E episodes resident in HBM) and the
reference-compatible single-episode facade ``GraphEnv``.

Mirrors the interface of the reference's ``grid/env_graph_gridv1.py`` (``GraphEnv`` :9-198,
``create_goals`` :465-501, ``create_obstacles`` :504-529): same constructor arguments, ``reset()``,
``step(actions, emb) -> (obs, {}, done, {})``, observation keys ``board / fov / adj_matrix /
distances / embeddings`` and ``computeMetrics()``.  All dynamics and observations are computed by
the CUDA kernels behind ``mapf_env_reset / mapf_env_step / mapf_env_metrics``; there is no host
implementation to fall back to.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib


def d2_limit_from_range(sensing_range: float) -> int:
    """Smallest integer d2 with float64 ``sqrt(d2) >= sensing_range``: the reference zeroes
    ``D[D >= sensing_range]`` on float64 roots (env_graph_gridv1.py:127-130)."""
    r = float(sensing_range)
    if r <= 0.0:
        return 0
    cand = max(0, int(math.floor(r * r)) - 2)
    while math.sqrt(float(cand)) < r:
        cand += 1
    return cand


def _as_i32(a, device, shape):
    t = torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a)
    t = t.to(device=device, dtype=torch.int32).reshape(shape).contiguous()
    return t


class BatchedGraphEnv:
    """E independent episodes of the reference ``GraphEnv`` stepped together on one GPU.

    State (device tensors): ``pos``/``goal`` int32[E,N,2] as (x, y); ``cells`` uint8[E,stride]
    (board value in the low bits, obstacle-set flag in bit 7); ``time`` int32[E]; ``done`` uint8[E].
    Observations: ``fov`` f32[E,N,2,5,5] and ``adj_matrix`` f32[E,N,N], rewritten in place by every
    ``step`` (clone them to keep a history).  Episodes that are done stay frozen.
    """

    def __init__(self, config, goals, starting_positions, obstacles=None, sensing_range=6, device="cuda"):
        self.config = config
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BatchedGraphEnv runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.load()
        self.nb_agents = int(config["num_agents"])
        self.board_size = int(config["board_size"][0])
        self.max_time = int(config["max_time"])
        self.sensing_range = sensing_range
        self.d2_limit = d2_limit_from_range(sensing_range)
        self.episodes_per_cta = 0        # work decomposition of the step kernel: 0 = the library's choice
        goals = torch.as_tensor(np.asarray(goals)) if not torch.is_tensor(goals) else goals
        if goals.dim() != 3 or goals.shape[1] != self.nb_agents or goals.shape[2] != 2:
            raise ValueError(f"goals must be [E,{self.nb_agents},2], got {tuple(goals.shape)}")
        E, N, H = int(goals.shape[0]), self.nb_agents, self.board_size
        self.episodes = E
        self.goal = _as_i32(goals, self.device, (E, N, 2))
        self.starts = _as_i32(starting_positions, self.device, (E, N, 2))
        if obstacles is not None and int(np.prod(tuple(obstacles.shape))) > 0:
            self.obstacles = _as_i32(obstacles, self.device, (E, -1, 2))
        else:
            self.obstacles = None
        self.cell_stride = (H * H + 15) // 16 * 16
        dev = self.device
        # positions, the action slot used by Rollout and the done flags share ONE allocation so that a host-driven
        # caller can fetch all three with a single device-to-host copy (HostSteppedRollout)
        nb_pos, nb_act = E * N * 8, E * N * 4
        self.host_block = torch.zeros((nb_pos + nb_act + (E + 15) // 16 * 16,), dtype=torch.uint8, device=dev)
        self.pos = self.host_block[:nb_pos].view(torch.int32).view(E, N, 2)
        self.actions = self.host_block[nb_pos:nb_pos + nb_act].view(torch.int32).view(E, N)
        self.done = self.host_block[nb_pos + nb_act:nb_pos + nb_act + E]
        self.cells = torch.empty((E, self.cell_stride), dtype=torch.uint8, device=dev)
        self.time = torch.empty((E,), dtype=torch.int32, device=dev)
        self.fov = torch.empty((E, N, 2, 5, 5), dtype=torch.float32, device=dev)
        self.adj_matrix = torch.empty((E, N, N), dtype=torch.float32, device=dev)
        self.reset()

    # -- reference: GraphEnv.reset, env_graph_gridv1.py:65-95 (explicit starts) ----------------
    def reset(self):
        a = _lib.EnvResetArgs()
        a.episodes, a.agents, a.board = self.episodes, self.nb_agents, self.board_size
        a.num_obstacles = 0 if self.obstacles is None else int(self.obstacles.shape[1])
        a.cell_stride = self.cell_stride
        a.starts, a.obstacles = _lib.ptr(self.starts), _lib.ptr(self.obstacles)
        a.pos, a.cells, a.time, a.done = (_lib.ptr(self.pos), _lib.ptr(self.cells), _lib.ptr(self.time),
                                          _lib.ptr(self.done))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mapf_env_reset(a, _lib.current_stream()), "env_reset")
        return self.observe()

    def step_args(self, actions=None, do_move=1):
        a = _lib.EnvStepArgs()
        a.episodes, a.agents, a.board = self.episodes, self.nb_agents, self.board_size
        a.cell_stride, a.d2_limit, a.do_move = self.cell_stride, self.d2_limit, do_move
        a.pos, a.goal, a.actions = _lib.ptr(self.pos), _lib.ptr(self.goal), _lib.ptr(actions)
        a.cells, a.time, a.done = _lib.ptr(self.cells), _lib.ptr(self.time), _lib.ptr(self.done)
        a.fov, a.adj = _lib.ptr(self.fov), _lib.ptr(self.adj_matrix)
        a.episodes_per_cta = self.episodes_per_cta
        return a

    # -- reference: GraphEnv.getObservations, env_graph_gridv1.py:97-106 ------------------------
    def observe(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mapf_env_step(self.step_args(None, 0), _lib.current_stream()), "env_observe")
        return {"fov": self.fov, "adj_matrix": self.adj_matrix, "distances": self.adj_matrix}

    # -- reference: GraphEnv.step, env_graph_gridv1.py:183-198 -----------------------------------
    def step(self, actions, freeze_done=True):
        """freeze_done=False reproduces the raw reference env, which keeps moving agents after all of them reached
        their goals (only its callers stop, evaluate.py:251); used when replaying expert trajectories."""
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions))
        actions = actions.to(device=self.device, dtype=torch.int32).reshape(self.episodes, self.nb_agents).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mapf_env_step(self.step_args(actions, 1 if freeze_done else 3), _lib.current_stream()),
                       "env_step")
        obs = {"fov": self.fov, "adj_matrix": self.adj_matrix, "distances": self.adj_matrix}
        return obs, self.done

    # -- reference: GraphEnv.computeMetrics, env_graph_gridv1.py:138-154,168-172 -----------------
    def computeMetrics(self):
        success = torch.empty((self.episodes,), dtype=torch.float32, device=self.device)
        flow = torch.empty((self.episodes,), dtype=torch.int32, device=self.device)
        a = _lib.EnvMetricsArgs()
        a.episodes, a.agents, a.max_time = self.episodes, self.nb_agents, self.max_time
        a.pos, a.goal, a.time = _lib.ptr(self.pos), _lib.ptr(self.goal), _lib.ptr(self.time)
        a.success, a.flow_time = _lib.ptr(success), _lib.ptr(flow)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mapf_env_metrics(a, _lib.current_stream()), "env_metrics")
        return success, flow

    def board(self):
        """uint8[E,H,H] board values (0 empty, 1 agent, 2 obstacle), env_graph_gridv1.py:68-70,273-275."""
        H = self.board_size
        return (self.cells[:, :H * H] & 3).reshape(self.episodes, H, H)


class GraphEnv:
    """Single-episode facade with the reference's constructor and method names
    (env_graph_gridv1.py:9-198) over a one-episode ``BatchedGraphEnv``.  Observations come back
    as float64 numpy arrays like the reference's; ``distances`` aliases ``adj_matrix`` (:133-136)."""

    def __init__(self, config, goal, max_time=23, board_size=10, sensing_range=6, pad=3,
                 starting_positions=None, obstacles=None, device="cuda"):
        self.config = config
        self.max_time = config["max_time"]
        self.min_time = config.get("min_time", 1)
        self.board_size = config["board_size"][0]
        self.nb_agents = config["num_agents"]
        self.obstacles = obstacles
        self.goal = np.asarray(goal)
        self.pad = pad
        self.sensing_range = sensing_range
        self.starting_positions = starting_positions
        self.device = device
        self.embedding = np.ones(self.nb_agents)
        self.time = 0
        self._env = None
        _ = self.reset()

    def _draw_starts(self):
        # env_graph_gridv1.py:77-90: x from obstacle-free columns, y from obstacle-free rows, with replacement
        xs = np.arange(self.board_size)
        ys = np.arange(self.board_size)
        if self.obstacles is not None:
            xs = xs[~np.isin(xs, self.obstacles[:, 0])]
            ys = ys[~np.isin(ys, self.obstacles[:, 1])]
        return np.stack([np.random.choice(xs, size=self.nb_agents), np.random.choice(ys, size=self.nb_agents)], axis=1)

    def reset(self):
        if self.starting_positions is not None:
            assert self.starting_positions.shape[0] == self.nb_agents, "Agents and positions are not equal"
            starts = np.array(self.starting_positions)
        else:
            starts = self._draw_starts()
        obst = None if self.obstacles is None else np.asarray(self.obstacles)[None]
        self._env = BatchedGraphEnv(self.config, self.goal[None], starts[None], obst, self.sensing_range, self.device)
        self.time = 0
        self.embedding = np.ones(self.nb_agents).reshape((self.nb_agents, 1))
        self._pull()
        self.positionX_temp, self.positionY_temp = self.positionX.copy(), self.positionY.copy()
        return self.getObservations()

    def _pull(self, positions=True):
        if positions:
            pos = self._env.pos[0].cpu().numpy()
            self.positionX, self.positionY = pos[:, 0].copy(), pos[:, 1].copy()
        self.board = self._env.board()[0].cpu().numpy().astype(np.float64)
        self.adj_matrix = self._env.adj_matrix[0].cpu().numpy().astype(np.float64)
        self.distance_matrix = self.adj_matrix
        self._fov = self._env.fov[0].cpu().numpy().astype(np.float64)
        self.time = int(self._env.time[0].item())

    # ``positionX / positionY / positionX_temp / positionY_temp`` are host-writable like the reference's attributes
    # (tests/test_env.py:97-121,164-194 poke them): every kernel call below first uploads the host values.
    def _host_pos(self, x, y):
        return torch.from_numpy(np.stack([np.asarray(x).reshape(-1), np.asarray(y).reshape(-1)], axis=1)
                                .astype(np.int32)).reshape(1, self.nb_agents, 2)

    def _push(self):
        self._env.pos.copy_(self._host_pos(self.positionX, self.positionY))

    def _apply_rules(self, stages):
        env = self._env
        self._push()
        temp = self._host_pos(self.positionX_temp, self.positionY_temp).to(env.device)
        a = _lib.EnvRulesArgs()
        a.episodes, a.agents, a.board, a.cell_stride, a.stages = 1, self.nb_agents, self.board_size, env.cell_stride, stages
        a.pos, a.pos_temp, a.cells = _lib.ptr(env.pos), _lib.ptr(temp), _lib.ptr(env.cells)
        with torch.cuda.device(env.device):
            _lib.check(env.lib.mapf_env_apply_rules(a, _lib.current_stream()), "env_apply_rules")
        pos = env.pos[0].cpu().numpy()
        self.positionX, self.positionY = pos[:, 0].copy(), pos[:, 1].copy()

    def check_collision_obstacle(self):
        """env_graph_gridv1.py:303-313: candidates inside the obstacle set revert to positionX/Y_temp."""
        self._apply_rules(1)

    def check_boundary(self):
        """env_graph_gridv1.py:267-271: clamp to the board."""
        self._apply_rules(2)

    def check_collisions(self):
        """env_graph_gridv1.py:282-301: every member of a shared cell reverts to positionX/Y_temp (single pass)."""
        self._apply_rules(4)

    def check_goals(self):
        """env_graph_gridv1.py:161-166 is a value no-op (agents at their goal are not frozen)."""

    def updateBoard(self):
        """env_graph_gridv1.py:273-275 on the device board: old cells <- 0, new cells <- 1 (obstacle-set flag kept)."""
        env, H = self._env, self.board_size
        old = torch.as_tensor(np.asarray(self.positionY_temp).reshape(-1) * H + np.asarray(self.positionX_temp).reshape(-1),
                              device=env.device, dtype=torch.long)
        new = torch.as_tensor(np.asarray(self.positionY).reshape(-1) * H + np.asarray(self.positionX).reshape(-1),
                              device=env.device, dtype=torch.long)
        env.cells[0, old] &= 0x80
        env.cells[0, new] = (env.cells[0, new] & 0x80) | 1
        self.board = env.board()[0].cpu().numpy().astype(np.float64)

    def _computeDistance(self):
        """env_graph_gridv1.py:117-136 (+ _computeClosest :174-181) for the current host positions."""
        self._push()
        self._env.observe()
        self._pull(positions=False)

    def getObservations(self):
        return {"board": self._board_with_goals(), "fov": self.preprocessObs(), "adj_matrix": self.adj_matrix,
                "distances": self.distance_matrix, "embeddings": self.embedding}

    def preprocessObs(self):
        """env_graph_gridv1.py:248-265 for the current host positions and the current board."""
        self._push()
        self._env.observe()
        self._fov = self._env.fov[0].cpu().numpy().astype(np.float64)
        return self._fov

    def getGraph(self):
        return self.adj_matrix

    def getEmbedding(self):
        return self.embedding.copy()

    def getPositions(self):
        return np.array([self.positionX, self.positionY]).T

    def step(self, actions, emb=None):
        """env_graph_gridv1.py:183-198.  Like the reference, the env itself never freezes: agents keep moving and the
        clock keeps running after all of them reached their goals (only the callers stop, evaluate.py:251)."""
        if emb is not None:
            self.embedding = emb
        self._push()
        self.positionX_temp, self.positionY_temp = np.array(self.positionX).copy(), np.array(self.positionY).copy()
        self._env.step(np.asarray(actions).reshape(1, self.nb_agents), freeze_done=False)
        self._pull()
        obs = {"board": self._board_with_goals(), "fov": self._fov, "adj_matrix": self.adj_matrix,
               "distances": self.distance_matrix, "embeddings": self.embedding}
        return obs, {}, self.checkAllInGoal(), {}

    def _board_with_goals(self):
        board = self.board.copy()
        board[self.goal[:, 1], self.goal[:, 0]] += 4          # updateBoardGoal :277-280
        return board

    def checkAllInGoal(self):
        return bool(np.array_equal(self.getPositions(), self.goal))

    def computeMetrics(self):
        self._push()
        success, flow = self._env.computeMetrics()
        # the kernel's fp32 ratio -> the reference's float64 ratio of the same integer count (tests compare == 0.4)
        hits = int(round(float(success[0].item()) * self.nb_agents))
        return hits / self.nb_agents, int(flow[0].item())

    def computeFlowTime(self):
        return self.computeMetrics()[1]


def create_obstacles(board_size, nb_obstacles):
    """Unique random obstacle cells, int array [M,2] as (x, y) (env_graph_gridv1.py:504-529)."""
    w, h = int(board_size[0]), int(board_size[1])
    if nb_obstacles > w * h:
        raise ValueError(f"Cannot place {nb_obstacles} obstacles on {w}x{h} board")
    idx = np.random.choice(w * h, size=nb_obstacles, replace=False)
    return np.stack([idx % w, idx // w], axis=1)


def create_goals(board_size, num_agents, obstacles=None):
    """Unique random goal cells avoiding obstacles, int array [N,2] (env_graph_gridv1.py:465-501)."""
    w, h = int(board_size[0]), int(board_size[1])
    free = np.ones(w * h, dtype=bool)
    if obstacles is not None and len(obstacles) > 0:
        obstacles = np.asarray(obstacles)
        free[obstacles[:, 1] * w + obstacles[:, 0]] = False
    valid = np.flatnonzero(free)
    if len(valid) < num_agents:
        raise ValueError(f"Not enough valid positions ({len(valid)}) for {num_agents} agents")
    idx = valid[np.random.choice(len(valid), size=num_agents, replace=False)]
    return np.stack([idx % w, idx // w], axis=1)
