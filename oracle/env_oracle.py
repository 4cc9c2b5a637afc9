"""Oracle (test infrastructure) -- numpy restatement of the reference grid environment.

Follows ``/root/reference/grid/env_graph_gridv1.py`` (cited per function below).  One
``EnvOracle`` is one episode, on the host, exactly like the reference's ``GraphEnv``; the
batched CUDA environment is checked against E of these run in a Python loop.

Two styles are kept side by side on purpose:

* ``EnvOracle`` mirrors the *structure* of the reference algorithm (per-row sort for the
  closest-4 rule, per-agent window slicing for the FOV, dictionary grouping for collisions)
  so that timing it is a fair stand-in for timing the reference on a box where
  ``/root/reference`` is absent (``bench.py`` ``cpu_baseline.kind == "port"``).
* ``*_closed_form`` functions restate the same results as exact integer formulas; they are
  what the CUDA kernels implement and are cross-checked against ``EnvOracle`` and the goldens.
"""
from __future__ import annotations

import math

import numpy as np

# action id -> (dx, dy); reference env_graph_gridv1.py:42-48
ACTION_DELTAS = np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1]], dtype=np.int64)

FOV_SIDE = 5      # (pad*2)-1 with pad=3, env_graph_gridv1.py:252
FOV_HALF = 2
BOARD_EMPTY, BOARD_AGENT, BOARD_OBSTACLE = 0, 1, 2   # env_graph_gridv1.py:70,275
GOAL_MARK = 3     # env_graph_gridv1.py:263


def d2_limit_from_range(sensing_range: float) -> int:
    """Smallest integer d2 such that float64 sqrt(d2) >= sensing_range.

    The reference zeroes ``D[D >= sensing_range]`` on float64 square roots
    (env_graph_gridv1.py:127-130); on integer squared distances that is ``d2 >= limit``.
    """
    r = float(sensing_range)
    if r <= 0.0:
        return 0
    cand = max(0, int(math.floor(r * r)) - 2)
    while not (np.sqrt(np.float64(cand)) >= r):
        cand += 1
    return cand


# ----------------------------------------------------------------------------------------
# closed forms (what the kernels compute)
# ----------------------------------------------------------------------------------------
def move_closed_form(pos, actions, obstacle_mask, board_size):
    """A1: env_graph_gridv1.py:200-213 as integer logic.  pos int[N,2] (x,y) -> new pos.

    obstacle_mask[y, x] is True where (x, y) is in the obstacle *set* (never changes).
    Order: add delta -> (check_goals is a value no-op :161-166) -> obstacle revert :303-313
    -> clamp :267-271 -> single-pass collision revert of every member of a shared cell :282-301.
    """
    pos = np.asarray(pos, dtype=np.int64)
    old = pos.copy()
    new = pos + ACTION_DELTAS[np.asarray(actions, dtype=np.int64)]
    H = int(board_size)
    inside = (new[:, 0] >= 0) & (new[:, 0] < H) & (new[:, 1] >= 0) & (new[:, 1] < H)
    hit = np.zeros(len(new), dtype=bool)
    hit[inside] = obstacle_mask[new[inside, 1], new[inside, 0]]
    new[hit] = old[hit]
    new = np.clip(new, 0, H - 1)
    key = new[:, 1] * H + new[:, 0]
    same = key[:, None] == key[None, :]
    crowded = same.sum(axis=1) > 1
    new[crowded] = old[crowded]
    return new


def adjacency_closed_form(pos, d2_limit):
    """A2: env_graph_gridv1.py:117-136,174-181 as integer logic -> uint8[N,N].

    A[i,j] = 1 iff 0 < d2 < d2_limit and d2 <= thr_i, where thr_i is the 4th smallest
    in-range non-zero d2 of row i (with multiplicity), or the largest when fewer than 4.
    """
    pos = np.asarray(pos, dtype=np.int64)
    diff = pos[:, None, :] - pos[None, :, :]
    d2 = (diff * diff).sum(axis=2)
    cand = (d2 > 0) & (d2 < d2_limit)
    n = len(pos)
    adj = np.zeros((n, n), dtype=np.uint8)
    for i in range(n):
        vals = np.sort(d2[i][cand[i]])
        if len(vals) == 0:
            continue
        thr = vals[3] if len(vals) >= 4 else vals[-1]
        adj[i] = cand[i] & (d2[i] <= thr)
    return adj


def fov_closed_form(board, pos, goal):
    """A3: env_graph_gridv1.py:218-265 as index arithmetic -> uint8[N,2,5,5].

    ch0[r,c] = board[y+2-r, x-2+c] (0 off-board); ch1 has a single 3 at
    (2 - clamp(gy-y,-2,2), 2 + clamp(gx-x,-2,2)).
    """
    pos = np.asarray(pos, dtype=np.int64)
    goal = np.asarray(goal, dtype=np.int64)
    H = board.shape[0]
    n = len(pos)
    out = np.zeros((n, 2, FOV_SIDE, FOV_SIDE), dtype=np.uint8)
    for a in range(n):
        x, y = pos[a]
        for r in range(FOV_SIDE):
            for c in range(FOV_SIDE):
                by, bx = y + FOV_HALF - r, x - FOV_HALF + c
                if 0 <= by < H and 0 <= bx < H:
                    out[a, 0, r, c] = board[by, bx]
        dx = int(np.clip(goal[a, 0] - x, -FOV_HALF, FOV_HALF))
        dy = int(np.clip(goal[a, 1] - y, -FOV_HALF, FOV_HALF))
        out[a, 1, FOV_HALF - dy, FOV_HALF + dx] = GOAL_MARK
    return out


# ----------------------------------------------------------------------------------------
# structural restatement (what the CPU baseline times)
# ----------------------------------------------------------------------------------------
class EnvOracle:
    """One episode.  Mirrors GraphEnv's observable behaviour (env_graph_gridv1.py:9-198).

    pos/goal are int arrays [N,2] as (x, y); obstacles int[M,2] or None.  ``starts`` must be
    given (the reference's random ``reset`` draw :77-90 is scenario generation, not hot path).
    """

    def __init__(self, num_agents, board_size, goal, starts, obstacles=None,
                 sensing_range=6, max_time=32):
        self.n = int(num_agents)
        self.H = int(board_size)
        self.goal = np.array(goal, dtype=np.int64).reshape(self.n, 2)
        self.starts = np.array(starts, dtype=np.int64).reshape(self.n, 2)
        self.obstacles = None if obstacles is None else np.array(obstacles, dtype=np.int64).reshape(-1, 2)
        self.obstacle_set = set() if self.obstacles is None else {tuple(o) for o in self.obstacles.tolist()}
        self.sensing_range = sensing_range
        self.max_time = int(max_time)
        self.reset()

    # env_graph_gridv1.py:65-95
    def reset(self):
        self.time = 0
        self.board = np.zeros((self.H, self.H))
        if self.obstacles is not None and len(self.obstacles):
            self.board[self.obstacles[:, 1], self.obstacles[:, 0]] = BOARD_OBSTACLE
        self.px = self.starts[:, 0].copy()
        self.py = self.starts[:, 1].copy()
        self._graph()
        return self.observe()

    # env_graph_gridv1.py:97-106 (board-with-goals and embeddings are not model inputs)
    def observe(self):
        return {"fov": self._fov(), "adj_matrix": self.adj}

    # env_graph_gridv1.py:117-136 + 174-181
    def _graph(self):
        p = np.stack([self.px, self.py], axis=1)
        delta = p[:, None, :] - p[None, :, :]
        dist = np.sqrt((delta * delta).sum(axis=2))
        dist[dist >= self.sensing_range] = 0
        for i in range(self.n):
            near = np.sort(dist[i][dist[i] != 0])
            if len(near) < 4:
                near = np.concatenate((np.zeros(4 - len(near)), near))
            dist[i][dist[i] > near[3]] = 0
        dist[dist != 0] = 1
        self.adj = dist

    # env_graph_gridv1.py:248-265 with map_goal :218-246
    def _fov(self):
        padded = np.pad(self.board, 3)
        out = np.zeros((self.n, 2, FOV_SIDE, FOV_SIDE))
        for a in range(self.n):
            x, y = int(self.px[a]), int(self.py[a])
            window = padded[y + 1:y + 6, x + 1:x + 6]
            out[a, 0] = window[::-1]
            gx, gy = int(self.goal[a, 0]), int(self.goal[a, 1])
            # column: inside the window -> offset; else clipped to the border
            if x - 2 < gx < x + 2:
                col = gx - x + 2
            elif gx <= x - 2:
                col = 0
            else:
                col = 4
            if y - 2 <= gy < y + 2:
                row = y - gy + 2
            elif gy <= y - 2:
                row = 4
            else:
                row = 0
            out[a, 1, row, col] = GOAL_MARK
        return out

    # env_graph_gridv1.py:200-213
    def _move(self, actions):
        ox, oy = self.px.copy(), self.py.copy()
        dx = np.array([ACTION_DELTAS[int(a)][0] for a in actions])
        dy = np.array([ACTION_DELTAS[int(a)][1] for a in actions])
        nx, ny = self.px + dx, self.py + dy
        if self.obstacles is not None:                       # :303-313
            for i in range(self.n):
                if (int(nx[i]), int(ny[i])) in self.obstacle_set:
                    nx[i], ny[i] = ox[i], oy[i]
        nx = np.clip(nx, 0, self.H - 1)                      # :267-271
        ny = np.clip(ny, 0, self.H - 1)
        cells = {}                                           # :282-301
        for i in range(self.n):
            cells.setdefault((int(nx[i]), int(ny[i])), []).append(i)
        for members in cells.values():
            if len(members) > 1:
                for i in members:
                    nx[i], ny[i] = ox[i], oy[i]
        self.board[oy, ox] = BOARD_EMPTY                     # :273-275
        self.board[ny, nx] = BOARD_AGENT
        self.px, self.py = nx, ny

    # env_graph_gridv1.py:183-198
    def step(self, actions):
        self._move(actions)
        self._graph()
        obs = self.observe()
        self.time += 1
        return obs, self.all_at_goal()

    # env_graph_gridv1.py:156-159
    def all_at_goal(self):
        return bool(np.array_equal(np.stack([self.px, self.py], axis=1), self.goal))

    # env_graph_gridv1.py:138-154,168-172
    def metrics(self):
        here = np.stack([self.px, self.py], axis=1)
        success = float(np.all(here == self.goal, axis=1).sum() / self.n)
        flow = self.time if self.all_at_goal() else self.n * self.max_time
        return success, flow

    @property
    def positions(self):
        return np.stack([self.px, self.py], axis=1)


# ----------------------------------------------------------------------------------------
# scenario generation shared by tests / bench (NOT the reference RNG stream; §8d says the
# same arrays are fed to both implementations)
# ----------------------------------------------------------------------------------------
def make_scenarios(rng, episodes, num_agents, board_size, num_obstacles):
    """Unique obstacles, then unique obstacle-free goals and unique obstacle-free starts
    (the order evaluate.py:207-208 uses), for ``episodes`` independent episodes.

    Returns starts int32[E,N,2], goals int32[E,N,2], obstacles int32[E,M,2] (x, y).
    """
    H = int(board_size)
    cells = H * H
    starts = np.empty((episodes, num_agents, 2), dtype=np.int32)
    goals = np.empty((episodes, num_agents, 2), dtype=np.int32)
    obst = np.empty((episodes, num_obstacles, 2), dtype=np.int32)
    for e in range(episodes):
        perm = rng.permutation(cells)
        o = perm[:num_obstacles]
        g = perm[num_obstacles:num_obstacles + num_agents]
        free = perm[num_obstacles:]
        s = rng.choice(free, size=num_agents, replace=False)
        obst[e, :, 0], obst[e, :, 1] = o % H, o // H
        goals[e, :, 0], goals[e, :, 1] = g % H, g // H
        starts[e, :, 0], starts[e, :, 1] = s % H, s // H
    return starts, goals, obst
