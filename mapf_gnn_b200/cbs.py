"""Conflict-Based Search expert with the reference's API.

``Environment(dimension, agents, obstacles)`` / ``CBS(env, verbose).search() -> {agent: [{t, x, y}]}`` as in
``cbs/cbs.py:148-157,316-410``; ``env.a_star.search(agent)`` as in ``cbs/a_star.py:29-102``.  The searches run in the
native library (``csrc/cbs.cu``: host C++, the plans are identical to the reference's, decision for decision); the small
value classes and the one-step helpers of ``Environment`` that the reference's callers and tests touch
(``get_neighbors``, ``get_first_conflict`` ...) are plain Python over the same rules.  On top of the reference surface:
``solve_batch`` plans many independent instances on all host cores in one native call, ``gen_input`` /
``parse_trajectories`` are the instance generator and the schedule -> action-matrix step of
``data_generation/dataset_gen.py:19-108`` and ``trajectory_parser.py:24-65``, and ``generate_cases`` chains them with the
GPU recorder (``dataset.record_cases``) into the reference's on-disk data set.
"""
from __future__ import annotations

import ctypes as C
from itertools import combinations, count
from math import fabs

import numpy as np

from . import _lib


# ---------------------------------------------------------------------------------------------------------------------
# value classes (cbs.py:24-145): equality / hashing by value, so they work as set members and dict keys
# ---------------------------------------------------------------------------------------------------------------------
class Location:
    def __init__(self, x=-1, y=-1):
        self.x, self.y = x, y

    def __eq__(self, other):
        return self.x == other.x and self.y == other.y

    def __hash__(self):
        return hash((self.x, self.y))

    def __str__(self):
        return str((self.x, self.y))

    __repr__ = __str__


class State:
    def __init__(self, time, location):
        self.time, self.location = time, location

    def __eq__(self, other):
        return self.time == other.time and self.location == other.location

    def __hash__(self):
        return hash((self.time, self.location.x, self.location.y))

    def is_equal_except_time(self, state):
        return self.location == state.location

    def __str__(self):
        return str((self.time, self.location.x, self.location.y))

    __repr__ = __str__


class Conflict:
    VERTEX, EDGE = 1, 2

    def __init__(self):
        self.time, self.type = -1, -1
        self.agent_1, self.agent_2 = "", ""
        self.location_1, self.location_2 = Location(), Location()

    def __str__(self):
        return f"({self.time}, {self.agent_1}, {self.agent_2}, {self.location_1}, {self.location_2})"


class VertexConstraint:
    def __init__(self, time, location):
        self.time, self.location = time, location

    def __eq__(self, other):
        return self.time == other.time and self.location == other.location

    def __hash__(self):
        return hash((self.time, self.location.x, self.location.y))

    def __str__(self):
        return f"({self.time}, {self.location})"


class EdgeConstraint:
    def __init__(self, time, location_1, location_2):
        self.time, self.location_1, self.location_2 = time, location_1, location_2

    def __eq__(self, other):
        return self.time == other.time and self.location_1 == other.location_1 and self.location_2 == other.location_2

    def __hash__(self):
        return hash((self.time, self.location_1.x, self.location_1.y, self.location_2.x, self.location_2.y))

    def __str__(self):
        return f"({self.time}, {self.location_1}, {self.location_2})"


class Constraints:
    def __init__(self):
        self.vertex_constraints = set()
        self.edge_constraints = set()

    def add_constraint(self, other):
        self.vertex_constraints |= other.vertex_constraints
        self.edge_constraints |= other.edge_constraints

    def __str__(self):
        return ("VC: " + str([str(vc) for vc in self.vertex_constraints]) + "EC: " +
                str([str(ec) for ec in self.edge_constraints]))


# ---------------------------------------------------------------------------------------------------------------------
# native calls
# ---------------------------------------------------------------------------------------------------------------------
def _obstacle_cells(obstacles):
    """The cells ``(x, y) in obstacles`` can match (cbs.py:262): the reference compares a tuple against the container's
    members, so only tuple members count -- a YAML list ``[x, y]`` never equals ``(x, y)`` and is no obstacle there."""
    return [(int(o[0]), int(o[1])) for o in obstacles if isinstance(o, tuple) and len(o) == 2]


class _Instance:
    """The flat arrays of one planning problem and its ``mapf_cbs_args`` (kept alive together)."""

    def __init__(self, dimension, starts, goals, obstacles, max_path=0, max_high=0, max_low=0):
        self.starts = np.ascontiguousarray(starts, dtype=np.int32).reshape(-1, 2)
        self.goals = np.ascontiguousarray(goals, dtype=np.int32).reshape(-1, 2)
        self.obstacles = np.ascontiguousarray(obstacles, dtype=np.int32).reshape(-1, 2)
        self.n = self.starts.shape[0]
        self.dimension = (int(dimension[0]), int(dimension[1]))
        self.max_high, self.max_low = int(max_high), int(max_low)
        self.alloc(max_path or 2 * (self.dimension[0] + self.dimension[1]) + 16)

    def alloc(self, max_path):
        self.max_path = int(max_path)
        self.path_len = np.zeros((self.n,), dtype=np.int32)
        self.paths = np.zeros((self.n, self.max_path, 2), dtype=np.int32)

    def fill(self, a):
        ptr = lambda arr: arr.ctypes.data_as(C.POINTER(C.c_int32))
        a.width, a.height = self.dimension
        a.num_agents = self.n
        a.starts, a.goals = ptr(self.starts), ptr(self.goals)
        a.num_obstacles = self.obstacles.shape[0]
        a.obstacles = ptr(self.obstacles) if self.obstacles.shape[0] else None
        a.max_high_level_iterations, a.max_low_level_iterations = self.max_high, self.max_low
        a.max_path = self.max_path
        a.path_len, a.paths = ptr(self.path_len), ptr(self.paths)

    def plan(self, a):
        """None when unsolved, else one int32 [len, 2] array of (x, y) per agent."""
        if not a.solved:
            return None
        return [self.paths[i, :self.path_len[i]].copy() for i in range(self.n)]


def solve_batch(instances, threads=0, max_high_level_iterations=0, max_low_level_iterations=0):
    """Plan many independent problems in ONE native call on ``threads`` host threads (0 = all cores).

    instances: iterable of ``(dimension, starts [N,2], goals [N,2], obstacles [M,2])`` with (x, y) cells.  Returns a list
    with, per instance, ``None`` (no plan within the reference's iteration caps, where ``CBS.search`` returns ``{}``) or
    ``(paths, cost)``: one int32 [len, 2] array of (x, y) per agent (index = time step) and the sum of path lengths."""
    lib = _lib.load()
    items = [_Instance(d, s, g, o, 0, max_high_level_iterations, max_low_level_iterations) for d, s, g, o in instances]
    out = [None] * len(items)
    todo = list(range(len(items)))
    while todo:
        arr = (_lib.CbsArgs * len(todo))()
        for k, i in enumerate(todo):
            items[i].fill(arr[k])
        _lib.check(lib.mapf_cbs_solve_batch(arr, len(todo), int(threads)), "cbs_solve_batch")
        again = []
        for k, i in enumerate(todo):
            if arr[k].truncated:                      # a path longer than the buffer: once more with its true length
                items[i].alloc(int(items[i].path_len.max()))
                again.append(i)
            else:
                plan = items[i].plan(arr[k])
                out[i] = None if plan is None else (plan, int(arr[k].cost))
        todo = again
    return out


class AStar:
    """cbs/a_star.py:12-102 over the environment's CURRENT ``constraints``."""

    def __init__(self, env):
        self.env = env
        self.agent_dict = env.agent_dict
        self.admissible_heuristic = env.admissible_heuristic
        self.is_at_goal = env.is_at_goal
        self.get_neighbors = env.get_neighbors

    def reconstruct_path(self, came_from, current):
        path = [current]
        while current in came_from:
            current = came_from[current]
            path.append(current)
        return path[::-1]

    def search(self, agent_name):
        """Space-time A* for one agent; a list of ``State`` from t = 0 to the first arrival at the goal, or ``False``."""
        env = self.env
        names = list(env.agent_dict)
        inst = env._instance()
        cons = env.constraints
        vc = np.array([(c.time, c.location.x, c.location.y) for c in cons.vertex_constraints], dtype=np.int32).reshape(-1, 3)
        ec = np.array([(c.time, c.location_1.x, c.location_1.y, c.location_2.x, c.location_2.y)
                       for c in cons.edge_constraints], dtype=np.int32).reshape(-1, 5)
        lib = _lib.load()
        ptr = lambda arr: arr.ctypes.data_as(C.POINTER(C.c_int32)) if arr.shape[0] else None
        agent = names.index(agent_name)
        while True:
            a = _lib.CbsArgs()
            inst.fill(a)
            found = C.c_int32(0)
            _lib.check(lib.mapf_cbs_a_star(C.byref(a), agent, ptr(vc), vc.shape[0], ptr(ec), ec.shape[0], C.byref(found)),
                       "cbs_a_star")
            if not found.value:
                return False
            n = int(inst.path_len[agent])
            if n <= inst.max_path:
                return [State(t, Location(int(inst.paths[agent, t, 0]), int(inst.paths[agent, t, 1]))) for t in range(n)]
            inst.alloc(n)


class Environment:
    """cbs.py:148-304: the grid, the agents' starts / goals and the constraint set the low-level search obeys."""

    def __init__(self, dimension, agents, obstacles):
        self.dimension = dimension
        self.obstacles = obstacles
        self.agents = agents
        self.agent_dict = {}
        self.make_agent_dict()
        self.constraints = Constraints()
        self.constraint_dict = {}
        self.a_star = AStar(self)

    def _instance(self, max_high=0, max_low=0):
        starts = [(v["start"].location.x, v["start"].location.y) for v in self.agent_dict.values()]
        goals = [(v["goal"].location.x, v["goal"].location.y) for v in self.agent_dict.values()]
        return _Instance(self.dimension, starts, goals, _obstacle_cells(self.obstacles), 0, max_high, max_low)

    def get_neighbors(self, state):
        """Successors in the reference's order: wait, y+1, y-1, x-1, x+1 (cbs.py:159-182)."""
        out = []
        x, y = state.location.x, state.location.y
        for k, (nx, ny) in enumerate(((x, y), (x, y + 1), (x, y - 1), (x - 1, y), (x + 1, y))):
            n = State(state.time + 1, Location(nx, ny) if k else state.location)
            if self.state_valid(n) and (k == 0 or self.transition_valid(state, n)):
                out.append(n)
        return out

    def get_first_conflict(self, solution):
        """Earliest step's first vertex conflict over the agent pairs in order, else that step's first swap; ``False``
        when the plans are conflict-free (cbs.py:184-215)."""
        horizon = max(len(plan) for plan in solution.values())
        for t in range(horizon):
            for a1, a2 in combinations(solution.keys(), 2):
                s1, s2 = self.get_state(a1, solution, t), self.get_state(a2, solution, t)
                if s1.is_equal_except_time(s2):
                    c = Conflict()
                    c.time, c.type, c.location_1, c.agent_1, c.agent_2 = t, Conflict.VERTEX, s1.location, a1, a2
                    return c
            for a1, a2 in combinations(solution.keys(), 2):
                s1a, s1b = self.get_state(a1, solution, t), self.get_state(a1, solution, t + 1)
                s2a, s2b = self.get_state(a2, solution, t), self.get_state(a2, solution, t + 1)
                if s1a.is_equal_except_time(s2b) and s1b.is_equal_except_time(s2a):
                    c = Conflict()
                    c.time, c.type, c.agent_1, c.agent_2 = t, Conflict.EDGE, a1, a2
                    c.location_1, c.location_2 = s1a.location, s1b.location
                    return c
        return False

    def create_constraints_from_conflict(self, conflict):
        """cbs.py:217-245: a vertex conflict bars both agents from (t, cell) -- they share ONE Constraints object, as in
        the reference; an edge conflict bars agent 1 from a->b and agent 2 from b->a at t."""
        out = {}
        if conflict.type == Conflict.VERTEX:
            c = Constraints()
            c.vertex_constraints |= {VertexConstraint(conflict.time, conflict.location_1)}
            out[conflict.agent_1] = c
            out[conflict.agent_2] = c
        elif conflict.type == Conflict.EDGE:
            c1, c2 = Constraints(), Constraints()
            c1.edge_constraints |= {EdgeConstraint(conflict.time, conflict.location_1, conflict.location_2)}
            c2.edge_constraints |= {EdgeConstraint(conflict.time, conflict.location_2, conflict.location_1)}
            out[conflict.agent_1] = c1
            out[conflict.agent_2] = c2
        return out

    def get_state(self, agent_name, solution, t):
        plan = solution[agent_name]
        return plan[t] if t < len(plan) else plan[-1]

    def state_valid(self, state):
        loc = state.location
        return (0 <= loc.x < self.dimension[0] and 0 <= loc.y < self.dimension[1]
                and VertexConstraint(state.time, loc) not in self.constraints.vertex_constraints
                and (loc.x, loc.y) not in self.obstacles)

    def transition_valid(self, state_1, state_2):
        return EdgeConstraint(state_1.time, state_1.location, state_2.location) not in self.constraints.edge_constraints

    def is_solution(self, agent_name):
        pass

    def admissible_heuristic(self, state, agent_name):
        goal = self.agent_dict[agent_name]["goal"].location
        return fabs(state.location.x - goal.x) + fabs(state.location.y - goal.y)

    def is_at_goal(self, state, agent_name):
        return state.is_equal_except_time(self.agent_dict[agent_name]["goal"])

    def make_agent_dict(self):
        for agent in self.agents:
            self.agent_dict.update({agent["name"]: {
                "start": State(0, Location(agent["start"][0], agent["start"][1])),
                "goal": State(0, Location(agent["goal"][0], agent["goal"][1]))}})

    def compute_solution(self):
        """One low-level plan per agent under ``constraint_dict`` (cbs.py:291-300); ``False`` if any agent has none."""
        solution = {}
        for agent in self.agent_dict.keys():
            self.constraints = self.constraint_dict.setdefault(agent, Constraints())
            plan = self.a_star.search(agent)
            if not plan:
                return False
            solution[agent] = plan
        return solution

    def compute_solution_cost(self, solution):
        return sum(len(path) for path in solution.values())


class HighLevelNode:
    def __init__(self):
        self.solution = {}
        self.constraint_dict = {}
        self.cost = 0

    def __lt__(self, other):
        return self.cost < other.cost


class CBS:
    """cbs.py:316-430.  ``search()`` runs the whole high-level search natively; ``closed_set`` afterwards holds one entry
    per expanded constraint-tree node (the reference stores the nodes' solution keys there; callers only size it)."""

    def __init__(self, environment, verbose=True):
        self.env = environment
        self.verbose = verbose
        self.open_list = []
        self.closed_set = set()
        self.counter = count()

    def search(self):
        env = self.env
        if self.verbose:
            print("Initializing search")
        inst = env._instance()
        lib = _lib.load()
        while True:
            a = _lib.CbsArgs()
            inst.fill(a)
            _lib.check(lib.mapf_cbs_solve(C.byref(a)), "cbs_solve")
            if not a.truncated:
                break
            inst.alloc(int(inst.path_len.max()))
        self.closed_set = set(range(int(a.expanded)))
        plan = inst.plan(a)
        if plan is None:
            if self.verbose:
                print("No solution found")
            return {}
        if self.verbose:
            print("Solution found!")
        return {name: [{"t": t, "x": int(x), "y": int(y)} for t, (x, y) in enumerate(path)]
                for name, path in zip(env.agent_dict, plan)}

    def _get_state_key(self, node):
        return tuple((agent, tuple((s.time, s.location.x, s.location.y) for s in path))
                     for agent, path in sorted(node.solution.items()))

    def generate_plan(self, solution):
        return {agent: [{"t": s.time, "x": s.location.x, "y": s.location.y} for s in path]
                for agent, path in solution.items()}


# ---------------------------------------------------------------------------------------------------------------------
# data generation (data_generation/dataset_gen.py, trajectory_parser.py)
# ---------------------------------------------------------------------------------------------------------------------
def gen_input(dimensions, nb_obstacles, nb_agents, max_attempts=10000):
    """A random instance drawn with the SAME sequence of ``np.random`` calls as dataset_gen.py:19-108 (obstacles first,
    then per agent a start and a goal, every cell free of everything placed before), so that a seed gives the
    reference's instance.  ``None`` when a cell cannot be placed."""
    w, h = int(dimensions[0]), int(dimensions[1])
    total = w * h
    taken = set()

    def draw(tries=1000):
        if len(taken) < total * 0.7:                       # sparse board: rejection sampling
            for _ in range(tries):
                pos = (np.random.randint(0, w), np.random.randint(0, h))
                if pos not in taken:
                    return pos
            return None
        free = list({(x, y) for x in range(w) for y in range(h)} - taken)      # dense board: pick among the free cells
        return free[np.random.randint(0, len(free))] if free else None

    if nb_obstacles + 2 * nb_agents > total * 0.9:
        print(f"Warning: Requesting {nb_obstacles + 2 * nb_agents} positions in {total} cells (>90% fill)")
    out = {"agents": [], "map": {"dimensions": [w, h], "obstacles": []}}
    for _ in range(nb_obstacles):
        pos = draw()
        if pos is None:
            return None
        taken.add(pos)
        out["map"]["obstacles"].append(pos)
    for k in range(nb_agents):
        start = draw()
        if start is None:
            return None
        taken.add(start)
        goal = draw()
        if goal is None:
            return None
        taken.add(goal)
        out["agents"].append({"start": list(start), "goal": list(goal), "name": f"agent{k}"})
    return out


def get_longest_path(schedule):
    return max((len(path) for path in schedule.values()), default=0)


_ACTION_OF_MOVE = {(0, 0): 0, (1, 0): 1, (0, 1): 2, (-1, 0): 3, (0, -1): 4}     # trajectory_parser.py:40-46


def parse_trajectories(schedule):
    """Schedule ``{agent: [{t, x, y}]}`` -> (actions int32 [agents, longest], starts int32 [agents, 2]); an agent that
    has arrived waits (action 0), and the last column is all waits (trajectory_parser.py:24-65)."""
    longest = get_longest_path(schedule)
    actions = np.zeros((len(schedule), longest), dtype=np.int32)
    starts = np.zeros((len(schedule), 2), dtype=np.int32)
    for j, path in enumerate(schedule.values()):
        starts[j] = (path[0]["x"], path[0]["y"])
        for i in range(len(path) - 1):
            actions[j, i] = _ACTION_OF_MOVE[(path[i + 1]["x"] - path[i]["x"], path[i + 1]["y"] - path[i]["y"])]
    return actions, starts


def data_gen(input_dict, output_path):
    """dataset_gen.py:111-148: plan one instance and write ``solution.yaml`` ({schedule, cost}) and ``input.yaml`` into
    ``output_path``; False (and no solution file) when there is no plan."""
    import os
    import yaml
    os.makedirs(output_path, exist_ok=True)
    env = Environment(input_dict["map"]["dimensions"], input_dict["agents"], input_dict["map"]["obstacles"])
    solution = CBS(env, verbose=False).search()
    if not solution:
        print(f"No solution found for case {os.path.basename(str(output_path))}")
        return False
    with open(os.path.join(output_path, "solution.yaml"), "w") as fh:
        yaml.safe_dump({"schedule": solution, "cost": env.compute_solution_cost(solution)}, fh)
    with open(os.path.join(output_path, "input.yaml"), "w") as fh:
        yaml.safe_dump(input_dict, fh)
    return True


def create_solutions(path, num_cases, config):
    """dataset_gen.py:151-206: cases ``case_<i>`` for i from the number of existing case directories up to ``num_cases``;
    instances from ``gen_input`` (np.random, in the reference's order -- drawing does not depend on the plans, so all
    instances are drawn first and planned in ONE batch call), failed cases leave no directory."""
    import os
    import shutil
    import yaml
    os.makedirs(path, exist_ok=True)
    ready = len([d for d in os.listdir(path) if d.startswith("case_") and os.path.isdir(os.path.join(path, d))])
    todo = []
    for i in range(ready, num_cases):
        inp = gen_input(tuple(config["map_shape"]), config["nb_obstacles"], config["nb_agents"])
        if inp is not None:
            todo.append((i, inp))
    plans = solve_batch([(inp["map"]["dimensions"], [a["start"] for a in inp["agents"]], [a["goal"] for a in inp["agents"]],
                          _obstacle_cells(inp["map"]["obstacles"])) for _, inp in todo])
    done = 0
    for (i, inp), plan in zip(todo, plans):
        case = os.path.join(path, f"case_{i}")
        if plan is None:
            shutil.rmtree(case, ignore_errors=True)
            continue
        os.makedirs(case, exist_ok=True)
        paths, cost = plan
        with open(os.path.join(case, "solution.yaml"), "w") as fh:
            yaml.safe_dump({"schedule": schedule_of(paths, [a["name"] for a in inp["agents"]]), "cost": cost}, fh)
        with open(os.path.join(case, "input.yaml"), "w") as fh:
            yaml.safe_dump(inp, fh)
        done += 1
    print(f"Generation complete: {done} successful, {num_cases - ready - done} failed")
    return done


def parse_dataset_trajectories(path):
    """trajectory_parser.py:68-106: ``trajectory.npy`` / ``startings.npy`` next to every ``case_*/solution.yaml``."""
    import os
    import yaml
    for name in sorted(d for d in os.listdir(path) if d.startswith("case_") and os.path.isdir(os.path.join(path, d))):
        sol = os.path.join(path, name, "solution.yaml")
        if not os.path.exists(sol):
            print(f"Warning: Solution file not found for {name}")
            continue
        with open(sol) as fh:
            schedule = yaml.load(fh, Loader=yaml.FullLoader)["schedule"]
        trajectory, startings = parse_trajectories(dict(schedule))
        np.save(os.path.join(path, name, "trajectory.npy"), trajectory)
        np.save(os.path.join(path, name, "startings.npy"), startings)


def schedule_of(paths, names=None):
    """``solve_batch`` paths -> the reference's schedule dict."""
    names = names or [f"agent{k}" for k in range(len(paths))]
    return {name: [{"t": t, "x": int(x), "y": int(y)} for t, (x, y) in enumerate(p)] for name, p in zip(names, paths)}


def generate_cases(num_cases, map_shape, nb_agents, nb_obstacles, threads=0):
    """``create_solutions`` (dataset_gen.py:151-206) without the files: draw instances with ``gen_input`` (np.random,
    seed it for reproducibility), plan them all with one ``solve_batch`` call and keep the solved ones.  Returns a list of
    dicts ``{input, schedule, cost, trajectory [N, T], startings [N, 2]}`` ready for ``dataset.record_cases``."""
    inputs = []
    for _ in range(num_cases):
        inp = gen_input(tuple(map_shape), nb_obstacles, nb_agents)
        if inp is not None:
            inputs.append(inp)
    plans = solve_batch([(inp["map"]["dimensions"], [a["start"] for a in inp["agents"]], [a["goal"] for a in inp["agents"]],
                          inp["map"]["obstacles"]) for inp in inputs], threads=threads)
    cases = []
    for inp, plan in zip(inputs, plans):
        if plan is None:
            continue
        paths, cost = plan
        schedule = schedule_of(paths, [a["name"] for a in inp["agents"]])
        trajectory, startings = parse_trajectories(schedule)
        cases.append({"input": inp, "schedule": schedule, "cost": cost, "trajectory": trajectory, "startings": startings})
    return cases
