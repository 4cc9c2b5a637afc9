"""CBS expert: the native search (csrc/cbs.cu behind mapf_gnn_b200/cbs.py) must return the reference's
plans exactly -- golden vectors of the unmodified reference (tests/golden/cbs_cases.npz), the Python oracle as a second
opinion on fresh instances, and the reference's API behaviours.  Host code: these run without a GPU."""
import os

import numpy as np
import pytest

import mapf_gnn_b200.cbs as cbs
from oracle import cbs_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cbs_cases.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _keys(g):
    return sorted({k.split("/")[0] for k in g.files if k.startswith(("seed", "hand_")) and not k.endswith("/none")})


def _instance(g, key):
    dims = tuple(int(v) for v in g[f"{key}/dims"])
    return dims, g[f"{key}/starts"], g[f"{key}/goals"], g[f"{key}/obstacles"]


def _check_plan(g, key, plan):
    if not int(g[f"{key}/solved"]):
        assert plan is None, key
        return
    assert plan is not None, key
    paths, cost = plan
    lens = g[f"{key}/path_len"]
    assert [len(p) for p in paths] == list(lens), key
    for i, p in enumerate(paths):
        assert np.array_equal(p, g[f"{key}/paths"][i, :lens[i]]), (key, i)
    assert cost == int(lens.sum())


def test_batch_solver_reproduces_every_reference_plan(gold):
    keys = _keys(gold)
    assert len(keys) > 200
    plans = cbs.solve_batch([_instance(gold, k) for k in keys])
    for k, plan in zip(keys, plans):
        _check_plan(gold, k, plan)
    assert sum(p is None for p in plans) >= 1          # the fixture holds unsolvable / capped instances too


def test_reference_api_returns_the_same_schedules(gold):
    """Environment / CBS objects as dataset_gen.py:125-133 uses them: schedule dicts, cost, closed-set size."""
    for key in _keys(gold)[:60] + ["hand_swap", "hand_walled", "hand_home"]:
        dims, starts, goals, obst = _instance(gold, key)
        agents = [{"name": f"agent{i}", "start": list(map(int, s)), "goal": list(map(int, t))}
                  for i, (s, t) in enumerate(zip(starts, goals))]
        env = cbs.Environment(list(dims), agents, [tuple(map(int, o)) for o in obst])
        search = cbs.CBS(env, verbose=False)
        plan = search.search()
        if not int(gold[f"{key}/solved"]):
            assert plan == {}
        else:
            lens = gold[f"{key}/path_len"]
            assert list(plan) == [a["name"] for a in agents]
            for i, a in enumerate(agents):
                want = [{"t": t, "x": int(x), "y": int(y)} for t, (x, y) in enumerate(gold[f"{key}/paths"][i, :lens[i]])]
                assert plan[a["name"]] == want
            assert env.compute_solution_cost(plan) == int(lens.sum())
            traj, st = cbs.parse_trajectories(plan)
            assert np.array_equal(traj, gold[f"{key}/trajectory"]) and traj.dtype == np.int32
            assert np.array_equal(st, gold[f"{key}/startings"])
        assert len(search.closed_set) == int(gold[f"{key}/expanded"]), key


def test_generator_draws_the_reference_instances(gold):
    """gen_input consumes np.random exactly like dataset_gen.py:19-108: same seed -> same instance."""
    # the (grid, agents, obstacles) cycle of tests/golden/make_golden_cbs.py
    combo = lambda seed: ([(8, 8), (10, 10), (6, 6), (16, 16), (5, 5)][seed % 5], [2, 3, 4, 5, 6][(seed // 5) % 5],
                          [0, 3, 6, 10][(seed // 25) % 4])
    n = int(gold["meta/n_seeds"])
    for seed in range(n):
        np.random.seed(seed)
        dims, na, no = combo(seed)
        inp = cbs.gen_input(dims, no, na)
        if f"seed{seed}/none" in gold.files:
            assert inp is None
            continue
        assert tuple(gold[f"seed{seed}/dims"]) == dims
        assert np.array_equal(np.array(inp["map"]["obstacles"], dtype=np.int32).reshape(-1, 2), gold[f"seed{seed}/obstacles"])
        assert np.array_equal(np.array([a["start"] for a in inp["agents"]]), gold[f"seed{seed}/starts"])
        assert np.array_equal(np.array([a["goal"] for a in inp["agents"]]), gold[f"seed{seed}/goals"])
        assert [a["name"] for a in inp["agents"]] == [f"agent{i}" for i in range(na)]


def test_low_level_search_under_constraints(gold):
    env = cbs.Environment([5, 5], [{"name": "a", "start": [0, 0], "goal": [4, 0]}], [(2, 1)])
    env.constraints.vertex_constraints |= {cbs.VertexConstraint(1, cbs.Location(1, 0)), cbs.VertexConstraint(2, cbs.Location(2, 0))}
    env.constraints.edge_constraints |= {cbs.EdgeConstraint(0, cbs.Location(0, 0), cbs.Location(0, 1))}
    path = env.a_star.search("a")
    assert [(s.location.x, s.location.y) for s in path] == [tuple(p) for p in gold["astar/path"]]
    assert [s.time for s in path] == list(range(len(path)))
    want = cbs_oracle.low_level((5, 5), {(2, 1)}, (0, 0), (4, 0), {(1, 1, 0), (2, 2, 0)}, {(0, 0, 0, 0, 1)})
    assert [(s.location.x, s.location.y) for s in path] == want


def test_oracle_is_pinned_to_the_reference(gold):
    """The Python restatement on the cheaper golden instances (it is ~30x slower than the native search)."""
    checked = 0
    for key in _keys(gold):
        if int(gold[f"{key}/expanded"]) > 40:
            continue
        dims, starts, goals, obst = _instance(gold, key)
        paths, cost, expanded = cbs_oracle.search(dims, starts.tolist(), goals.tolist(), obst.tolist())
        _check_plan(gold, key, None if paths is None else ([np.array(p, dtype=np.int32).reshape(-1, 2) for p in paths], cost))
        assert expanded == int(gold[f"{key}/expanded"])
        checked += 1
    assert checked > 100


def test_native_equals_oracle_on_fresh_crowded_instances():
    rng = np.random.default_rng(11)
    insts = []
    for _ in range(40):
        w, h = int(rng.integers(3, 8)), int(rng.integers(3, 8))
        n = min(int(rng.integers(2, 6)), w * h // 2)
        cells = rng.permutation(w * h)
        m = int(rng.integers(0, max(1, w * h - 2 * n)))
        pick = lambda idx: np.stack([cells[idx] % w, cells[idx] // w], axis=1).astype(np.int32)
        insts.append(((w, h), pick(slice(0, n)), pick(slice(n, 2 * n)), pick(slice(2 * n, 2 * n + min(m, 6)))))
    plans = cbs.solve_batch(insts, threads=3, max_high_level_iterations=300)
    solved = 0
    for (dims, s, g, o), plan in zip(insts, plans):
        paths, cost, _ = cbs_oracle.search(dims, s.tolist(), g.tolist(), o.tolist(), max_nodes=300)
        if paths is None:
            assert plan is None
            continue
        solved += 1
        assert plan is not None and plan[1] == cost
        for p, q in zip(plan[0], paths):
            assert [tuple(c) for c in p] == q
    assert solved > 10


def test_plans_are_conflict_free_and_reach_the_goals(gold):
    for key in _keys(gold)[:80]:
        if not int(gold[f"{key}/solved"]):
            continue
        dims, starts, goals, obst = _instance(gold, key)
        (paths, _), = cbs.solve_batch([(dims, starts, goals, obst)])
        T = max(len(p) for p in paths)
        pos = np.stack([np.concatenate([p, np.repeat(p[-1:], T - len(p), axis=0)]) for p in paths])      # [N, T, 2]
        obstacles = {tuple(o) for o in obst}
        for i, p in enumerate(paths):
            assert tuple(p[0]) == tuple(starts[i]) and tuple(p[-1]) == tuple(goals[i])
            assert np.abs(np.diff(p, axis=0)).sum(axis=1).max(initial=0) <= 1
            assert not any(tuple(c) in obstacles for c in p)
        for t in range(T):
            assert len({tuple(c) for c in pos[:, t]}) == len(paths)                       # no vertex conflict
        for i in range(len(paths)):
            for j in range(i + 1, len(paths)):
                swap = (pos[i, :-1] == pos[j, 1:]).all(axis=1) & (pos[i, 1:] == pos[j, :-1]).all(axis=1)
                assert not swap.any()


def test_long_paths_outgrow_the_first_buffer():
    """A 1 x 40 corridor: the path (40 cells) is longer than the default buffer of a 1-row grid's... the call retries with the
    true length (mapf_cbs_args.truncated)."""
    dims = (60, 1)
    (paths, cost), = cbs.solve_batch([(dims, [[0, 0]], [[59, 0]], np.zeros((0, 2), np.int32))])
    assert len(paths[0]) == 60 and cost == 60
    inst = cbs._Instance(dims, [[0, 0]], [[59, 0]], np.zeros((0, 2), np.int32), max_path=8)
    from mapf_gnn_b200 import _lib
    a = _lib.CbsArgs()
    inst.fill(a)
    assert _lib.load().mapf_cbs_solve(a) == 0
    assert a.solved == 1 and a.truncated == 1 and inst.path_len[0] == 60
    assert np.array_equal(inst.paths[0, :, 0], np.arange(8))


def test_list_obstacles_are_ignored_like_in_the_reference():
    """cbs.py:262 tests ``(x, y) in obstacles``: a YAML list [x, y] never equals the tuple, so it blocks nothing."""
    agents = [{"name": "a", "start": [0, 0], "goal": [2, 0]}]
    blocked = cbs.CBS(cbs.Environment([3, 1], agents, [(1, 0)]), verbose=False).search()
    ignored = cbs.CBS(cbs.Environment([3, 1], agents, [[1, 0]]), verbose=False).search()
    assert blocked == {} and len(ignored["a"]) == 3
    assert cbs.Environment([3, 1], agents, [[1, 0]]).state_valid(cbs.State(1, cbs.Location(1, 0)))


def test_argument_errors():
    from mapf_gnn_b200 import _lib
    lib = _lib.load()
    assert lib.mapf_cbs_solve(None) == -1
    a = _lib.CbsArgs()
    assert lib.mapf_cbs_solve(a) == -2            # zero-sized grid
    assert lib.mapf_cbs_solve_batch(None, 0, 0) == 0 and lib.mapf_cbs_solve_batch(None, 2, 0) == -1


def test_generate_cases_feeds_the_recorder_format():
    np.random.seed(5)
    cases = cbs.generate_cases(12, (10, 10), 4, 6, threads=2)
    assert 8 <= len(cases) <= 12
    for c in cases:
        n, t = c["trajectory"].shape
        assert n == 4 and t == max(len(p) for p in c["schedule"].values())
        assert (c["trajectory"][:, -1] == 0).all()                                         # the last column is all waits
        pos = c["startings"].copy()
        delta = np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1]])
        for k in range(t):
            pos += delta[c["trajectory"][:, k]]
        assert np.array_equal(pos, np.array([a["goal"] for a in c["input"]["agents"]]))


def test_low_level_search_matches_oracle_under_random_constraints():
    """mapf_cbs_a_star (Environment.a_star.search with env.constraints set) against the Python restatement on random small
    grids with random obstacles, vertex and edge constraints -- including searches that fail."""
    rng = np.random.default_rng(5)
    found = failed = 0
    for trial in range(300):
        w, h = int(rng.integers(2, 7)), int(rng.integers(2, 7))
        cells = [(x, y) for x in range(w) for y in range(h)]
        rng.shuffle(cells)
        start, goal = cells[0], cells[1]
        obstacles = set(cells[2:2 + int(rng.integers(0, max(1, len(cells) // 3)))])
        vertex = {(int(rng.integers(0, 8)), int(rng.integers(0, w)), int(rng.integers(0, h))) for _ in range(int(rng.integers(0, 12)))}
        edge = set()
        for _ in range(int(rng.integers(0, 10))):
            x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
            dx, dy = [(0, 1), (0, -1), (1, 0), (-1, 0)][int(rng.integers(0, 4))]
            edge.add((int(rng.integers(0, 8)), x, y, x + dx, y + dy))
        want = cbs_oracle.low_level((w, h), obstacles, start, goal, vertex, edge, max_pops=2000)
        env = cbs.Environment([w, h], [{"name": "a", "start": list(start), "goal": list(goal)}], sorted(obstacles))
        env.constraints.vertex_constraints |= {cbs.VertexConstraint(t, cbs.Location(x, y)) for t, x, y in vertex}
        env.constraints.edge_constraints |= {cbs.EdgeConstraint(t, cbs.Location(a, b), cbs.Location(c, d)) for t, a, b, c, d in edge}
        inst = env._instance(max_low=2000)
        env._instance = lambda *a, **k: inst            # same pop budget as the oracle call
        got = env.a_star.search("a")
        if want is None:
            assert got is False, trial
            failed += 1
        else:
            assert [(s.location.x, s.location.y) for s in got] == want, trial
            found += 1
    assert found > 100 and failed >= 3
