"""GPU parity of the batched environment kernels (through the C ABI) against the reference's golden
traces and the numpy oracle.  Integer work: everything must be bit-exact."""
import numpy as np
import pytest
import torch

from oracle import env_oracle as eo

pytestmark = pytest.mark.gpu


def _cfg(n, h, max_time=32):
    return {"num_agents": n, "board_size": [h, h], "max_time": max_time, "min_time": 1}


def _run_case_batched(c, reps):
    """Replicate one golden case `reps` times in one batch (covers multi-episode CTAs and tails)."""
    import mapf_gnn_b200 as m
    n, h, mm, T, max_time = (int(v) for v in c["meta"])
    tile = lambda a: np.repeat(np.asarray(a)[None], reps, axis=0)
    env = m.BatchedGraphEnv(_cfg(n, h, max_time), tile(c["goals"]), tile(c["starts"]),
                            tile(c["obstacles"]) if mm else None, sensing_range=float(c["range"][0]))
    obs = env.observe()
    out = {"fov": [obs["fov"].cpu().numpy().copy()], "adj": [obs["adj_matrix"].cpu().numpy().copy()],
           "pos": [env.pos.cpu().numpy().copy()], "done": []}
    for t in range(T):
        obs, done = env.step(tile(c["actions"][t]))
        out["fov"].append(obs["fov"].cpu().numpy().copy())
        out["adj"].append(obs["adj_matrix"].cpu().numpy().copy())
        out["pos"].append(env.pos.cpu().numpy().copy())
        out["done"].append(done.cpu().numpy().copy())
    return env, out


@pytest.mark.parametrize("reps", [1, 3, 37])
def test_golden_traces_bit_exact(env_cases, reps):
    for name, c in env_cases.items():
        n, h, mm, T, max_time = (int(v) for v in c["meta"])
        env, out = _run_case_batched(c, reps)
        for r in (0, reps - 1):
            for t in range(T + 1):
                assert np.array_equal(out["pos"][t][r], c["pos"][t]), (name, t)
                assert np.array_equal(out["fov"][t][r], c["fov"][t].astype(np.float32)), (name, t)
                assert np.array_equal(out["adj"][t][r], c["adj"][t].astype(np.float32)), (name, t)
        # golden traces keep stepping after done (the reference env does not freeze itself); the batched env
        # freezes finished episodes, so compare done up to and including the first True
        ref_done = c["done"].astype(bool)
        first = int(np.argmax(ref_done)) if ref_done.any() else T
        for t in range(min(first + 1, T)):
            assert bool(out["done"][t][0]) == bool(ref_done[t]), (name, t)
        if not ref_done.any():
            assert np.array_equal(env.board()[0].cpu().numpy(), c["board_final"]), name
            success, flow = env.computeMetrics()
            assert abs(float(success[0]) - c["metrics"][0]) < 1e-7 and int(flow[0]) == int(c["metrics"][1]), name


def test_done_freezes_episode(env_cases):
    import mapf_gnn_b200 as m
    c = env_cases["goalrun_3a_8"]
    n, h, mm, T, max_time = (int(v) for v in c["meta"])
    env = m.BatchedGraphEnv(_cfg(n, h, max_time), c["goals"][None], c["starts"][None], None,
                            sensing_range=float(c["range"][0]))
    first = int(np.argmax(c["done"]))
    for t in range(first + 1):
        _, done = env.step(c["actions"][t][None])
    assert bool(done[0])
    pos, time = env.pos.clone(), env.time.clone()
    fov = env.fov.clone()
    for _ in range(3):
        env.step(np.full((1, n), 1))
    assert torch.equal(env.pos, pos) and torch.equal(env.time, time) and torch.equal(env.fov, fov)
    success, flow = env.computeMetrics()
    assert float(success[0]) == 1.0 and int(flow[0]) == first + 1


CONFIGS = [  # name, E, N, H, M, r, T, seed   (SURVEY 8d parity subsets)
    ("C1", 64, 5, 18, 5, 6, 32, 0),
    ("C2", 64, 5, 28, 8, 6, 32, 1),
    ("C3", 48, 20, 32, 40, 6, 16, 2),
    ("C5", 10, 64, 64, 200, 6, 8, 4),
    ("r4", 33, 7, 9, 4, 4, 12, 5),
    ("n100", 3, 100, 40, 30, 5, 4, 6),
]


@pytest.mark.parametrize("name,E,N,H,M,r,T,seed", CONFIGS)
def test_random_rollouts_match_oracle(name, E, N, H, M, r, T, seed):
    import mapf_gnn_b200 as m
    rng = np.random.default_rng(seed)
    starts, goals, obst = eo.make_scenarios(rng, E, N, H, M)
    env = m.BatchedGraphEnv(_cfg(N, H), goals, starts, obst, sensing_range=r)
    oracles = [eo.EnvOracle(N, H, goals[e], starts[e], obst[e], sensing_range=r) for e in range(E)]
    obs = env.observe()
    for e in range(E):
        o = oracles[e].observe()
        assert np.array_equal(obs["fov"][e].cpu().numpy(), o["fov"].astype(np.float32))
        assert np.array_equal(obs["adj_matrix"][e].cpu().numpy(), o["adj_matrix"].astype(np.float32))
    frozen = np.zeros(E, dtype=bool)
    for t in range(T):
        actions = rng.integers(0, 5, size=(E, N))
        obs, done = env.step(actions)
        fov, adj, pos = obs["fov"].cpu().numpy(), obs["adj_matrix"].cpu().numpy(), env.pos.cpu().numpy()
        done = done.cpu().numpy().astype(bool)
        for e in range(E):
            if frozen[e]:
                continue
            o, d = oracles[e].step(actions[e])
            assert np.array_equal(pos[e], oracles[e].positions), (name, t, e)
            assert np.array_equal(fov[e], o["fov"].astype(np.float32)), (name, t, e)
            assert np.array_equal(adj[e], o["adj_matrix"].astype(np.float32)), (name, t, e)
            assert done[e] == d
            frozen[e] = d
    success, flow = env.computeMetrics()
    for e in range(E):
        s, f = oracles[e].metrics()
        assert abs(float(success[e]) - s) < 1e-7 and int(flow[e]) == f


def test_facade_matches_reference_known_answers():
    """tests/test_env.py:196-220 (action map) and :77-122 (three-way collision) through the facade."""
    import mapf_gnn_b200 as m
    goals = np.array([[9, 9], [8, 8], [7, 7], [6, 6], [5, 5]])
    starts = np.array([[5, 5], [4, 4], [3, 3], [2, 2], [1, 1]])
    env = m.GraphEnv(_cfg(5, 10, 20), goal=goals, starting_positions=starts.copy())
    obs, info, done, _ = env.step([0, 1, 2, 3, 4], np.ones((5, 1)))
    assert env.getPositions().tolist() == [[5, 5], [5, 4], [3, 4], [1, 2], [1, 0]]
    assert obs["fov"].shape == (5, 2, 5, 5) and obs["adj_matrix"].shape == (5, 5) and done is False
    assert obs["distances"] is obs["adj_matrix"]
    starts = np.array([[3, 4], [4, 5], [5, 4], [7, 7], [9, 9]])
    env = m.GraphEnv(_cfg(5, 10, 20), goal=goals, starting_positions=starts.copy())
    env.step([1, 4, 3, 0, 0], None)
    assert env.getPositions().tolist() == starts.tolist()
    assert env.computeMetrics()[0] == 0.0


def test_bad_arguments_raise():
    import mapf_gnn_b200 as m
    with pytest.raises(ValueError):
        m.BatchedGraphEnv(_cfg(5, 10), np.zeros((2, 4, 2)), np.zeros((2, 4, 2)))
    with pytest.raises(RuntimeError):
        m.BatchedGraphEnv(_cfg(200, 10), np.zeros((1, 200, 2)), np.zeros((1, 200, 2)))   # N > 128 unsupported


@pytest.mark.parametrize("E,N,H,M", [(64, 5, 18, 5), (33, 20, 32, 40), (16, 64, 64, 200), (200, 6, 4, 4), (8, 1, 2, 3)])
def test_device_scenario_generation(E, N, H, M):
    """Validity properties of the on-device generator (distributional parity with create_obstacles / create_goals):
    in-board, obstacles unique, goals unique and obstacle-free, starts unique and obstacle-free, deterministic in the
    seed, different across episodes and seeds; roughly uniform cell usage."""
    import mapf_gnn_b200 as m
    s, g, o = (t.cpu().numpy() for t in m.make_scenarios_device(E, N, H, M, seed=7))
    s2, g2, o2 = (t.cpu().numpy() for t in m.make_scenarios_device(E, N, H, M, seed=7))
    s3, g3, o3 = (t.cpu().numpy() for t in m.make_scenarios_device(E, N, H, M, seed=8))
    assert np.array_equal(s, s2) and np.array_equal(g, g2) and np.array_equal(o, o2)
    assert not (np.array_equal(g, g3) and np.array_equal(s, s3))
    for arr in (s, g, o):
        assert arr.min() >= 0 and arr.max() < H
    for e in range(E):
        so = set(map(tuple, o[e]))
        assert len(so) == M
        for pts in (g[e], s[e]):
            sp = set(map(tuple, pts))
            assert len(sp) == N and not (sp & so)
    if E >= 33:
        assert len({tuple(map(tuple, g[e])) for e in range(E)}) > E // 2
    if E * N >= 300 and H * H <= 400:
        counts = np.bincount(g[:, :, 1].ravel() * H + g[:, :, 0].ravel(), minlength=H * H)
        assert counts.max() <= 12 * max(1.0, E * N / (H * H))


def test_device_scenarios_drive_a_rollout():
    import mapf_gnn_b200 as m
    with pytest.raises(ValueError):
        m.make_scenarios_device(4, 10, 3, 5)
    s, g, o = m.make_scenarios_device(50, 5, 12, 6, seed=1)
    env = m.BatchedGraphEnv(_cfg(5, 12), g, s, o, sensing_range=6)
    orc = [eo.EnvOracle(5, 12, g[e].cpu().numpy(), s[e].cpu().numpy(), o[e].cpu().numpy(), sensing_range=6) for e in range(50)]
    for e in range(50):
        ob = orc[e].observe()
        assert np.array_equal(env.fov[e].cpu().numpy(), ob["fov"].astype(np.float32))


@pytest.mark.parametrize("E,N,H,M,r", [(67, 5, 18, 5, 6), (41, 20, 32, 40, 6), (9, 64, 64, 200, 6), (13, 33, 24, 20, 5),
                                       (5, 100, 40, 30, 5), (130, 2, 6, 3, 3)])
def test_step_decomposition_independent(E, N, H, M, r):
    """mapf_env_step_args.episodes_per_cta only changes the work decomposition (and whether a CTA's output slices are
    16-byte aligned, i.e. leave as TMA bulk stores or as the scalar copy): every value gives the oracle's result, and
    nothing is written behind the observation tensors (NaN canaries)."""
    import mapf_gnn_b200 as m
    rng = np.random.default_rng(100 + N)
    starts, goals, obst = eo.make_scenarios(rng, E, N, H, M)
    lanes = 8
    while lanes < N:
        lanes *= 2
    actions = rng.integers(0, 5, size=(3, E, N))
    oracles = [eo.EnvOracle(N, H, goals[e], starts[e], obst[e], sensing_range=r) for e in range(E)]
    ref = []
    for t in range(3):
        ref.append([oracles[e].step(actions[t][e])[0] for e in range(E)])
    for epb in sorted({0, 1, 2, 3, 4, 5, 128 // lanes}):
        if epb > 128 // lanes:
            continue
        env = m.BatchedGraphEnv(_cfg(N, H), goals, starts, obst, sensing_range=r)
        env.episodes_per_cta = epb
        fbuf = torch.full((E * N * 50 + 256,), float("nan"), device="cuda")
        abuf = torch.full((E * N * N + 256,), float("nan"), device="cuda")
        env.fov = fbuf[:E * N * 50].view(E, N, 2, 5, 5)
        env.adj_matrix = abuf[:E * N * N].view(E, N, N)
        for t in range(3):
            env.step(actions[t], freeze_done=False)
            fov, adj = env.fov.cpu().numpy(), env.adj_matrix.cpu().numpy()
            for e in range(E):
                assert np.array_equal(fov[e], ref[t][e]["fov"].astype(np.float32)), (epb, t, e)
                assert np.array_equal(adj[e], ref[t][e]["adj_matrix"].astype(np.float32)), (epb, t, e)
        assert torch.isnan(fbuf[E * N * 50:]).all() and torch.isnan(abuf[E * N * N:]).all(), epb
    env = m.BatchedGraphEnv(_cfg(N, H), goals, starts, obst, sensing_range=r)
    env.episodes_per_cta = 128 // lanes + 1
    with pytest.raises(RuntimeError):
        env.observe()


def test_facade_keeps_stepping_after_done(env_cases):
    """The reference GraphEnv never freezes itself (env_graph_gridv1.py:183-198: only its callers break on done): the
    facade must follow the golden trace THROUGH and PAST the step at which all agents are on their goals."""
    import mapf_gnn_b200 as m
    c = env_cases["goalrun_3a_8"]
    n, h, mm, T, max_time = (int(v) for v in c["meta"])
    env = m.GraphEnv(_cfg(n, h, max_time), goal=c["goals"], starting_positions=c["starts"].copy(),
                     obstacles=c["obstacles"] if mm else None, sensing_range=float(c["range"][0]))
    first = int(np.argmax(c["done"]))
    assert first + 1 < T, "the golden case must continue after done"
    for t in range(T):
        obs, _, done, _ = env.step(c["actions"][t], None)
        assert np.array_equal(env.getPositions(), c["pos"][t + 1]), t
        assert np.array_equal(obs["fov"], c["fov"][t + 1]), t
        assert np.array_equal(obs["adj_matrix"], c["adj"][t + 1]), t
        assert done == bool(c["done"][t]), t
        assert env.time == t + 1


def test_facade_rule_methods_and_writable_positions():
    """check_collision_obstacle / check_boundary / check_collisions / _computeDistance / computeMetrics on host-written
    positionX / positionY (the attributes tests/test_env.py:97-121,164-194 poke), against the oracle's closed forms."""
    import mapf_gnn_b200 as m
    rng = np.random.default_rng(11)
    N, H = 12, 9
    obstacles = np.array([[2, 3], [4, 4], [6, 1], [0, 8]])
    goals = eo.make_scenarios(rng, 1, N, H, 0)[1][0]
    env = m.GraphEnv(_cfg(N, H), goal=goals, starting_positions=eo.make_scenarios(rng, 1, N, H, 0)[0][0],
                     obstacles=obstacles, sensing_range=4)
    obst_set = set(map(tuple, obstacles.tolist()))
    for trial in range(20):
        temp = rng.integers(0, H, size=(N, 2))
        cand = temp + rng.integers(-2, 3, size=(N, 2))
        if trial % 3 == 0:
            cand[: N // 2] = cand[0]                      # many agents in one cell
        # obstacle rule
        env.positionX_temp, env.positionY_temp = temp[:, 0].copy(), temp[:, 1].copy()
        env.positionX, env.positionY = cand[:, 0].copy(), cand[:, 1].copy()
        env.check_collision_obstacle()
        want = np.array([temp[i] if tuple(cand[i]) in obst_set else cand[i] for i in range(N)])
        assert np.array_equal(env.getPositions(), want)
        # boundary rule
        env.check_boundary()
        want = np.clip(want, 0, H - 1)
        assert np.array_equal(env.getPositions(), want)
        # collision rule: every member of a shared cell reverts (single pass)
        env.check_collisions()
        keys = [tuple(p) for p in want]
        want2 = np.array([temp[i] if keys.count(keys[i]) > 1 else want[i] for i in range(N)])
        assert np.array_equal(env.getPositions(), want2)
        # adjacency of the host-written positions
        env._computeDistance()
        adj = eo.adjacency_closed_form(want2, eo.d2_limit_from_range(4))
        assert np.array_equal(env.adj_matrix, adj.astype(np.float64))
        assert env.distance_matrix is env.adj_matrix
        sr, _ = env.computeMetrics()
        assert abs(sr - float(np.mean(np.all(want2 == goals, axis=1)))) < 1e-7


def test_device_scenarios_distribution_matches_reference_generator():
    """Distributional parity of the on-device generator with the reference's numpy generator (create_obstacles :504-529,
    create_goals :465-501, evaluate.py:207-208 order) over 10^5 episodes: the reference draws obstacles uniformly without
    replacement, goals uniformly among the obstacle-free cells and (our explicit starts) likewise, so every cell marginal
    is uniform.  Chi-square goodness of fit against the uniform law for the obstacle / goal / start cells, for single
    slots (no slot is special), and two-sample chi-square of joint statistics (goal-start Manhattan distance of one agent,
    goal-obstacle distance) against samples of the numpy generator itself.  Critical values at p = 1e-5."""
    from scipy.stats import chi2
    import mapf_gnn_b200 as m
    E, N, H, M = 120_000, 5, 10, 8
    cells = H * H
    s, g, o = (t.cpu().numpy().astype(np.int64) for t in m.make_scenarios_device(E, N, H, M, seed=123))

    def gof(points, what):
        counts = np.bincount((points[..., 1] * H + points[..., 0]).ravel(), minlength=cells).astype(np.float64)
        exp = counts.sum() / cells
        stat = ((counts - exp) ** 2 / exp).sum()
        assert stat < chi2.ppf(1 - 1e-5, cells - 1), (what, stat)

    gof(o, "obstacles"), gof(g, "goals"), gof(s, "starts")
    for slot in range(N):
        gof(g[:, slot], f"goal slot {slot}"), gof(s[:, slot], f"start slot {slot}")
    gof(o[:, 0], "obstacle slot 0"), gof(o[:, M - 1], f"obstacle slot {M - 1}")
    # numpy generator of the reference's scheme (vectorised restatement in the oracle: make_scenarios)
    rs, rg, ro = (a.astype(np.int64) for a in eo.make_scenarios(np.random.default_rng(9), E, N, H, M))

    def two_sample(a, b, bins, what):
        ca = np.bincount(a, minlength=bins).astype(np.float64)
        cb = np.bincount(b, minlength=bins).astype(np.float64)
        keep = (ca + cb) >= 20
        ca, cb = ca[keep], cb[keep]
        k1, k2 = np.sqrt(cb.sum() / ca.sum()), np.sqrt(ca.sum() / cb.sum())
        stat = ((k1 * ca - k2 * cb) ** 2 / (ca + cb)).sum()
        assert stat < chi2.ppf(1 - 1e-5, keep.sum() - 1), (what, stat)

    man = lambda p, q: np.abs(p - q).sum(axis=-1)
    two_sample(man(g[:, 0], s[:, 0]), man(rg[:, 0], rs[:, 0]), 2 * H, "goal-start distance of agent 0")
    two_sample(man(g[:, 1], g[:, 3]), man(rg[:, 1], rg[:, 3]), 2 * H, "goal-goal distance")
    two_sample(man(g[:, 2], o[:, 5]), man(rg[:, 2], ro[:, 5]), 2 * H, "goal-obstacle distance")
    two_sample(man(s[:, 4], o[:, 0]), man(rs[:, 4], ro[:, 0]), 2 * H, "start-obstacle distance")
    # exclusions hold in every episode (vectorised): obstacles unique, goals / starts unique and obstacle-free
    key = lambda p: p[..., 1] * H + p[..., 0]
    ko, kg, ks = np.sort(key(o), axis=1), np.sort(key(g), axis=1), np.sort(key(s), axis=1)
    assert (np.diff(ko, axis=1) > 0).all() and (np.diff(kg, axis=1) > 0).all() and (np.diff(ks, axis=1) > 0).all()
    assert not (key(g)[:, :, None] == key(o)[:, None, :]).any() and not (key(s)[:, :, None] == key(o)[:, None, :]).any()
