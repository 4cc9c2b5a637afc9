"""Pin the env oracle to the reference: golden traces + the reference's own known-answer tests."""
import numpy as np
import pytest

from oracle import env_oracle as eo


def _obst_mask(case):
    n, h = int(case["meta"][0]), int(case["meta"][1])
    mask = np.zeros((h, h), dtype=bool)
    if len(case["obstacles"]):
        mask[case["obstacles"][:, 1], case["obstacles"][:, 0]] = True
    return mask


def test_all_cases_present(env_cases):
    assert len(env_cases) == 12


def test_structural_oracle_matches_reference_traces(env_cases):
    for name, c in env_cases.items():
        n, h, m, T, max_time = (int(v) for v in c["meta"])
        env = eo.EnvOracle(n, h, c["goals"], c["starts"], c["obstacles"] if m else None,
                           sensing_range=float(c["range"][0]), max_time=max_time)
        obs = env.observe()
        assert np.array_equal(obs["fov"], c["fov"][0]), name
        assert np.array_equal(obs["adj_matrix"], c["adj"][0]), name
        for t in range(T):
            obs, done = env.step(c["actions"][t])
            assert np.array_equal(env.positions, c["pos"][t + 1]), (name, t)
            assert np.array_equal(obs["fov"], c["fov"][t + 1]), (name, t)
            assert np.array_equal(obs["adj_matrix"], c["adj"][t + 1]), (name, t)
            assert done == bool(c["done"][t]), (name, t)
        assert np.array_equal(env.board.astype(np.uint8), c["board_final"]), name
        success, flow = env.metrics()
        assert success == c["metrics"][0] and flow == c["metrics"][1], name


def test_closed_forms_match_reference_traces(env_cases):
    for name, c in env_cases.items():
        n, h, m, T, _ = (int(v) for v in c["meta"])
        mask = _obst_mask(c)
        lim = eo.d2_limit_from_range(float(c["range"][0]))
        board = np.where(mask, eo.BOARD_OBSTACLE, eo.BOARD_EMPTY).astype(np.uint8)
        pos = c["starts"].astype(np.int64)
        assert np.array_equal(eo.adjacency_closed_form(pos, lim), c["adj"][0]), name
        assert np.array_equal(eo.fov_closed_form(board, pos, c["goals"]), c["fov"][0]), name
        for t in range(T):
            new = eo.move_closed_form(pos, c["actions"][t], mask, h)
            assert np.array_equal(new, c["pos"][t + 1]), (name, t)
            board[pos[:, 1], pos[:, 0]] = eo.BOARD_EMPTY
            board[new[:, 1], new[:, 0]] = eo.BOARD_AGENT
            pos = new
            assert np.array_equal(eo.adjacency_closed_form(pos, lim), c["adj"][t + 1]), (name, t)
            assert np.array_equal(eo.fov_closed_form(board, pos, c["goals"]), c["fov"][t + 1]), (name, t)
        assert np.array_equal(board, c["board_final"]), name


def test_some_episode_finishes(env_cases):
    assert env_cases["goalrun_3a_8"]["done"].any()


@pytest.mark.parametrize("r,expect", [(6, 36), (4, 16), (2.5, 7), (2, 4), (1, 1), (0, 0), (1.5, 3)])
def test_d2_limit(r, expect):
    assert eo.d2_limit_from_range(r) == expect


# --- known-answer cases of the reference's own tests/test_env.py ---------------------------
def test_ref_action_map():
    # tests/test_env.py:196-220
    goals = np.array([[9, 9], [8, 8], [7, 7], [6, 6], [5, 5]])
    starts = np.array([[5, 5], [4, 4], [3, 3], [2, 2], [1, 1]])
    env = eo.EnvOracle(5, 10, goals, starts)
    env.step([0, 1, 2, 3, 4])
    assert env.positions.tolist() == [[5, 5], [5, 4], [3, 4], [1, 2], [1, 0]]
    mask = np.zeros((10, 10), dtype=bool)
    assert eo.move_closed_form(starts, [0, 1, 2, 3, 4], mask, 10).tolist() == env.positions.tolist()


def test_ref_three_agent_collision_reverts_all():
    # tests/test_env.py:77-122: three agents step into (4,4); all three revert
    starts = np.array([[3, 4], [4, 5], [5, 4], [7, 7], [9, 9]])
    mask = np.zeros((10, 10), dtype=bool)
    new = eo.move_closed_form(starts, [1, 4, 3, 0, 0], mask, 10)
    assert new.tolist() == starts.tolist()


def test_ref_obstacle_revert_and_clamp():
    # tests/test_env.py:149-194
    mask = np.zeros((10, 10), dtype=bool)
    mask[5, 5] = True
    starts = np.array([[4, 5], [0, 3], [9, 9], [5, 0]])
    new = eo.move_closed_form(starts, [1, 3, 2, 4], mask, 10)
    assert new.tolist() == [[4, 5], [0, 3], [9, 9], [5, 0]]


def test_ref_success_rate():
    # tests/test_env.py:124-147
    goals = np.array([[5, 5], [6, 6], [7, 7], [8, 8], [9, 9]])
    env = eo.EnvOracle(5, 10, goals, np.array([[5, 5], [6, 6], [2, 2], [3, 3], [4, 4]]), max_time=20)
    assert env.metrics() == (0.4, 100)


def test_ref_distance_facts():
    # tests/test_env.py:252-298 with sensing_range=2: unit neighbours kept, sqrt(2) pair kept (< 2)
    adj = eo.adjacency_closed_form(np.array([[0, 0], [1, 0], [0, 1]]), eo.d2_limit_from_range(2))
    assert adj.tolist() == [[0, 1, 1], [1, 0, 1], [1, 1, 0]]
    assert np.all(np.diag(adj) == 0)
