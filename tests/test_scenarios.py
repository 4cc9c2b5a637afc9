"""Host scenario generator (evaluate.py:207-208 order; train.py:111-112 validation order) -- CPU only."""
import numpy as np
import pytest

from mapf_gnn_b200.scenarios import make_scenarios


def _cells(a):
    return set(map(tuple, a))


@pytest.mark.parametrize("E,N,H,M", [(200, 5, 18, 5), (50, 20, 8, 20), (30, 1, 2, 3), (10, 64, 64, 300)])
def test_evaluate_order_is_valid(E, N, H, M):
    s, g, o = make_scenarios(np.random.default_rng(E), E, N, H, M)
    assert s.dtype == g.dtype == o.dtype == np.int32 and s.shape == (E, N, 2) and o.shape == (E, M, 2)
    for e in range(E):
        so = _cells(o[e])
        assert len(so) == M and len(_cells(g[e])) == N and len(_cells(s[e])) == N
        assert not (_cells(g[e]) & so) and not (_cells(s[e]) & so)
    assert min(s.min(), g.min(), o.min()) >= 0 and max(s.max(), g.max(), o.max()) < H


def test_validation_order_draws_goals_independently():
    s, g, o = make_scenarios(np.random.default_rng(0), 400, 5, 6, 12, goals_ignore_obstacles=True)
    under = sum(bool(_cells(g[e]) & _cells(o[e])) for e in range(400))
    assert under > 200                                   # 12 of 36 cells blocked: most episodes have a buried goal
    assert all(len(_cells(g[e])) == 5 and not (_cells(s[e]) & _cells(o[e])) for e in range(400))


def test_too_small_board_raises_like_create_goals():
    with pytest.raises(ValueError):
        make_scenarios(np.random.default_rng(0), 4, 10, 3, 5)


def test_round_aligned_group_sizes():
    """Host logic of PipelinedHostRollout: groups whose encoder launches are whole rounds of the persistent kernel."""
    from mapf_gnn_b200.rollout import PipelinedHostRollout as P
    assert P.round_aligned_sizes(4096, 5, 2, 148) == [1894, 2202]          # 1894 * 5 = 9470 images = 2 rounds of 4736
    assert P.round_aligned_sizes(90, 5, 3, 148) == [30, 30, 30]            # too small for a round: equal split
    for E, N, G in [(4096, 5, 3), (8192, 64, 2), (4096, 20, 4), (7, 3, 7), (100000, 5, 8)]:
        sizes = P.round_aligned_sizes(E, N, G, 148)
        assert len(sizes) == G and sum(sizes) == E and min(sizes) >= 1
