"""Parity at BASELINE.json's full single-GPU sizes through size-independent properties (the oracle cannot step
thousands of episodes in test time): a policy-driven rollout of the whole batch, with
  * invariants checked on EVERY episode and step (rules of grid/env_graph_gridv1.py:200-313, 117-136, 218-265),
  * a random sample of episodes replayed through the numpy oracle with the same actions (bit-exact),
  * batch-composition independence: the sampled episodes stepped alone, teacher-forced, reproduce the same bits,
  * logits of the sampled episodes against the oracle network at every step (1e-4 relative)."""
import numpy as np
import pytest
import torch

from helpers import load_network, model_config, oracle_params, rel_err
from oracle import env_oracle as eo
from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu
REL = 1e-4   # fp32 tolerance of the north star: max|a-b| / max|b|


def _invariants(env, obst_mask, H, N, r2, first):
    """obst_mask: bool [E, H, H] indexed [y, x]."""
    pos = env.pos.cpu().numpy()
    fov = env.fov.cpu().numpy()
    adj = env.adj_matrix.cpu().numpy()
    E = pos.shape[0]
    assert pos.min() >= 0 and pos.max() < H                                   # check_boundary :267-271
    e_idx = np.repeat(np.arange(E), N)
    assert not obst_mask[e_idx, pos[:, :, 1].ravel(), pos[:, :, 0].ravel()].any()   # check_collision_obstacle :303-313
    # (two agents CAN end a step on one cell: the collision revert of :282-301 is a single pass)
    d2 = ((pos[:, :, None, :] - pos[:, None, :, :]).astype(np.int64) ** 2).sum(-1)
    assert set(np.unique(adj)) <= {0.0, 1.0}
    assert (adj[:, np.arange(N), np.arange(N)] == 0).all()                       # :117-136: no self edges
    assert not (adj > 0)[d2 >= r2].any()                                         # only inside the sensing range
    in_range = ((d2 > 0) & (d2 < r2)).sum(-1)
    assert ((adj > 0).sum(-1) == np.minimum(in_range, 4))[in_range <= 4].all()   # closest-4 rule without ties
    assert ((adj > 0).sum(-1) >= 4)[in_range > 4].all()
    assert ((fov[:, :, 1] == 3).sum(axis=(2, 3)) == 1).all()                     # map_goal :218-246: one goal marker
    assert set(np.unique(fov[:, :, 0])) <= {0.0, 1.0, 2.0}
    if not first:
        assert (fov[:, :, 0, 2, 2] == 1).all()                                   # self at the FOV centre (Appendix B.3)


@pytest.mark.parametrize("name,E,N,H,M,r,T,sample", [("gnn_k2", 4096, 5, 28, 8, 6, 12, 24),       # C2a
                                                      ("gnn_k4_seed4", 8192, 64, 64, 200, 8, 2, 3)])  # C5, one GPU's share
def test_full_size_rollout_properties(name, E, N, H, M, r, T, sample):
    import mapf_gnn_b200 as m
    rng = np.random.default_rng(E + N)
    starts, goals, obst = m.make_scenarios(rng, E, N, H, M)
    cfg = model_config(name, N, board_size=[H, H], obstacles=M, sensing_range=r)
    net = load_network(name, num_agents=N).eval()
    params = oracle_params(name)
    env = m.BatchedGraphEnv(cfg, goals, starts, obst, sensing_range=r)
    ro = m.Rollout(net, env)
    obst_mask = np.zeros((E, H, H), dtype=bool)
    obst_mask[np.repeat(np.arange(E), M), obst[:, :, 1].ravel(), obst[:, :, 0].ravel()] = True
    pick = np.sort(rng.choice(E, size=sample, replace=False))
    oracles = [eo.EnvOracle(N, H, goals[e], starts[e], obst[e], sensing_range=r) for e in pick]
    sub = m.BatchedGraphEnv(cfg, goals[pick], starts[pick], obst[pick], sensing_range=r)
    _invariants(env, obst_mask, H, N, r * r, first=True)
    prev_done = np.zeros(E, dtype=bool)
    prev_time = env.time.cpu().numpy().copy()
    for t in range(T):
        fov_in, adj_in = env.fov[pick].cpu(), env.adj_matrix[pick].cpu()
        ro.run(1)
        acts = ro.actions.cpu().numpy()
        with torch.no_grad():
            ref = mo.forward(params, fov_in.float(), adj_in.float()).numpy()
        assert rel_err(ro.logits[pick].cpu().numpy(), ref) < REL, (name, t)
        _invariants(env, obst_mask, H, N, r * r, first=False)
        done, time = env.done.cpu().numpy().astype(bool), env.time.cpu().numpy()
        assert (done | ~prev_done).all()                                       # done is monotone
        assert (time - prev_time == (~prev_done).astype(time.dtype)).all()      # frozen episodes do not tick
        at_goal = (env.pos.cpu().numpy() == goals).all(axis=(1, 2))
        assert (done == (at_goal | prev_done)).all()                            # checkAllInGoal :156-159
        # sampled episodes: oracle replay and the same episodes stepped alone, both with the batch's actions
        sub.step(acts[pick], freeze_done=True)
        pos, fov, adj = env.pos.cpu().numpy(), env.fov.cpu().numpy(), env.adj_matrix.cpu().numpy()
        assert np.array_equal(sub.pos.cpu().numpy(), pos[pick]) and np.array_equal(sub.fov.cpu().numpy(), fov[pick])
        assert np.array_equal(sub.adj_matrix.cpu().numpy(), adj[pick])
        for i, e in enumerate(pick):
            if prev_done[e]:
                continue
            o, d = oracles[i].step(acts[e])
            assert np.array_equal(pos[e], oracles[i].positions), (name, t, e)
            assert np.array_equal(fov[e], o["fov"].astype(np.float32)), (name, t, e)
            assert np.array_equal(adj[e], o["adj_matrix"].astype(np.float32)), (name, t, e)
            assert bool(done[e]) == bool(d)
        prev_done, prev_time = done.copy(), time.copy()
    success, flow = env.computeMetrics()
    for i, e in enumerate(pick):
        s, f = oracles[i].metrics()
        assert abs(float(success[e]) - s) < 1e-7 and int(flow[e]) == f
