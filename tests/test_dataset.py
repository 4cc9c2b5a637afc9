"""Dataset wire format: loader logic on CPU (files written in the reference layout), cross-checked against the
reference's own CreateDataset when /root/reference is present; recording parity on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import env_oracle as eo


def _write_cases(root, rng, n_cases, N, lengths):
    os.makedirs(root, exist_ok=True)
    for i, T in enumerate(lengths[:n_cases]):
        d = os.path.join(root, f"case_{i}")
        os.makedirs(d)
        np.save(os.path.join(d, "states.npy"), rng.integers(0, 4, size=(T, N, 2, 5, 5)).astype(np.float64))
        np.save(os.path.join(d, "gso.npy"), (rng.random((T, N, N, N)) < 0.3).astype(np.float64))
        np.save(os.path.join(d, "trajectory_record.npy"), rng.integers(0, 5, size=(N, T)).astype(np.float64))
    os.makedirs(os.path.join(root, "case_without_files"))


def test_loader_matches_reference_semantics(tmp_path):
    from mapf_gnn_b200.dataset import GNNDataLoader, load_cases
    rng = np.random.default_rng(0)
    N, min_time = 5, 6
    root = str(tmp_path / "ds" / "train")
    _write_cases(root, rng, 7, N, [12, 9, 4, 30, 8, 7, 10])     # one too short (4), all within max_time_dl
    states, targets, gso, count = load_cases(root, min_time, N, 25)
    assert states.shape[1:] == (N, 2, 5, 5) and targets.shape[1] == N and gso.shape[1:] == (N, N)
    assert states.shape[0] == count * min_time == targets.shape[0] == gso.shape[0]
    assert count == 6
    ref_path = "/root/reference"
    if os.path.isdir(ref_path):
        sys.path.insert(0, ref_path)
        try:
            from data_loader import CreateDataset
        finally:
            sys.path.remove(ref_path)
        cfg = {"train": {"root_dir": str(tmp_path / "ds"), "min_time": min_time, "nb_agents": N, "max_time_dl": 25}}
        ds = CreateDataset(cfg, "train")
        assert ds.count == count
        assert np.array_equal(ds.states.astype(np.float32), states.numpy())
        assert np.array_equal(ds.trajectories.astype(np.float32), targets.numpy())
        assert np.array_equal(ds.gsos.astype(np.float32), gso.numpy())
    cfg = {"train": {"root_dir": str(tmp_path / "ds"), "min_time": min_time, "nb_agents": N, "max_time_dl": 25},
           "batch_size": 16}
    loader = GNNDataLoader(cfg, "train", device="cpu", seed=1)
    seen = 0
    for s, t, g in loader.train_loader:
        assert s.shape[0] == t.shape[0] == g.shape[0] <= 16 and s.dtype == torch.float32
        diag_ok = g.diagonal(dim1=1, dim2=2).min(dim=1).values >= 1.0          # + I (data_loader.py:74)
        hole = g.abs().sum(dim=(1, 2)) == 0     # the reference's skipped-case holes stay zero (reproduced quirk)
        assert bool((diag_ok | hole).all())
        seen += s.shape[0]
    assert seen == count * min_time and len(loader) == (seen + 15) // 16
    # reference_len=True: the reference's Dataset.__len__ returns the case count (data_loader.py:112-113), so its
    # DataLoader draws only flattened sample indices below `count` and runs ceil(count / batch) iterations per epoch
    ref_loader_ = GNNDataLoader(cfg, "train", device="cpu", seed=1, reference_len=True)
    assert len(ref_loader_) == (count + 15) // 16
    rows = torch.cat([s for s, _, _ in ref_loader_])
    assert rows.shape[0] == count
    first = ref_loader_.states[:count]
    assert sorted(map(lambda r: r.numpy().tobytes(), rows)) == sorted(map(lambda r: r.numpy().tobytes(), first))
    if os.path.isdir("/root/reference"):
        # the reference's own DataLoader over its Dataset object visits exactly `count` rows per epoch
        from torch.utils.data import DataLoader
        assert len(ds) == count
        assert sum(b[0].shape[0] for b in DataLoader(ds, batch_size=16, shuffle=True)) == count


@pytest.mark.gpu
def test_record_cases_matches_oracle_replay():
    """record.py:66-88 on the GPU vs the numpy oracle env replaying the same expert actions (incl. the quirks:
    first action dropped, last slot overwritten, no freezing after done)."""
    import mapf_gnn_b200 as m
    from mapf_gnn_b200.dataset import record_cases
    E, N, H, M, T = 12, 5, 10, 4, 9
    rng = np.random.default_rng(4)
    starts, goals, obst = m.make_scenarios(rng, E, N, H, M)
    traj = rng.integers(0, 5, size=(E, N, T + 1))
    traj[0, :, 3:] = 0
    cfg = {"num_agents": N, "board_size": [H, H], "max_time": 32, "min_time": 1}
    states, gso, rec = record_cases(cfg, goals, starts, obst, traj, sensing_range=4)
    assert states.shape == (E, T, N, 2, 5, 5) and gso.shape == (E, T, N, N, N) and rec.shape == (E, N, T)
    for e in range(E):
        env = eo.EnvOracle(N, H, goals[e], starts[e], obst[e], sensing_range=4)
        obs = env.observe()
        exp_s, exp_a = np.zeros((T, N, 2, 5, 5)), np.zeros((T, N, N))
        for i in range(T):
            exp_s[i], exp_a[i] = obs["fov"], obs["adj_matrix"]
            obs, _ = env.step(traj[e, :, 1 + i])
        exp_s[T - 1], exp_a[T - 1] = obs["fov"], obs["adj_matrix"]
        assert np.array_equal(states[e], exp_s), e
        for a in range(N):
            assert np.array_equal(gso[e, :, a], exp_a), e
        assert np.array_equal(rec[e], traj[e, :, 1:])


@pytest.mark.gpu
def test_generate_dataset_end_to_end(tmp_path):
    """main_data.py:24-27 in one call: CBS plans (native, batch) -> action matrices -> GPU recording -> the reference's
    files; the recordings equal the numpy oracle env replaying each case's expert actions, the expert actions lead every
    agent to its goal, and the loader serves the cases."""
    import yaml
    from mapf_gnn_b200.dataset import GNNDataLoader, generate_dataset
    N, H, M = 4, 12, 6
    cfg = {"num_agents": N, "nb_agents": N, "map_shape": [H, H], "board_size": [H, H], "nb_obstacles": M, "sensor_range": 4,
           "max_time": 32, "min_time": 3}
    np.random.seed(3)
    root = tmp_path / "train"
    count = generate_dataset(str(root), 16, cfg, threads=2)
    assert 10 <= count <= 16
    for i in range(count):
        d = root / f"case_{i}"
        inp = yaml.safe_load(open(d / "input.yaml"))
        sol = yaml.safe_load(open(d / "solution.yaml"))
        traj, rec = np.load(d / "trajectory.npy"), np.load(d / "trajectory_record.npy")
        states, gso = np.load(d / "states.npy"), np.load(d / "gso.npy")
        T = traj.shape[1] - 1
        assert sol["cost"] == sum(len(p) for p in sol["schedule"].values())
        assert np.array_equal(rec, traj[:, 1:]) and states.shape == (T, N, 2, 5, 5) and gso.shape == (T, N, N, N)
        goals = np.array([a["goal"] for a in inp["agents"]])
        starts = np.array([a["start"] for a in inp["agents"]])
        assert np.array_equal(starts, np.load(d / "startings.npy"))
        env = eo.EnvOracle(N, H, goals, starts, np.array(inp["map"]["obstacles"]).reshape(-1, 2), sensing_range=4)
        full = eo.EnvOracle(N, H, goals, starts, np.array(inp["map"]["obstacles"]).reshape(-1, 2), sensing_range=4)
        for k in range(traj.shape[1]):
            full.step(traj[:, k])
        assert np.array_equal(np.stack([full.px, full.py], axis=1), goals)      # the expert arrives
        obs = env.observe()
        for k in range(T):
            if k < T - 1:
                assert np.array_equal(states[k], obs["fov"]) and np.array_equal(gso[k, 0], obs["adj_matrix"])
            obs, _ = env.step(traj[:, 1 + k])
        if T > 0:
            assert np.array_equal(states[T - 1], obs["fov"])
    loader = GNNDataLoader({**cfg, "batch_size": 8, "train": {"root_dir": str(tmp_path), "mode": "train", "min_time": 3,
                                                              "max_time_dl": 25, "nb_agents": N}}, device="cuda")
    assert loader.num_samples > 0
    states, actions, gso = next(iter(loader))
    assert states.shape[1:] == (N, 2, 5, 5) and actions.shape[1:] == (N,) and gso.shape[1:] == (N, N) and states.is_cuda


@pytest.mark.gpu
def test_record_env_over_a_case_directory(tmp_path):
    """The three scripts of main_data.py:24-27 one after the other on a directory: create_solutions ->
    parse_dataset_trajectories -> record_env; the recordings equal generate_dataset's (same seed) file for file."""
    from mapf_gnn_b200 import cbs
    from mapf_gnn_b200.dataset import generate_dataset, record_env
    N, H, M = 3, 9, 4
    cfg = {"num_agents": N, "nb_agents": N, "map_shape": [H, H], "board_size": [H, H], "nb_obstacles": M, "sensor_range": 4,
           "max_time": 32, "min_time": 3}
    a, b = tmp_path / "a", tmp_path / "b"
    np.random.seed(11)
    cbs.create_solutions(str(a), 10, cfg)
    cbs.parse_dataset_trajectories(str(a))
    count = record_env(str(a), cfg)
    np.random.seed(11)
    assert generate_dataset(str(b), 10, cfg) == count > 5
    names_a = sorted(d for d in os.listdir(a) if d.startswith("case_"))
    names_b = sorted(d for d in os.listdir(b) if d.startswith("case_"))
    assert len(names_a) == len(names_b) == count
    # create_solutions numbers the cases by their draw (failed ones leave holes), generate_dataset consecutively
    for da, db in zip(sorted(names_a, key=lambda d: int(d.split("_")[1])), sorted(names_b, key=lambda d: int(d.split("_")[1]))):
        for f in ("trajectory.npy", "startings.npy", "states.npy", "gso.npy", "trajectory_record.npy"):
            assert np.array_equal(np.load(a / da / f), np.load(b / db / f)), (da, db, f)
    assert (a / "stats.txt").exists()
    assert record_env(str(a), cfg, skip_last_case=True) == count - 1
