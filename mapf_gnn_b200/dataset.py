"""Dataset wire format of the reference: recording expert trajectories into
``case_*/{states,gso,trajectory_record}.npy`` (``data_generation/record.py:39-97``) and loading them back the way
``data_loader.py:24-125`` does.  Recording is the batched GPU environment replaying expert actions for all cases at
once; loading is host-side numpy (file I/O), followed by pinned-memory uploads of whole batches.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .env import BatchedGraphEnv


@torch.no_grad()
def record_cases(config, goals, starts, obstacles, trajectories, sensing_range=6, device="cuda"):
    """Replay expert action matrices through the environment (record.py:66-88) for E cases of equal length.

    trajectories: int [E, N, T+1] expert actions per agent (``trajectory.npy``); the first action is dropped
    (record.py:67).  Returns numpy arrays in the on-disk layout: states f64 [E, T, N, 2, 5, 5], gso f64
    [E, T, N, N, N] (the adjacency repeated along the first agent axis, record.py:80), trajectory_record [E, N, T].
    Quirk kept (record.py:87-88): the observation after the last action overwrites the last slot."""
    traj = np.asarray(trajectories)[:, :, 1:]
    E, N, T = traj.shape
    env = BatchedGraphEnv(config, goals, starts, obstacles, sensing_range=sensing_range, device=device)
    states = torch.empty((T, E, N, 2, 5, 5), dtype=torch.float32, device=env.device)
    adj = torch.empty((T, E, N, N), dtype=torch.float32, device=env.device)
    actions = torch.as_tensor(traj, device=env.device).to(torch.int32)
    for i in range(T):
        states[i].copy_(env.fov)
        adj[i].copy_(env.adj_matrix)
        env.step(actions[:, :, i].contiguous(), freeze_done=False)
    if T > 0:
        states[T - 1].copy_(env.fov)
        adj[T - 1].copy_(env.adj_matrix)
    states_np = states.permute(1, 0, 2, 3, 4, 5).cpu().numpy().astype(np.float64)
    adj_np = adj.permute(1, 0, 2, 3).cpu().numpy().astype(np.float64)
    gso_np = np.broadcast_to(adj_np[:, :, None, :, :], (E, T, N, N, N)).copy()
    return states_np, gso_np, traj


def generate_dataset(path, num_cases, config, threads=0, device="cuda"):
    """The reference's whole data-generation driver (data_generation/main_data.py:24-27): random instances -> CBS plans
    (``create_solutions``, dataset_gen.py:151-206: native batch search on the host cores, ``cbs.generate_cases``) ->
    action matrices (``parse_dataset_trajectories``, trajectory_parser.py:68-106) -> recorded observations
    (``record_env``, record.py:39-97: here ALL cases of one length replay together on the GPU).  Writes, per solved
    instance, ``case_<i>/{input.yaml, solution.yaml, trajectory.npy, startings.npy, states.npy, gso.npy,
    trajectory_record.npy}`` in the reference's formats and returns the number of cases.  ``config`` carries main_data.py's
    keys (map_shape, nb_agents, nb_obstacles, sensor_range, board_size, num_agents, max_time, min_time); instances are
    drawn from ``np.random`` (seed it).  Deviation, on purpose: record.py's loop bound ``len(cases) - 1`` leaves the
    last case without recordings; every case is recorded here."""
    import yaml
    from . import cbs
    cases = cbs.generate_cases(num_cases, config["map_shape"], config["nb_agents"], config["nb_obstacles"], threads=threads)
    os.makedirs(path, exist_ok=True)
    by_len = {}
    for i, c in enumerate(cases):
        case_dir = os.path.join(path, f"case_{i}")
        os.makedirs(case_dir, exist_ok=True)
        inp = {"agents": c["input"]["agents"],
               "map": {"dimensions": c["input"]["map"]["dimensions"], "obstacles": [list(o) for o in c["input"]["map"]["obstacles"]]}}
        with open(os.path.join(case_dir, "input.yaml"), "w") as fh:
            yaml.safe_dump(inp, fh)
        with open(os.path.join(case_dir, "solution.yaml"), "w") as fh:
            yaml.safe_dump({"schedule": c["schedule"], "cost": c["cost"]}, fh)
        np.save(os.path.join(case_dir, "trajectory.npy"), c["trajectory"])
        np.save(os.path.join(case_dir, "startings.npy"), c["startings"])
        by_len.setdefault(c["trajectory"].shape[1], []).append(i)
    for T, idx in by_len.items():
        if T < 2:            # a plan of one state has no action to replay (record.py would write empty arrays)
            for i in idx:
                n = cases[i]["trajectory"].shape[0]
                save_case(os.path.join(path, f"case_{i}"), np.zeros((0, n, 2, 5, 5)), np.zeros((0, n, n, n)),
                          cases[i]["trajectory"][:, 1:])
            continue
        goals = np.array([[a["goal"] for a in cases[i]["input"]["agents"]] for i in idx], dtype=np.int32)
        starts = np.array([[a["start"] for a in cases[i]["input"]["agents"]] for i in idx], dtype=np.int32)
        obst = np.array([cases[i]["input"]["map"]["obstacles"] for i in idx], dtype=np.int32).reshape(len(idx), -1, 2)
        traj = np.stack([cases[i]["trajectory"] for i in idx])
        states, gso, rec = record_cases(config, goals, starts, obst, traj, sensing_range=config["sensor_range"], device=device)
        for k, i in enumerate(idx):
            save_case(os.path.join(path, f"case_{i}"), states[k], gso[k], rec[k])
    return len(cases)


def record_env(path, config, device="cuda", skip_last_case=False):
    """``record_env`` of data_generation/record.py:39-97 over a directory of ``case_<i>/{input.yaml, trajectory.npy}``
    (written by ``cbs.create_solutions`` + ``cbs.parse_dataset_trajectories`` or by the reference's scripts): all cases
    of one trajectory length and obstacle count replay together on the GPU, then ``states.npy`` / ``gso.npy`` /
    ``trajectory_record.npy`` are written per case and ``stats.txt`` like the reference's.  ``skip_last_case=True``
    reproduces the reference's loop bound (``range(len(cases) - 1)``, record.py:44,59: the case with the highest
    number gets no recordings)."""
    import yaml
    cases = sorted((d for d in os.listdir(path) if d.startswith("case_") and os.path.isdir(os.path.join(path, d))),
                   key=lambda d: int(d.split("_")[1]))
    if skip_last_case:
        cases = cases[:-1]
    groups, lengths = {}, []
    for name in cases:
        tpath = os.path.join(path, name, "trajectory.npy")
        if not os.path.exists(tpath):
            continue
        with open(os.path.join(path, name, "input.yaml")) as fh:
            params = yaml.load(fh, Loader=yaml.FullLoader)
        traj = np.load(tpath, allow_pickle=True)
        goals = np.array([[int(a["goal"][0]), int(a["goal"][1])] for a in params["agents"]], dtype=np.int32)
        starts = np.array([[int(a["start"][0]), int(a["start"][1])] for a in params["agents"]], dtype=np.int32)
        obst = np.array([[int(o[0]), int(o[1])] for o in params["map"]["obstacles"]], dtype=np.int32).reshape(-1, 2)
        lengths.append(traj.shape[1])
        key = (traj.shape[0], traj.shape[1], obst.shape[0], int(params["map"]["dimensions"][0]))
        groups.setdefault(key, []).append((name, goals, starts, obst, traj))
    if lengths:
        with open(os.path.join(path, "stats.txt"), "w") as fh:
            fh.write(f"max steps {float(np.max(lengths))}\nmin steps {float(np.min(lengths))}\nmean steps {np.mean(lengths)}\n")
    for (n, t, m, board), items in groups.items():
        cfg = dict(config, num_agents=n, board_size=[board, board])
        if t < 2:
            for name, *_ in items:
                save_case(os.path.join(path, name), np.zeros((0, n, 2, 5, 5)), np.zeros((0, n, n, n)), np.zeros((n, 0), np.int32))
            continue
        states, gso, rec = record_cases(cfg, np.stack([i[1] for i in items]), np.stack([i[2] for i in items]),
                                        np.stack([i[3] for i in items]), np.stack([i[4] for i in items]),
                                        sensing_range=config["sensor_range"], device=device)
        for k, (name, *_) in enumerate(items):
            save_case(os.path.join(path, name), states[k], gso[k], rec[k])
    return sum(len(v) for v in groups.values())


def save_case(case_dir, states, gso, trajectory_record):
    """Write one case in the reference's file layout (record.py:90-94)."""
    os.makedirs(case_dir, exist_ok=True)
    np.save(os.path.join(case_dir, "states.npy"), states)
    np.save(os.path.join(case_dir, "gso.npy"), gso)
    np.save(os.path.join(case_dir, "trajectory_record.npy"), trajectory_record)


def load_cases(dir_path, min_time, nb_agents, max_time_dl):
    """``CreateDataset.__init__`` (data_loader.py:24-101): per case keep states[1:min_time+1], the first min_time
    expert actions and gso[:min_time, 0] + I; skip cases shorter than min_time or longer than max_time_dl; flatten
    time into the sample axis.  Like the reference, skipped cases leave holes that shift later cases out of the
    kept prefix (`states[:count]`, :95-97) -- reproduced, not fixed.
    Returns (states f32 [S,N,2,5,5], targets f32 [S,N], gso f32 [S,N,N], count)."""
    cases = os.listdir(dir_path)
    states = np.zeros((len(cases), min_time, nb_agents, 2, 5, 5))
    trajectories = np.zeros((len(cases), min_time, nb_agents))
    gsos = np.zeros((len(cases), min_time, nb_agents, nb_agents))
    count = 0
    for i, case in enumerate(cases):
        path = os.path.join(dir_path, case, "states.npy")
        if not os.path.exists(path):
            continue
        state = np.load(path)[1:min_time + 1]
        tray = np.load(os.path.join(dir_path, case, "trajectory_record.npy"))[:, :min_time]
        gso = np.load(os.path.join(dir_path, case, "gso.npy"))[:min_time, 0, :, :] + np.eye(nb_agents)
        if state.shape[0] < min_time or tray.shape[1] < min_time:
            continue
        if state.shape[0] > max_time_dl or tray.shape[1] > max_time_dl:
            continue
        states[i], trajectories[i], gsos[i] = state, tray.T, gso
        count += 1
    states = states[:count].reshape((-1, nb_agents, 2, 5, 5))
    trajectories = trajectories[:count].reshape((-1, nb_agents))
    gsos = gsos[:count].reshape((-1, nb_agents, nb_agents))
    return (torch.from_numpy(states).float(), torch.from_numpy(trajectories).float(), torch.from_numpy(gsos).float(),
            count)


class GNNDataLoader:
    """``GNNDataLoader(config).train_loader`` (data_loader.py:9-21): iterating yields
    ``(states[B,N,2,5,5], actions[B,N], gso[B,N,N])`` float32 batches.  The whole dataset is kept in pinned host
    memory and every batch is one asynchronous upload to ``device`` (the reference uses DataLoader workers +
    pin_memory; here a batch is a contiguous gather, so no worker processes are needed).

    ``reference_len=True`` reproduces the reference's epoch: its ``Dataset.__len__`` returns the CASE count
    (data_loader.py:112-113) although the arrays are flattened to ``count * min_time`` samples, so its shuffling
    ``DataLoader`` only ever draws sample indices below ``count`` -- the first ``count`` flattened rows, i.e. the first
    ``count / min_time`` cases -- and runs ``ceil(count / batch_size)`` iterations per epoch.  The default
    (``reference_len=False``) serves every sample; this is a deliberate deviation (the reference trains on ~1/min_time
    of its data), and the epoch length -- hence the per-epoch cosine LR schedule of train.py:65,151 -- differs
    accordingly."""

    def __init__(self, config, mode="train", device="cuda", seed=None, reference_len=False):
        cfg = config[mode]
        sub = {"train": "train", "valid": "val"}.get(mode, mode)
        self.states, self.targets, self.gso, self.count = load_cases(
            os.path.join(cfg["root_dir"], sub), cfg["min_time"], cfg["nb_agents"], cfg["max_time_dl"])
        self.batch_size = int(config["batch_size"])
        self.device = torch.device(device)
        self.rng = np.random.default_rng(seed)
        if self.device.type == "cuda":
            self.states, self.targets, self.gso = (t.pin_memory() for t in (self.states, self.targets, self.gso))
        self.train_loader = self
        self.reference_len = bool(reference_len)

    @property
    def num_samples(self):
        """Samples an epoch draws from: every flattened row, or the first `count` rows in reference_len mode."""
        return min(self.count, self.states.shape[0]) if self.reference_len else self.states.shape[0]

    def __len__(self):
        return (self.num_samples + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        order = torch.from_numpy(self.rng.permutation(self.num_samples))      # shuffle=True (data_loader.py:17)
        for lo in range(0, len(order), self.batch_size):
            idx = order[lo:lo + self.batch_size]
            yield tuple(t[idx].to(self.device, non_blocking=True) for t in (self.states, self.targets, self.gso))
