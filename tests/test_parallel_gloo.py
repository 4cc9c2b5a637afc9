"""world_size-2 gloo tests (CPU) of the host-side multi-GPU plumbing: sharding, throughput aggregation
(max time over ranks, units summed), gradient averaging, metric gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mapf_gnn_b200 import parallel as par
    assert par.env_world() == (rank, rank, world)
    lo, hi = par.shard_range(4097, rank, world)
    value, tmax, units = par.aggregate_throughput(units=(hi - lo) * 5, seconds=1.0 + rank)
    g = torch.full((8,), float(rank + 1))
    par.average_gradients(g)
    s, f = par.gather_metrics(torch.full((3,), float(rank)), torch.full((3,), rank, dtype=torch.int32))
    out[rank] = dict(lo=lo, hi=hi, value=value, tmax=tmax, units=units, g=g.tolist(), s=s.tolist(), f=f.tolist())
    dist.destroy_process_group()


def test_two_rank_plumbing():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert (a["lo"], a["hi"], b["lo"], b["hi"]) == (0, 2049, 2049, 4097)
    for r in (a, b):
        assert r["units"] == 4097 * 5 and r["tmax"] == 2.0 and abs(r["value"] - 4097 * 5 / 2.0) < 1e-9
        assert r["g"] == [1.5] * 8
        assert r["s"] == [0.0] * 3 + [1.0] * 3 and r["f"] == [0] * 3 + [1] * 3


def test_shard_range_covers_everything():
    from mapf_gnn_b200.parallel import shard_range
    for total in (1, 7, 8, 65536, 65537):
        for world in (1, 2, 4, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
