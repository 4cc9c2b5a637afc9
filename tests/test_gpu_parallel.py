"""2-rank NCCL parity of the multi-GPU training step (SURVEY 8e option i = DDP semantics of the reference): batch rows
sharded over the ranks, per-shard BatchNorm statistics, gradient = mean of the ranks' gradients, identical clip + Adam on
every rank.  Expected values: the torch-CPU oracle (`loss_and_grads` per shard, averaged, `clip_and_adam`).
Needs 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import numpy as np
import pytest
import torch

from helpers import oracle_params, rel_err
from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(name, B, N, seed):
    rng = np.random.default_rng(seed)
    states = torch.tensor(rng.integers(0, 4, size=(B, N, 2, 5, 5))).float()
    adj = torch.tensor((rng.random((B, N, N)) < 0.3).astype(np.float32))
    adj[:, torch.arange(N), torch.arange(N)] = 0
    gso = adj + torch.eye(N)
    targets = torch.tensor(rng.integers(0, 5, size=(B, N)))
    return states, gso, targets


def _worker(rank, world, port, name, B, N, steps, use_graph, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from helpers import load_network
    import mapf_gnn_b200.train as tr
    from mapf_gnn_b200.parallel import shard_range
    net = load_network(name, num_agents=N, device=f"cuda:{rank}").train()
    trainer = tr.Trainer(net, use_graph=use_graph)
    states, gso, targets = _data(name, B, N, 0)
    lo, hi = shard_range(B, rank, world)
    dev = torch.device("cuda", rank)
    losses = []
    for _ in range(steps):
        loss = trainer.step(states[lo:hi].to(dev), None if name == "baseline" else gso[lo:hi].to(dev),
                            targets[lo:hi].to(dev))
        losses.append(float(loss.item()))
    torch.cuda.synchronize()
    out[rank] = {"params": {k: v.detach().cpu().numpy() for k, v in net.named_parameters()},
                 "grad": trainer.flat_grad.cpu().numpy(), "loss": losses, "norm": float(trainer.grad_norm.item()),
                 "world": trainer.world}
    dist.destroy_process_group()


@pytest.mark.parametrize("name,use_graph", [("gnn_msg_k3", False), ("gnn_k3", True), ("baseline", True)])
def test_two_rank_trainer_matches_ddp_of_the_reference(name, use_graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world, B, N, steps = 2, 96, 5, 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), name, B, N, steps, use_graph, out), nprocs=world, join=True)
    # oracle: DDP of the reference = per-shard train-mode forward / backward, gradients averaged, one optimiser
    from mapf_gnn_b200.parallel import shard_range
    states, gso, targets = _data(name, B, N, 0)
    params = [oracle_params(name) for _ in range(world)]      # running stats diverge per rank, trainables stay equal
    keys = mo.trainable_keys(params[0])
    m = {k: torch.zeros_like(params[0][k]) for k in keys}
    v = {k: torch.zeros_like(params[0][k]) for k in keys}
    ref_losses = [[] for _ in range(world)]
    for step in range(1, steps + 1):
        grads = []
        for r in range(world):
            lo, hi = shard_range(B, r, world)
            loss, g, bn = mo.loss_and_grads(params[r], states[lo:hi], gso[lo:hi], targets[lo:hi])
            grads.append(g)
            ref_losses[r].append(float(loss))
            params[r] = dict(params[r], **{k: t.detach() for k, t in bn.items()})
        avg = {k: sum(g[k] for g in grads) / world for k in keys}
        newp, m, v, total = mo.clip_and_adam(params[0], avg, m, v, step=step)
        for r in range(world):
            params[r] = dict(params[r], **newp)
    for r in range(world):
        res = out[r]
        assert res["world"] == world
        for a, b in zip(res["loss"], ref_losses[r]):
            assert abs(a - b) < 1e-4 * abs(b), (r, a, b)
        assert abs(res["norm"] - float(total)) < 1e-3 * float(total)
        for k in keys:
            got, ref = res["params"][k], params[0][k].numpy()
            if k.startswith("convLayers") and k.endswith(".bias") and int(k.split(".")[1]) % 3 == 0:
                # conv biases in front of train-mode BatchNorm: true gradient 0, Adam amplifies rounding noise to +-lr
                assert np.abs(got - ref).max() <= 2 * steps * 3e-4 + 1e-7, k
            else:
                assert rel_err(got, ref) < 1e-4, (r, k)
    # every rank holds the same parameters bit for bit (same averaged gradient, same kernel)
    for k in keys:
        assert np.array_equal(out[0]["params"][k], out[1]["params"][k]), k
