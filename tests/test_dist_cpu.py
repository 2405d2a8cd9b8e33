"""world_size-2 `gloo` test of the N>1 bench path on CPU: clouds are sharded by rank with no
data-path collective, the only exchanges are the voxel-count all-reduce, the
max-over-ranks timing reduction and (in training) the gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bench import make_batch
    clouds = 2
    coords, feats, labels = make_batch(clouds, 500, seed0=rank * clouds)
    # shards are disjoint by construction: different seeds, own batch indices 0..clouds-1
    n = torch.tensor([float(coords.shape[0])])
    dist.all_reduce(n)
    ms = torch.tensor([10.0 + rank])          # pretend per-rank step time
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # gradient all-reduce as DDP would do it (mean over ranks)
    g = torch.full((4,), float(rank + 1))
    dist.all_reduce(g)
    g /= world
    digest = int(coords.long().sum().item())
    gathered = [None] * world
    dist.all_gather_object(gathered, digest)
    if rank == 0:
        out.put((float(n.item()), float(ms.item()), g.tolist(), gathered,
                 int(coords[:, 0].max().item())))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    total, ms, g, digests, max_batch = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == 2 * 2 * 500            # every rank contributes clouds x voxels
    assert ms == 11.0                      # max over ranks
    assert g == [1.5] * 4                  # mean of rank gradients
    assert digests[0] != digests[1]        # ranks really hold different clouds
    assert max_batch == 1                  # batch indices are local to the rank
