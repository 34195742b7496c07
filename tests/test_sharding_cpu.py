"""CPU: pair-parallel sharding logic of the N>1 path (one process per GPU, no data-path collective), exercised with
gloo at world size 2 — every pair is owned by exactly one rank, and the max-over-ranks timing reduction the bench uses."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gims_b200.engine import shard_pairs


def test_shard_pairs_partition():
    for n_pairs in (0, 1, 7, 8, 255, 256):
        for world in (1, 2, 3, 4, 8):
            owned = [list(shard_pairs(n_pairs, world, r)) for r in range(world)]
            flat = [i for o in owned for i in o]
            assert flat == list(range(n_pairs))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= (n_pairs + world - 1) // world


def _worker(rank, world, port, n_pairs, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = list(shard_pairs(n_pairs, world, rank))
    # stand-in for the per-rank result of the matcher: one int per owned pair
    local = torch.tensor([i * i for i in mine], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(mine)], dtype=torch.int64))
    t = torch.tensor([1.0 + rank])           # "elapsed ms" of this rank
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [torch.zeros(int(c), dtype=torch.int64) for c in counts]
    dist.all_gather(gathered, local) if len(set(int(c) for c in counts)) == 1 else None
    if rank == 0:
        out.put((sum(int(c) for c in counts), float(t), [g.tolist() for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    n_pairs = 8
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, tmax, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == n_pairs
    assert tmax == 2.0                         # max over ranks
    assert gathered[0] + gathered[1] == [i * i for i in range(n_pairs)]
