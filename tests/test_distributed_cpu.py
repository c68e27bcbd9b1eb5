"""CPU, world_size 2, gloo: the batch-sharding host logic used by the multi-GPU runs
(no collective on the data path; gather + MAX-reduce for reporting only)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sella_b200.sharding import shard_range, shard_sizes, gather_rows, max_over_ranks


def test_shard_range_partitions_exactly():
    for total in (1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(total, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(total))
            sizes = shard_sizes(total, world)
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    # each rank "optimises" its own systems: result row i = global index i
    local = torch.arange(lo, hi, dtype=torch.float64).unsqueeze(1).repeat(1, 3)
    full = gather_rows(local, total, dst=0)
    tmax = max_over_ranks(1.0 + rank)
    if rank == 0:
        q.put((full.tolist(), tmax))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_and_max_reduce_world2():
    world, total = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, tmax = q.get(timeout=100)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert full == [[float(i)] * 3 for i in range(total)]
    assert tmax == 2.0
