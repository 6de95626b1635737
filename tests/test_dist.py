"""CPU tier: the N>1 host logic (channel sharding, stats all-reduce, record gather) with gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from m17_sdr_b200 import dist as md


def test_shard_range_partitions():
    for total in (0, 1, 7, 1024, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            r = [md.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert md.owner_of(0, 10, 4) == 0 and md.owner_of(9, 10, 4) == 3
    with pytest.raises(ValueError):
        md.shard_range(10, 4, 4)


def _worker(rank, world, port, total, cap, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c0, c1 = md.shard_range(total, rank, world)
    g = torch.Generator().manual_seed(7)
    allf = torch.randint(0, 256, (total, cap, 64), generator=g, dtype=torch.uint8)
    alln = torch.randint(0, cap + 1, (total,), generator=g, dtype=torch.int32)
    alls = torch.randint(0, 1000, (total, 8), generator=g, dtype=torch.int64)
    f, n = md.gather_records(allf[c0:c1].clone(), alln[c0:c1].clone(), total, dst=0)
    tot = md.reduce_stats(alls[c0:c1].clone())
    ok = bool(torch.equal(tot, alls.sum(0)))
    if rank == 0:
        ok = ok and torch.equal(f, allf) and torch.equal(n, alln)
    else:
        ok = ok and f is None
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gather_and_reduce_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 11, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=60) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
