"""World-size-2 gloo test (CPU) of the multi-GPU host logic: test-grid sharding and result gathering."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from linpde_gp_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_bounds(n, rank, world)
    full = torch.arange(n, dtype=torch.float64) ** 2
    got = parallel.gather_concat(full[lo:hi].clone(), n)
    mx = parallel.max_over_ranks(float(rank + 1))
    q.put((rank, bool(torch.equal(got, full)), mx))
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 262144, 262147):
        for world in (1, 2, 3, 8):
            b = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_gather_concat_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1001, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert all(mx == 2.0 for _, _, mx in res)


def test_single_process_is_identity():
    x = torch.arange(5, dtype=torch.float64)
    assert parallel.gather_concat(x, 5) is x
    assert parallel.max_over_ranks(3.5) == 3.5
