"""Distributed (block-row cyclic) Cholesky: orchestration logic under gloo, world_size 1 and 2, with the host test
double for the local kernels (the CUDA kernels themselves are covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from linpde_gp_b200.distributed import BlockRowLayout, DistributedCholesky
from tests.dist_double import HostOps


def _spd(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(n, n + 8, dtype=torch.float64, generator=g)
    return X @ X.T / n + 0.5 * torch.eye(n, dtype=torch.float64)


def _run(rank, world, port, n, nb, q):
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    G = _spd(n, 3)
    ch = DistributedCholesky(n, nb=nb, ops=HostOps())
    for i in ch.layout.local_blocks(ch.rank):
        lo, hi = ch.layout.block_bounds(i)
        ch.local_block_rows(i)[:, :hi].copy_(G[lo:hi, :hi])  # lower part only
    L_full = torch.zeros((n, (n + 15) // 16 * 16), dtype=torch.float64)[:, :n]
    ch.factor(L_full)                      # the gathered panels fill the replicated factor on the fly
    L_rep = torch.zeros_like(L_full)
    ch.replicate_into(L_rep)               # explicit replication of the distributed block rows: same result
    L_ref = torch.linalg.cholesky(G)
    err = max(float((torch.tril(L_full) - L_ref).abs().max()), float((torch.tril(L_rep) - L_ref).abs().max()))
    # inverted leaf blocks are replicated too
    l0 = torch.linalg.inv(L_ref[:128, :128])
    werr = float((ch.dinv[: 128 * 128].view(128, 128) - l0).abs().max())
    # distributed solve and streamed posterior variance (nothing replicated) against dense references
    g = torch.Generator().manual_seed(11)
    b = torch.randn(n, dtype=torch.float64, generator=g)
    x = ch.solve(b)
    err = max(err, float((x - torch.cholesky_solve(b[:, None], L_ref)[:, 0]).abs().max()))
    m = 37 + 5 * rank  # ragged, different per rank
    K = torch.randn(m, (n + 15) // 16 * 16, dtype=torch.float64, generator=g)[:, :n]
    V_ref = torch.linalg.solve_triangular(L_ref, K.T.contiguous(), upper=False).T
    var = ch.post_var(K, 3.0)
    err = max(err, float((var - (3.0 - (V_ref * V_ref).sum(1))).abs().max()) * 1e-2)
    q.put((rank, err, werr))
    if world > 1:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_layout_bookkeeping():
    lay = BlockRowLayout(1280 + 256, 512, 2)
    assert lay.nblk == 3 and lay.local_blocks(0) == [0, 2] and lay.local_blocks(1) == [1]
    assert lay.n_local(0) == 512 + 512 and lay.n_local(1) == 512
    assert lay.block_bounds(2) == (1024, 1536)
    assert lay.first_local_block_after(0, 0) == 1 and lay.first_local_block_after(1, 0) == 0
    assert lay.rows_after(0, 0) == 512 and lay.rows_after(1, 0) == 512 and lay.rows_after(1, 1) == 0
    with pytest.raises(ValueError):
        BlockRowLayout(1000, 500, 2)


@pytest.mark.parametrize("n,nb", [(512, 128), (1280, 256), (1100, 256)])
def test_single_process(n, nb):
    import queue

    q = queue.Queue()
    _run(0, 1, 0, n, nb, q)
    _, err, werr = q.get()
    assert err < 1e-11 and werr < 1e-10


def test_not_positive_definite_raises():
    n, nb = 640, 128
    G = _spd(n, 5)
    G[300, 300] = -1.0
    ch = DistributedCholesky(n, nb=nb, ops=HostOps())
    for i in ch.layout.local_blocks(0):
        lo, hi = ch.layout.block_bounds(i)
        ch.local_block_rows(i)[:, :hi].copy_(G[lo:hi, :hi])
    with pytest.raises(np.linalg.LinAlgError):
        ch.factor()


@pytest.mark.parametrize("n,nb", [(1280, 256), (1100, 128), (1666, 128)])
def test_world2_gloo(n, nb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, 2, port, n, nb, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(err < 1e-11 and werr < 1e-10 for _, err, werr in res), res


# ---- 2-D block-cyclic variant ------------------------------------------------------------------------------------------
def _run2d(rank, world, port, n, nb, pr, pc, q):
    from linpde_gp_b200.distributed import BlockCyclic2DCholesky

    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    G = _spd(n, 5)
    ch = BlockCyclic2DCholesky(n, nb, pr, pc, ops=HostOps())
    lay = ch.layout
    for i in range(lay.nblk):
        for j in range(i + 1):
            if ch.owns(i, j):
                ch.local_block(i, j).copy_(G[i * nb : (i + 1) * nb, j * nb : (j + 1) * nb])
    ch.factor()
    L = torch.zeros((n, n), dtype=torch.float64)
    ch.gather_full(L)
    err = float((torch.tril(L) - torch.linalg.cholesky(G)).abs().max())
    # a matrix that is not positive definite: every rank raises, with LAPACK's info
    bad = BlockCyclic2DCholesky(n, nb, pr, pc, ops=HostOps())
    Gb = G.clone()
    Gb[nb + 3, nb + 3] = -1.0
    for i in range(lay.nblk):
        for j in range(i + 1):
            if bad.owns(i, j):
                bad.local_block(i, j).copy_(Gb[i * nb : (i + 1) * nb, j * nb : (j + 1) * nb])
    raised = False
    try:
        bad.factor()
    except np.linalg.LinAlgError:
        raised = True
    q.put((rank, err, raised))
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,pr,pc,n,nb", [(1, 1, 1, 512, 128), (2, 2, 1, 768, 128), (2, 1, 2, 768, 128), (4, 2, 2, 1152, 128),
                                              (6, 2, 3, 1024, 128), (6, 3, 2, 896, 128)])
def test_2d_block_cyclic_cholesky(world, pr, pc, n, nb):
    """``BlockCyclic2DCholesky`` on pr x pc process grids (row broadcast + column all-gather of the panel, staircase row
    limits, lookahead bookkeeping) against torch's dense Cholesky; non-positive-definite input raises on every rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run2d, args=(rk, world, port, n, nb, pr, pc, q)) for rk in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for _, err, raised in res:
        assert err < 1e-11 and raised
