"""1-D block-row cyclic vs 2-D block-cyclic distributed Cholesky on the same Gram matrix (run under torchrun):
   torchrun --nproc-per-node P tools/dist_bench2d.py N nb pr:pc [pr:pc ...]
Prints the max-over-ranks factorisation time of `DistributedCholesky` (1-D) and of `BlockCyclic2DCholesky` for every
process grid given, and (when the replicated 1-D factor fits) the largest deviation of the 2-D factor's local blocks from it."""
import os
import sys

sys.path.insert(0, ".")
os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
import torch
import torch.distributed as dist

import bench
import linpde_gp_b200 as lg
from linpde_gp_b200 import backend, distributed
from linpde_gp_b200.linfuncops import diffops
from linpde_gp_b200.randprocs import _conditional, covfuncs

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, nb = int(sys.argv[1]), int(sys.argv[2])
grids = [tuple(int(v) for v in a.split(":")) for a in sys.argv[3:]]
nbc_edge = N // 128
prob = bench.make_problem(N - 4 * nbc_edge, nbc_edge, 16)
k = bench.SIGMA2 * covfuncs.TensorProduct(covfuncs.Matern((), nu=bench.NU, lengthscales=prob["ell"]),
                                          covfuncs.Matern((), nu=bench.NU, lengthscales=prob["ell"]))
prior = lg.GaussianProcess(lg.functions.Zero(input_shape=(2,)), k)
lap = -1.0 * diffops.Laplacian((2,))
CGP = lg.ConditionalGaussianProcess
blocks, off = [], 0
for Y, X, L in [(Yb, Xb, None) for Xb, Yb in zip(prob["edges"], prob["Y_bc"])] + [(prob["Y_pde"], prob["X_pde"], lap)]:
    atoms = CGP._preprocess_observations(prior=prior, Y=Y, X=X, L=L, b=None)[3]
    blk = _conditional._Block(None, None, 2, off, atoms=atoms)
    blocks.append(blk)
    off += blk.n_phys
noises = [None] * len(blocks)


def tmax(ms):
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return tmax(e0.elapsed_time(e1))


replicate = N * N * 8 * (1.0 + 1.0 / world) + (8 << 30) < 0.5 * torch.cuda.get_device_properties(local).total_memory
L_full = backend.alloc_matrix(N, N) if replicate else None
t1 = None
for rep in range(2):
    ch = distributed.DistributedCholesky(N, nb=nb)
    for i in ch.layout.local_blocks(ch.rank):
        g0, g1 = ch.layout.block_bounds(i)
        CGP._assemble_range(prior, blocks, noises, ch.local_block_rows(i), g0, g1)
    t1 = timed(lambda: ch.factor(L_full))
    del ch
if rank == 0:
    print(f"P={world} N={N} nb={nb}  1-D block-row cyclic:      factor {t1:8.1f} ms ({N**3 / 3 / t1 * 1e-9:7.2f} TFLOP/s aggregate)", flush=True)
strip = backend.alloc_matrix(nb, N)
for pr, pc in grids:
    t2 = None
    for rep in range(2):
        ch2 = distributed.BlockCyclic2DCholesky(N, nb, pr, pc)
        lay = ch2.layout
        for li in range(ch2.nbr):
            i = li * pr + ch2.r
            CGP._assemble_range(prior, blocks, noises, strip, i * nb, (i + 1) * nb)
            for lj in range(ch2.nbc):
                j = lj * pc + ch2.c
                if j <= i:
                    ch2.A_loc[li * nb : (li + 1) * nb, lj * nb : (lj + 1) * nb].copy_(strip[:, j * nb : (j + 1) * nb])
        t2 = timed(ch2.factor)
        dev = 0.0
        if rep == 1 and L_full is not None:
            for li in range(ch2.nbr):
                i = li * pr + ch2.r
                for lj in range(ch2.nbc):
                    j = lj * pc + ch2.c
                    if j <= i:
                        a = ch2.A_loc[li * nb : (li + 1) * nb, lj * nb : (lj + 1) * nb]
                        b = L_full[i * nb : (i + 1) * nb, j * nb : (j + 1) * nb]
                        dev = max(dev, float(((a - b) if j < i else torch.tril(a - b)).abs().max()))
            dev = tmax(dev)
        del ch2
    if rank == 0:
        print(f"P={world} N={N} nb={nb}  2-D block-cyclic {pr} x {pc}:     factor {t2:8.1f} ms ({N**3 / 3 / t2 * 1e-9:7.2f} TFLOP/s aggregate)"
              + (f"   max |L_2D - L_1D| = {dev:.2e}" if L_full is not None else ""), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
