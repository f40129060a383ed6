"""Timing and per-stage timeline of the multi-GPU assembly + Cholesky (run under torchrun):
   torchrun --nproc-per-node P tools/dist_bench.py N nb [nb ...]
Prints, per nb, the max-over-ranks device time of block-row assembly and of the distributed factorisation
(aggregate TFLOP/s = N^3/3 / time), the CUDA-event time of every stage of the panel chain on rank 0 (what the critical
path is made of: leaf chain / NCCL / copies / GEMMs), and the single-GPU recursive potrf of the same matrix on rank 0."""
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
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
nbs = [int(a) for a in sys.argv[2:]] or [512]
nbc_edge = N // 128
prob = bench.make_problem(N - 4 * nbc_edge, nbc_edge, 16)
k = bench.SIGMA2 * covfuncs.TensorProduct(covfuncs.Matern((), nu=bench.NU, lengthscales=prob["ell"]),
                                          covfuncs.Matern((), nu=bench.NU, lengthscales=prob["ell"]))
prior = lg.GaussianProcess(lg.functions.Zero(input_shape=(2,)), k)
lap = -1.0 * diffops.Laplacian((2,))
batches = [(Yb, Xb, None) for Xb, Yb in zip(prob["edges"], prob["Y_bc"])] + [(prob["Y_pde"], prob["X_pde"], lap)]
CGP = lg.ConditionalGaussianProcess
blocks, off = [], 0
for Y, X, L in batches:
    atoms = CGP._preprocess_observations(prior=prior, Y=Y, X=X, L=L, b=None)[3]
    blk = _conditional._Block(None, None, 2, off, atoms=atoms)
    blocks.append(blk)
    off += blk.n_phys
assert off == N
noises = [None] * len(blocks)


def tmax(ms):
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


replicate = N * N * 8 * (1.0 + 1.0 / world) + (8 << 30) < 0.5 * torch.cuda.get_device_properties(local).total_memory
L_full = backend.alloc_matrix(N, N) if replicate else None
for nb in nbs:
    for rep in range(2):
        ch = distributed.DistributedCholesky(N, nb=nb)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev[0].record()
        for i in ch.layout.local_blocks(ch.rank):
            g0, g1 = ch.layout.block_bounds(i)
            CGP._assemble_range(prior, blocks, noises, ch.local_block_rows(i), g0, g1)
        ev[1].record()
        prof = {} if rep == 1 else None
        ch.factor(L_full, profile=prof)
        ev[2].record()
        torch.cuda.synchronize()
        t_asm, t_fac = tmax(ev[0].elapsed_time(ev[1])), tmax(ev[1].elapsed_time(ev[2]))
        if rank == 0 and rep == 1:
            print(f"P={world} N={N} nb={nb} replicate={replicate}: assemble {t_asm:.2f} ms ({N * (N + 1) / 2 / t_asm * 1e-6:.1f} "
                  f"Gentries/s aggregate), factor {t_fac:.1f} ms ({N**3 / 3 / t_fac * 1e-9:.2f} TFLOP/s aggregate)", flush=True)
            stages = ("potrf", "bcast", "trsm", "gather", "rotate", "update_next", "update_rest")
            panel = sum(prof.get(s, 0.0) for s in stages[:-1])
            print("   rank-0 timeline [ms over %d panels]: " % prof["panels"]
                  + ", ".join(f"{s} {prof.get(s, 0.0):.1f}" for s in stages)
                  + f" | panel stream busy {panel:.1f}, update stream busy {prof.get('update_rest', 0.0):.1f}, total {prof['total']:.1f}",
                  flush=True)
        del ch
if rank == 0 and replicate:
    # single-GPU recursive factorisation of the same matrix, and agreement of the two factors
    f = backend.DeviceFactor([N])
    CGP._assemble_range(prior, blocks, noises, f.L, 0, N)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    f.potrf()
    e1.record()
    torch.cuda.synchronize()
    t1 = e0.elapsed_time(e1)
    diff = 0.0
    for r in range(0, N, 4096):
        diff = max(diff, float((torch.tril(f.L[r : r + 4096], diagonal=r) - torch.tril(L_full[r : r + 4096], diagonal=r)).abs().max()))
    print(f"single-GPU potrf N={N}: {t1:.1f} ms ({N**3 / 3 / t1 * 1e-9:.2f} TFLOP/s); max |L_dist - L_single| = {diff:.2e}", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
