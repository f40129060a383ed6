#!/usr/bin/env python
"""bench.py -- seconds per GP-PDE solve (2-D Poisson, N = 65,536) on B200, with roofline and CPU baseline.

One *step* = one full pass of the hot path over one synthetic problem (BASELINE.json configs[3]):
  2-D Poisson-Dirichlet on [0,1]^2, prior 4 * TensorProduct(Matern-5/2(l), Matern-5/2(l)), l = 4/sqrt(N_pde);
  4 boundary batches (N_bc = 2,048) + 1 PDE-collocation batch (N_pde = 63,488, seed 2) => N = 65,536;
  assemble all Gram blocks -> bordered FP64 Cholesky -> representer weights -> posterior mean AND pointwise
  variance on the 512 x 512 test grid.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--npde ..] [--nbc-edge ..] [--grid ..]

`value`      device-timed seconds per solve with all inputs already resident in HBM (CUDA events, max over ranks);
`e2e`        the same solve through the public Python API with HOST numpy buffers (pinned H2D of the points /
             right-hand sides and D2H of mean + variance inside the timed region);
`roofline`   the DMMA GEMM kernel (Cholesky trailing updates + posterior-variance triangular solve) against the
             FP64 tensor-pipe issue rate measured live on the box (lpgp_dmma_peak_probe);
`cpu_baseline` the oracle (numpy/scipy restatement of the reference) timed on the box's host cores on a bounded
             sample and extrapolated to the full solve.
With --gpus N > 1 (torchrun) the Gram matrix is assembled and factorised over all ranks (block-row cyclic layout,
NCCL panel exchange, linpde_gp_b200/distributed.py), the factor is then replicated and the test grid is sharded N
ways (strong scaling of the fixed N = 65,536 problem); result rows are gathered with NCCL.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIGMA2 = 4.0
NU = 2.5


# ----------------------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------------------
def make_problem(n_pde: int, n_bc_edge: int, grid: int, seed: int = 2):
    rng = np.random.default_rng(seed)
    ell = 4.0 / np.sqrt(n_pde)
    s = np.linspace(0.0, 1.0, n_bc_edge, endpoint=False)
    h = 1.0 / n_bc_edge
    edges = [
        np.stack([s, np.zeros_like(s)], -1),
        np.stack([np.ones_like(s), s], -1),
        np.stack([s + h, np.ones_like(s)], -1),
        np.stack([np.zeros_like(s), s + h], -1),
    ]
    Xp = rng.uniform(0.0, 1.0, size=(n_pde, 2))
    g = np.linspace(0.0, 1.0, grid)
    Xt = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    return {
        "ell": ell,
        "edges": [np.ascontiguousarray(e) for e in edges],
        "Y_bc": [np.zeros(len(e)) for e in edges],
        "X_pde": Xp,
        "Y_pde": np.full(n_pde, 2.0),
        "Xt": Xt,
        "N": n_pde + 4 * n_bc_edge,
        "M": Xt.shape[0],
    }


def oracle_kernel(ell):
    f = {"kind": "matern", "input_shape": [], "nu": NU, "lengthscales": float(ell)}
    return {"scale": SIGMA2, "base": {"kind": "tensor_product", "factors": [f, dict(f)]}}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores, bounded sample, extrapolated to one full solve
# ----------------------------------------------------------------------------------------------------------------
def cpu_sample(prob, budget_s: float = 20.0):
    """Bounded sample of the reference's CPU path (the oracle port) on ALL host cores, extrapolated to one solve:
    Gram assembly on independent row tiles spread over a thread pool (numpy releases the GIL inside its ufuncs; the
    reference itself is single-threaded here), LAPACK dpotrf / dtrsm through scipy with every BLAS thread."""
    import concurrent.futures as cf

    import scipy.linalg
    from threadpoolctl import threadpool_limits

    from oracle import covfuncs as ocf

    cores = len(os.sched_getaffinity(0))
    N, M = prob["N"], prob["M"]
    kern = oracle_kernel(prob["ell"])
    lap = [(-1.0, ("wl", np.ones(2)))]
    X = prob["X_pde"]
    t_budget = budget_s / 3.0
    # (a) Gram assembly L k L^* on row tiles (the un-tiled reference needs ~10 N x N temporaries)
    rows = 64
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        ocf.matrix(kern, lap, lap, X[:rows], X)
        t_one = max(time.perf_counter() - t0, 1e-3)
        per_worker = max(1, int(t_budget / t_one))
        workers = max(1, min(cores, 32))

        def work(wid):
            for i in range(per_worker):
                r0 = ((wid * per_worker + i) * rows) % (len(X) - rows)
                ocf.matrix(kern, lap, lap, X[r0 : r0 + rows], X)
            return per_worker * rows

        t0 = time.perf_counter()
        with cf.ThreadPoolExecutor(workers) as ex:
            done = sum(ex.map(work, range(workers)))
        t_asm = time.perf_counter() - t0
    eps = done * len(X) / t_asm
    with threadpool_limits(limits=cores):
        # (b) dpotrf via scipy at a host-sized N
        nc = 8192
        rng = np.random.default_rng(0)
        A = rng.standard_normal((nc, nc))
        G = A @ A.T / nc + 2.0 * np.eye(nc)
        t0 = time.perf_counter()
        Lc = scipy.linalg.cholesky(G, lower=True)
        t_chol = time.perf_counter() - t0
        chol_flops = nc**3 / 3.0 / t_chol
        # (c) dtrsm (posterior variance in the N x M_c form, _conditional.py:245-251)
        mc = 4096
        B = rng.standard_normal((nc, mc))
        t0 = time.perf_counter()
        scipy.linalg.solve_triangular(Lc, B, lower=True, check_finite=False)
        t_trsm = time.perf_counter() - t0
        trsm_flops = nc * nc * mc / t_trsm
    est = N * N / eps + M * N / eps + (N**3 / 3.0) / chol_flops + (float(M) * N * N) / trsm_flops
    detail = {
        "gram_entries_per_s": eps,
        "cholesky_gflops": chol_flops * 1e-9,
        "trsm_gflops": trsm_flops * 1e-9,
        "sample": f"LkL assembly of {done}x{len(X)} entries in {rows}-row tiles on {workers} threads ({t_asm:.1f} s), "
                  f"scipy dpotrf n={nc} ({t_chol:.1f} s), dtrsm {nc}x{mc} ({t_trsm:.1f} s), BLAS threads {cores}; "
                  f"extrapolated to N={N}, M={M}",
    }
    return est, detail


def host_threads():
    """(threads used by the sample, host cores available): cpu_sample drives every core explicitly, whatever
    OMP_NUM_THREADS torchrun may have exported."""
    cores = len(os.sched_getaffinity(0))
    return cores, cores


# ----------------------------------------------------------------------------------------------------------------
# GPU clocks during the timed region
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.path = tempfile.mktemp(prefix="lpgp_clocks_", suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
                power.append(float(p[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        load = [c for c, w in zip(sm, power) if w > 0.5 * max(power)] if power else sm
        return {
            "sm_mhz": float(np.median(load)) if load else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "power_w_max": float(max(power)) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ----------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------
class DeviceSolve:
    """The hot path driven at the device-buffer level (inputs resident in HBM): this is what `value` times."""

    def __init__(self, prob, rank: int, world: int):
        import torch

        import linpde_gp_b200 as lg
        from linpde_gp_b200 import backend
        from linpde_gp_b200._lowering import Factor1D, lower

        self.torch, self.be = torch, backend
        fac = [Factor1D("matern", prob["ell"], nu=NU), Factor1D("matern", prob["ell"], nu=NU)]
        lap = {(2, 0): -1.0, (0, 2): -1.0}
        self.d_k = lower(fac, None, None, SIGMA2)
        self.d_kL = lower(fac, None, lap, SIGMA2)   # test/boundary side x PDE side
        self.d_Lk = lower(fac, lap, None, SIGMA2)   # PDE rows x boundary columns
        self.d_LkL = lower(fac, lap, lap, SIGMA2)
        self.edges = [backend.to_device(e) for e in prob["edges"]]
        self.Xp = backend.to_device(prob["X_pde"])
        self.y = torch.cat([backend.to_device(y) for y in prob["Y_bc"]] + [backend.to_device(prob["Y_pde"])])
        Xt = prob["Xt"]
        from linpde_gp_b200 import parallel

        lo, hi = parallel.shard_bounds(len(Xt), rank, world)
        self.Xt = backend.to_device(Xt[lo:hi])
        self.N, self.M = prob["N"], prob["M"]
        self.t = {}
        self.distributed = world > 1

    def _timed(self, name, fn):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self.t.setdefault(name, []).append((e0, e1))
        return out

    def step(self):
        be = self.be
        sizes = [e.shape[0] for e in self.edges] + [self.Xp.shape[0]]
        blocksX = self.edges + [self.Xp]
        factor = None
        off = 0
        for i, (X, n) in enumerate(zip(blocksX, sizes)):
            is_pde = i == len(sizes) - 1
            factor = be.DeviceFactor([n]) if factor is None else factor.extended(n)

            def assemble():
                rows = factor.L[off : off + n]
                c = 0
                for Xj, nj in zip(blocksX[:i], sizes[:i]):
                    be.gram(self.d_Lk if is_pde else self.d_k, X, Xj, out=rows[:, c : c + nj])
                    c += nj
                be.gram(self.d_LkL if is_pde else self.d_k, X, None, out=rows[:, off : off + n], lower=True)

            self._timed("assemble", assemble)
            self._timed("factor", factor.potrf if i == 0 else factor.append_last)
            off += n
        w = self._timed("solve", lambda: factor.potrs(self.y.clone().reshape(1, -1)).reshape(-1))
        descs = [self.d_k] * len(self.edges) + [self.d_kL]
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]])
        blocks = be.ObsBlocks(descs, blocksX, offs)
        mean = self._timed("mean", lambda: be.post_mean(blocks, w, self.Xt))
        chunk = be.var_chunk_rows(factor.n, self.Xt.shape[0])
        var = self._timed("var", lambda: be.post_var(blocks, factor, self.Xt, self.d_k.diag_value, chunk=chunk))
        return mean, var

    # -- multi-GPU step: block-row cyclic assembly + distributed Cholesky, replicated factor, sharded test grid --
    def _desc_for(self, bi: int, bj: int):
        last = len(self.edges)
        if bi == last:
            return self.d_LkL if bj == last else self.d_Lk
        return self.d_k

    def _assemble_block_rows(self, out, g0: int, g1: int, blocksX, offs, sizes):
        """rows g0..g1 of the lower triangle of the block-structured Gram matrix -> out[:, :g1]"""
        be = self.be
        for bi, (Xi, oi, ni) in enumerate(zip(blocksX, offs, sizes)):
            r_lo, r_hi = max(g0, oi), min(g1, oi + ni)
            if r_lo >= r_hi:
                continue
            rows = out[r_lo - g0 : r_hi - g0]
            Xr = Xi[r_lo - oi : r_hi - oi]
            for bj in range(bi + 1):
                oj, nj = offs[bj], sizes[bj]
                c_hi = nj if bj < bi else r_hi - oi
                be.gram(self._desc_for(bi, bj), Xr, blocksX[bj][:c_hi], out=rows[:, oj : oj + c_hi])

    def step_distributed(self, nb: int, replicate: bool = True):
        from linpde_gp_b200 import distributed

        be, torch = self.be, self.torch
        sizes = [e.shape[0] for e in self.edges] + [self.Xp.shape[0]]
        blocksX = self.edges + [self.Xp]
        offs = [int(o) for o in np.concatenate([[0], np.cumsum(sizes)[:-1]])]
        n = int(sum(sizes))
        ch = distributed.DistributedCholesky(n, nb=nb)

        def assemble():
            for i in ch.layout.local_blocks(ch.rank):
                g0, g1 = ch.layout.block_bounds(i)
                self._assemble_block_rows(ch.local_block_rows(i), g0, g1, blocksX, offs, sizes)

        self._timed("assemble", assemble)
        descs = [self.d_k] * len(self.edges) + [self.d_kL]
        blocks = be.ObsBlocks(descs, blocksX, offs)
        if replicate:
            factor = be.DeviceFactor([n])
            self._timed("factor", lambda: ch.factor(factor.L))  # every rank ends up with the whole factor
            factor.dinv[: ch.dinv.numel() - 8].copy_(ch.dinv[:-8])
            del ch
            w = self._timed("solve", lambda: factor.potrs(self.y.clone().reshape(1, -1)).reshape(-1))
            mean = self._timed("mean", lambda: be.post_mean(blocks, w, self.Xt))
            chunk = be.var_chunk_rows(factor.n, self.Xt.shape[0])
            var = self._timed("var", lambda: be.post_var(blocks, factor, self.Xt, self.d_k.diag_value, chunk=chunk))
        else:  # the factor stays distributed: owner-computes solve, block rows of L streamed for the variance
            self._timed("factor", ch.factor)
            fac = distributed.DistributedFactor(ch)
            w = self._timed("solve", lambda: ch.solve(self.y))
            mean = self._timed("mean", lambda: be.post_mean(blocks, w, self.Xt))
            var = self._timed("var", lambda: fac.post_var(blocks, self.Xt, self.d_k.diag_value))
        return mean, var

    def phase_ms(self, last_k: int):
        self.torch.cuda.synchronize()
        per = lambda k: 5 if (k in ("assemble", "factor") and not self.distributed) else 1
        return {k: sum(e0.elapsed_time(e1) for e0, e1 in v[-last_k * per(k):]) / last_k for k, v in self.t.items()}


def api_solve(prob, rank: int, world: int, nb: int = 1024, replicate: bool = True):
    """The same solve through the public reference-style API with host (numpy) buffers -> `e2e`."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.linfuncops import diffops
    from linpde_gp_b200.randprocs import covfuncs

    k = SIGMA2 * covfuncs.TensorProduct(covfuncs.Matern((), nu=NU, lengthscales=prob["ell"]),
                                        covfuncs.Matern((), nu=NU, lengthscales=prob["ell"]))
    post = lg.GaussianProcess(lg.functions.Zero(input_shape=(2,)), k)
    lap = -1.0 * diffops.Laplacian((2,))
    if world > 1:  # one-shot conditioning on all batches, Gram assembly + Cholesky distributed over the ranks
        batches = [(Yb, Xb) for Xb, Yb in zip(prob["edges"], prob["Y_bc"])] + [(prob["Y_pde"], prob["X_pde"], lap)]
        post = lg.ConditionalGaussianProcess.from_observation_batches(post, batches, nb=nb, replicate=replicate)
    else:
        for Xb, Yb in zip(prob["edges"], prob["Y_bc"]):
            post = post.condition_on_observations(Yb, X=Xb)
        post = post.condition_on_observations(prob["Y_pde"], X=prob["X_pde"], L=lap)
    from linpde_gp_b200 import parallel

    lo, hi = parallel.shard_bounds(prob["M"], rank, world)
    Xt = prob["Xt"][lo:hi]
    return post.mean(Xt), post.var(Xt)


def dmma_peak_tflops(torch, backend):
    import ctypes

    from linpde_gp_b200._lib import lib

    sms = torch.cuda.get_device_properties(0).multi_processor_count
    scratch = torch.empty(sms * 2 * 256, dtype=torch.float64, device="cuda")
    flops = ctypes.c_double(0.0)
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.lpgp_dmma_peak_probe(ctypes.c_void_p(scratch.data_ptr()), sms * 2, 20000, ctypes.byref(flops),
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0
        best = max(best, flops.value / e0.elapsed_time(e1) * 1e-9)
    return best


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")  # panel broadcasts / all-gathers are the critical path
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from linpde_gp_b200 import backend
    from linpde_gp_b200._lib import lib

    prob = make_problem(args.npde, args.nbc_edge, args.grid)
    N, M = prob["N"], prob["M"]
    if args.nb <= 0:
        args.nb = 1024 if world <= 2 else 512

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from linpde_gp_b200 import parallel

    def gather(mean, var):
        """the path's only exchange step: all-gather the sharded result rows (NCCL)"""
        if world == 1:
            return mean, var
        if not hasattr(mean, "cpu"):
            mean, var = backend.to_device(mean), backend.to_device(var)
        return parallel.gather_concat(mean, M), parallel.gather_concat(var, M)

    peak = dmma_peak_tflops(torch, backend) if rank == 0 else None
    ds = DeviceSolve(prob, rank, world)
    # the replicated n x n factor, this rank's block rows and the variance workspace must fit next to each other
    if args.replicate == "auto":
        replicate = (N * N * 8) * (1.0 + 1.0 / world) + (8 << 30) < 0.5 * torch.cuda.get_device_properties(local).total_memory
    else:
        replicate = args.replicate == "yes"
    step = ds.step if world == 1 else (lambda: ds.step_distributed(args.nb, replicate))
    for _ in range(args.warmup):
        m, v = step()
        gather(m, v)
    barrier()
    lib.lpgp_launch_count(1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        m, v = step()
        gm, gv = gather(m, v)
    e1.record()
    barrier()
    launches = lib.lpgp_launch_count(0)
    ms_dev = e0.elapsed_time(e1) / args.steps
    phases = ds.phase_ms(args.steps)
    # the assembly kernel on its own (one L k L* block of 16,384 x N_pde entries, written to HBM), best of 3
    gram_kernel = None
    if rank == 0:
        nrow = min(16384, ds.Xp.shape[0])
        blk = backend.alloc_matrix(nrow, ds.Xp.shape[0])
        best = 1e30
        for _ in range(4):
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            backend.gram(ds.d_LkL, ds.Xp[:nrow], ds.Xp, out=blk)
            g1.record()
            torch.cuda.synchronize()
            best = min(best, g0.elapsed_time(g1))
        ent = nrow * ds.Xp.shape[0] / (best * 1e-3)
        gram_kernel = {"entries_per_s": ent, "hbm_write_gbps": ent * 8e-9, "block": [int(nrow), int(ds.Xp.shape[0])],
                       "kernel": "gram_sep_kernel<2,3,false> (L k L*, product Matern-5/2)"}
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm = json.load(open(peaks_path)).get("hbm_gbs") if os.path.exists(peaks_path) else None
        gram_kernel["hbm_peak_gbps"] = hbm if hbm else 6650.0
        gram_kernel["hbm_peak_source"] = "MEASURED_PEAKS.json" if hbm else "fallback (B200_PROFILING.md: 6.65 TB/s)"
        gram_kernel["frac_of_hbm_peak"] = gram_kernel["hbm_write_gbps"] / gram_kernel["hbm_peak_gbps"]
        del blk
    del ds
    torch.cuda.empty_cache()

    # ---- e2e through the public API, host buffers (pinned H2D + D2H inside the timed region) ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        em, ev = api_solve(prob, rank, world, args.nb, replicate)
        gem, gev = gather(em, ev)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop() if rank == 0 else None
    times = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = (float(x) for x in times.cpu())

    if rank == 0:
        tonp = lambda a: a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)
        gm, gv, gem, gev = tonp(gm), tonp(gv), tonp(gem), tonp(gev)
        agree = float(max(np.max(np.abs(gm - gem)), np.max(np.abs(gv - gev))))
        m_shard = parallel.shard_bounds(M, 0, world)[1]
        flops_tensor = N**3 / 3.0 / world + float(m_shard) * N * N  # per rank: Cholesky share + variance TRSM
        t_tensor = (phases["factor"] + phases["var"]) * 1e-3
        achieved = flops_tensor / t_tensor * 1e-12
        h2d = sum(e.nbytes for e in prob["edges"]) + sum(y.nbytes for y in prob["Y_bc"]) + prob["X_pde"].nbytes \
            + prob["Y_pde"].nbytes + prob["Xt"].nbytes * 2 // world
        # the CPU baseline is reported by the single-GPU run only (rank 0 at N = 1)
        cpu_est, cpu_detail = cpu_sample(prob, budget_s=20.0) if world == 1 else (None, {})
        threads, cores = host_threads()
        out = {
            "metric": "s per GP-PDE solve (2D Poisson N=64k)",
            "value": ms_dev * 1e-3,
            "unit": "s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_dev,
            "higher_is_better": False,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": f"2D Poisson Dirichlet synthetic N={N} (N_pde={args.npde} seed 2, N_bc={4 * args.nbc_edge}), "
                            f"product Matern-5/2 prior, 5 conditioning batches, mean+variance on {args.grid}x{args.grid} grid "
                            f"(BASELINE.json configs[{4 if N >= 131072 else 3 if N >= 65536 else 1}])",
                "parallelism": (f"block-row cyclic Gram assembly + Cholesky over {world} ranks (nb={args.nb}, NCCL panel "
                                f"exchange), factor {'replicated' if replicate else 'left distributed (block rows streamed for the variance)'}, "
                                "test grid sharded") if world > 1 else "single GPU",
                "l2": f"working set {N * N * 8 / 1e9:.1f} GB Gram >> 126 MB L2 (no flush needed)",
            },
            "phases_ms": phases,
            "gram_entries_per_s": (N * (N + 1) / 2 + 0.0) / (phases["assemble"] * 1e-3),  # whole assembly phase
            "gram_kernel": gram_kernel,
            "cholesky_tflops": N**3 / 3.0 / (phases["factor"] * 1e-3) * 1e-12,  # aggregate over all ranks
            "variance_trsm_tflops": float(m_shard) * N * N / (phases["var"] * 1e-3) * 1e-12,
            "e2e": {"value": ms_e2e * 1e-3, "unit": "s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(2 * 8 * M // world), "api_vs_device_max_abs_diff": agree},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                # dram__bytes_read+write of ONE representative launch (8192 x 8192 x 2048, beta = 1) from the committed
                # ncu --set full capture (profiles/ncu_kernels_r01d.md, banded tile rasterisation; 5.53e9 before it);
                # its algorithmic bytes are 1.34e9 (A + B once, C read + written) -- the excess is B re-streamed
                # once per band of tile rows, at 0.35 TB/s far from the HBM bound of this tensor-bound kernel
                "traffic": 2.786e9,
                "traffic_launch": "gemm_nt_kernel 8192x8192x2048 beta=1: 2.26 GB read + 0.53 GB written (algorithmic 1.34 GB)",
                "kernel": "gemm_nt_kernel (DMMA m8n8k4.f64) inside lpgp_potrf/lpgp_chol_append + lpgp_post_var",
                "peak_source": "FP64 tensor-pipe issue rate measured live (lpgp_dmma_peak_probe); MEASURED_PEAKS.json "
                               "holds no FP64 figure",
                "flops": flops_tensor,
            },
            "cpu_baseline": ({"value": cpu_est, "unit": "s", "cores": cores, "threads": threads, "kind": "port",
                              **cpu_detail} if cpu_est is not None else None),
        }
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob = make_problem(args.npde, args.nbc_edge, args.grid)
    for _ in range(min(args.warmup, 1)):
        cpu_sample(prob, budget_s=6.0)
    vals, detail = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, detail = cpu_sample(prob, budget_s=20.0)
        vals.append(v)
    wall = time.perf_counter() - t0
    threads, cores = host_threads()
    val = float(np.mean(vals))
    out = {
        "impl": "reference",
        "metric": "s per GP-PDE solve (2D Poisson N=64k)",
        "value": val, "unit": "s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall * 1e3 / max(args.steps, 1), "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"2D Poisson Dirichlet synthetic N={prob['N']}, mean+variance on {args.grid}x{args.grid} grid "
                               "(BASELINE.json configs[3]); bounded sample extrapolated to the full solve"},
        "cpu_baseline": {"value": val, "unit": "s", "cores": cores, "threads": threads, "kind": "port", **detail},
        "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(out))


_STDOUT_FD = None


def emit(line: str) -> None:
    """The ONE JSON line goes to the real stdout; everything else written to fd 1 by this process or by libraries
    (NCCL prints its version banner there when the box sets NCCL_DEBUG) has been redirected to stderr by main()."""
    sys.stdout.flush()
    if _STDOUT_FD is None:
        print(line, flush=True)
    else:
        os.write(_STDOUT_FD, (line + "\n").encode())


def main():
    global _STDOUT_FD  # pylint: disable=global-statement
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--npde", type=int, default=63488)
    ap.add_argument("--nbc-edge", type=int, default=512)
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--replicate", default="auto", choices=["auto", "yes", "no"],
                    help="N > 1 GPU: replicate the factor on every rank (auto: if it fits) or keep it distributed")
    ap.add_argument("--nb", type=int, default=0,
                    help="block-row height of the distributed Cholesky (N > 1 GPU); 0 = 1024 up to 2 GPUs, 512 beyond "
                         "(shorter panel chain / better balance, profiles/dist_cholesky_r01.txt)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
