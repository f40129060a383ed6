"""Ad-hoc kernel timings on the B200 (development aid; bench.py is the contract)."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from linpde_gp_b200 import backend as be
from linpde_gp_b200._lowering import Factor1D, lower

def tm(f, reps=3, warm=1):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

res = {}
which = sys.argv[1:] or ["gram", "gemm", "potrf", "trsm", "mean"]
ell = 0.03
fac = [Factor1D("matern", ell, nu=2.5), Factor1D("matern", ell, nu=2.5)]
lap = {(2, 0): -1.0, (0, 2): -1.0}
d_LkL, d_kL, d_k = lower(fac, lap, lap, 4.0), lower(fac, None, lap, 4.0), lower(fac, None, None, 4.0)
if "gram" in which:
    for n in (16384, 32768):
        X = torch.rand(n, 2, dtype=torch.float64, device="cuda")
        out = be.alloc_matrix(n, n)
        for name, dsc in (("LkL", d_LkL), ("kL", d_kL), ("k", d_k)):
            ms = tm(lambda: be.gram(dsc, X, None, out=out))
            print(f"gram {name} n={n} full: {ms:.2f} ms  {n*n/ms*1e-6:.1f} Gentries/s  ({n*n*8/ms*1e-6:.0f} GB/s)", flush=True)
        ms = tm(lambda: be.gram(d_LkL, X, None, out=out, lower=True))
        print(f"gram LkL n={n} lower: {ms:.2f} ms  {n*(n+1)/2/ms*1e-6:.1f} Gentries/s (lower entries)", flush=True)
        del out
if "gemm" in which:
    for (m, n, k, low) in ((8192, 8192, 8192, False), (16384, 16384, 512, False), (32768, 32768, 512, True), (32768, 32768, 2048, True), (8192, 32768, 16384, False)):
        A = be.alloc_matrix(m, k).normal_(); B = be.alloc_matrix(n, k).normal_(); C = be.alloc_matrix(m, n).zero_()
        ms = tm(lambda: be.gemm_nt(A, B, C, -1.0, 1.0, lower=low))
        fl = 2.0 * m * n * k * (0.5 if low else 1.0)
        print(f"gemm_nt m={m} n={n} k={k} lower={low}: {ms:.2f} ms  {fl/ms*1e-9:.2f} TFLOP/s", flush=True)
        del A, B, C
if "potrf" in which or "potrf64" in which:
    from linpde_gp_b200 import _lib
    sizes = (8192, 16384, 32768) if "potrf" in which else ()
    if "potrf64" in which:
        sizes = sizes + (65536,)
    for n in sizes:
        f = be.DeviceFactor([n])
        # SPD test matrix built in place, block-wise (no second n x n buffer at n = 65536)
        G = be.alloc_matrix(n, n)
        Xs = torch.randn(n, 512, dtype=torch.float64, device="cuda")
        torch.mm(Xs, Xs.T, out=G)  # n is a multiple of 16: G is contiguous
        G.mul_(1.0 / 512); G.diagonal().add_(2.0); del Xs
        def run():
            f.L.copy_(G); f.potrf()
        tcopy = tm(lambda: f.L.copy_(G))
        for flag, refine, label in ((1, 1, "one-stream recursion"), (0, 0, "lookahead pipeline, unrefined panel solves"),
                                    (0, 3, "lookahead pipeline, every leaf refined"), (0, 1, "lookahead pipeline")):
            _lib.lib.lpgp_set_option(_lib.OPT_NO_LOOKAHEAD, flag)
            _lib.lib.lpgp_set_option(_lib.OPT_TRSM_REFINE, refine)
            ms = tm(run, reps=2) - tcopy
            print(f"potrf n={n} [{label}]: {ms:.1f} ms  {n**3/3/ms*1e-9:.2f} TFLOP/s", flush=True)
        _lib.lib.lpgp_set_option(_lib.OPT_NO_LOOKAHEAD, 0)
        _lib.lib.lpgp_set_option(_lib.OPT_TRSM_REFINE, 1)
        if "trsm" in which:
            m = 8192
            Xr = be.alloc_matrix(m, n).normal_()
            ms = tm(lambda: f.trsm_rlt(Xr), reps=2)
            print(f"trsm_rlt m={m} n={n}: {ms:.1f} ms  {m*n*n/ms*1e-9:.2f} TFLOP/s", flush=True)
            b = torch.randn(1, n, dtype=torch.float64, device="cuda")
            ms = tm(lambda: f.potrs(b), reps=2)
            print(f"potrs n={n}: {ms:.1f} ms", flush=True)
            del Xr
        del G, f
if "mean" in which:
    n, m = 16384, 262144
    X = torch.rand(n, 2, dtype=torch.float64, device="cuda"); Xt = torch.rand(m, 2, dtype=torch.float64, device="cuda")
    w = torch.randn(n, dtype=torch.float64, device="cuda")
    blocks = be.ObsBlocks([d_kL], [X], [0])
    ms = tm(lambda: be.post_mean(blocks, w, Xt))
    print(f"post_mean m={m} n={n}: {ms:.2f} ms  {m*n/ms*1e-6:.1f} Gevals/s", flush=True)
if "kron" in which:
    # tensor-grid path: Gram block of -Laplace k -Laplace on a g1 x g2 collocation grid, 4 Kronecker terms
    for (g1, g2) in ((128, 128), (256, 128), (256, 256)):
        n = g1 * g2
        terms = [(1.0 + t, be.alloc_matrix(g1, g1).normal_(), be.alloc_matrix(g2, g2).normal_()) for t in range(4)]
        out = be.alloc_matrix(n, n)
        ms = tm(lambda: be.kron_sum(terms, out=out))
        print(f"kron_sum 4 terms n={n} full: {ms:.2f} ms  {n*n/ms*1e-6:.1f} Gentries/s  ({n*n*8/ms*1e-6:.0f} GB/s)", flush=True)
        ms = tm(lambda: be.kron_sum(terms[:1], out=out))
        print(f"kron_sum 1 term  n={n} full: {ms:.2f} ms  {n*n/ms*1e-6:.1f} Gentries/s  ({n*n*8/ms*1e-6:.0f} GB/s)", flush=True)
        ms = tm(lambda: be.kron_sum(terms, out=out, lower=True))
        print(f"kron_sum 4 terms n={n} lower: {ms:.2f} ms  {n*(n+1)/2/ms*1e-6:.1f} Gentries/s (lower entries)", flush=True)
        del out, terms
