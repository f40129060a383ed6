"""Timing of the INT8-emulated triangular solve against the DMMA path (development aid; bench.py is the contract)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from linpde_gp_b200 import backend as be  # noqa: E402


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    g = torch.Generator(device="cuda").manual_seed(0)
    f = be.DeviceFactor([n], reserve_rows=0)
    # a well-conditioned SPD matrix without an O(n^3) host step: diagonally dominant random symmetric
    f.L.copy_(torch.randn((n, n), generator=g, device="cuda", dtype=torch.float64) * 0.01)
    f.L.diagonal().add_(2.0 + 0.01 * np.sqrt(n))
    t0 = time.perf_counter()
    f.potrf()
    torch.cuda.synchronize()
    print(f"potrf n={n}: {time.perf_counter() - t0:.2f} s")
    X0 = torch.randn((m, n), generator=g, device="cuda", dtype=torch.float64)
    X = be.alloc_matrix(m, n)
    flops = float(m) * n * n

    def dmma():
        X.copy_(X0)
        f.trsm_rlt(X)

    t_copy = timed(lambda: X.copy_(X0))
    t = timed(dmma) - t_copy
    Xd = X.clone()
    print(f"DMMA trsm  m={m} n={n}: {t:9.1f} ms  {flops / t * 1e-9:7.2f} TFLOP/s")
    for kblock in (512, 1024):
        for S in (6, 7):
            XP = be.OzakiPlanes(m, n, S, kblock)
            t_split = timed(lambda: f.__dict__.pop("_ozaki_cache", None) or f.ozaki_planes(S, kblock), reps=1)

            def oz():
                X.copy_(X0)
                f.trsm_rlt_ozaki(X, S, kblock, XP)

            t = timed(oz) - t_copy
            err = float((X - Xd).abs().max() / Xd.abs().max())
            print(f"INT8 S={S} kblock={kblock}: {t:9.1f} ms  {flops / t * 1e-9:7.2f} TFLOP/s-equivalent  rel diff vs DMMA {err:.1e}"
                  f"  (factor split {t_split:.1f} ms)")
            del XP


if __name__ == "__main__":
    main()
