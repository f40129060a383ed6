"""Small run of the factor / solve kernels for compute-sanitizer (memcheck, racecheck):
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import ctypes
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from linpde_gp_b200 import _lib, backend as be

torch.manual_seed(0)
for refine in (1, 3):
    _lib.lib.lpgp_set_option(_lib.OPT_TRSM_REFINE, refine)
    sizes = (300, 46, 130)  # ragged leaves, three segments
    n = sum(sizes)
    X = torch.randn(n, n + 8, dtype=torch.float64, device="cuda")
    G = X @ X.T / n + 0.5 * torch.eye(n, dtype=torch.float64, device="cuda")
    f, off = None, 0
    for s in sizes:
        f = be.DeviceFactor([s]) if f is None else f.extended(s)
        f.L[off : off + s, : off + s].copy_(G[off : off + s, : off + s])
        f.potrf() if off == 0 else f.append_last()
        off += s
    L = torch.tril(f.L)
    err = (L @ L.T - G).abs().max().item()
    b = torch.randn(2, n, dtype=torch.float64, device="cuda")
    x = f.potrs(b.clone())
    res = (x @ G - b).abs().max().item()
    Y = be.alloc_matrix(37, n).normal_()
    Y0 = Y.clone()
    st = f._struct()
    _lib.check(_lib.lib.lpgp_trsm_rlt_refined(ctypes.byref(st), n, ctypes.c_void_p(Y.data_ptr()), 37, Y.stride(0), be._stream()), "trsm")
    res2 = (Y @ L.T - Y0).abs().max().item()
    print(f"refine={refine}: |LL^T-G|={err:.2e} potrs residual={res:.2e} trsm residual={res2:.2e}")
    assert err < 1e-12 and res < 1e-10 and res2 < 1e-10
_lib.lib.lpgp_set_option(_lib.OPT_TRSM_REFINE, 1)
torch.cuda.synchronize()
print("done")
