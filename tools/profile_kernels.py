"""Launch each hot kernel once at a representative size (for `ncu --set full`):
   ncu ... python tools/profile_kernels.py"""
import sys
sys.path.insert(0, ".")
import torch
from linpde_gp_b200 import backend as be
from linpde_gp_b200._lowering import Factor1D, lower

ell = 0.03
fac = [Factor1D("matern", ell, nu=2.5), Factor1D("matern", ell, nu=2.5)]
lap = {(2, 0): -1.0, (0, 2): -1.0}
d_LkL, d_kL, d_k = lower(fac, lap, lap, 4.0), lower(fac, None, lap, 4.0), lower(fac, None, None, 4.0)
n = 16384
X = torch.rand(n, 2, dtype=torch.float64, device="cuda")
out = be.alloc_matrix(n, n)
be.gram(d_LkL, X, None, out=out)                    # gram_tile_kernel<2,3,false>  (268M entries)
be.gram(d_k, X, None, out=out, lower=True)
A = be.alloc_matrix(8192, 2048).normal_(); B = be.alloc_matrix(8192, 2048).normal_(); C = be.alloc_matrix(8192, 8192).zero_()
be.gemm_nt(A, B, C, -1.0, 1.0)                       # gemm_nt_kernel 8192 x 8192 x 2048 (4096 tiles)
P = be.alloc_matrix(n, 512).normal_()
be.gemm_nt(P, P, out, -1.0, 1.0, lower=True)         # SYRK-style trailing update, K = 512
Xt = torch.rand(65536, 2, dtype=torch.float64, device="cuda")
w = torch.randn(n, dtype=torch.float64, device="cuda")
be.post_mean(be.ObsBlocks([d_kL], [X], [0]), w, Xt)  # post_mean_kernel<2,3,false>
G = torch.randn(128, 128, dtype=torch.float64, device="cuda"); G = G @ G.T + 128 * torch.eye(128, dtype=torch.float64, device="cuda")
f = be.DeviceFactor([128]); f.L.copy_(G); f.potrf()  # potrf_leaf_kernel
be.row_sumsq(out)
torch.cuda.synchronize()
print("done")
