"""One single-RHS solve at N = 32768 for ncu (fused substitution kernels):
   ncu --set full -k regex:fwd_step --launch-skip 100 -c 1 python tools/profile_potrs.py"""
import sys
sys.path.insert(0, ".")
import torch
from linpde_gp_b200 import backend as be

n = 32768
f = be.DeviceFactor([n])
Xs = torch.randn(n, 512, dtype=torch.float64, device="cuda")
torch.mm(Xs, Xs.T, out=f.L)
f.L.mul_(1.0 / 512); f.L.diagonal().add_(2.0); del Xs
f.potrf()
b = torch.randn(1, n, dtype=torch.float64, device="cuda")
f.potrs(b)
torch.cuda.synchronize()
print("done")
