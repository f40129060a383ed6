"""One launch of the Gram assembly kernel at n = 32768 (for `ncu --set full --import-source on -k regex:gram_sep`)."""
import sys
sys.path.insert(0, ".")
import torch
from linpde_gp_b200 import backend as be
from linpde_gp_b200._lowering import Factor1D, lower

ell = 0.03
fac = [Factor1D("matern", ell, nu=2.5), Factor1D("matern", ell, nu=2.5)]
lap = {(2, 0): -1.0, (0, 2): -1.0}
d_LkL = lower(fac, lap, lap, 4.0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
X = torch.rand(n, 2, dtype=torch.float64, device="cuda")
out = be.alloc_matrix(n, n)
for _ in range(3):
    be.gram(d_LkL, X, None, out=out)
torch.cuda.synchronize()
print("done")
