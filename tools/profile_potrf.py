"""One potrf at size N (after a warm-up) -- for an ncu launch list:  ncu --metrics gpu__time_duration.sum ... python tools/profile_potrf.py N"""
import sys
sys.path.insert(0, ".")
import torch
from linpde_gp_b200 import backend as be
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
X = torch.randn(n, n, dtype=torch.float64, device="cuda")
G = X @ X.T / n
del X
G.diagonal().add_(2.0)
f = be.DeviceFactor([n])
for rep in range(2):
    f.L.copy_(G)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f.potrf(); e1.record(); torch.cuda.synchronize()
    print(f"potrf n={n}: {e0.elapsed_time(e1):.2f} ms")
