"""One emulated GEMM launch of the shape the variance solve issues most (m x kblock x K), for `ncu --set full`:
   ncu ... -k regex:ozaki_gemm python tools/profile_ozaki.py [m] [n] [k] [S]; also prints the CUDA-event time."""
import sys
sys.path.insert(0, ".")
import torch
from linpde_gp_b200 import backend as be

m = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
k = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
S = int(sys.argv[4]) if len(sys.argv) > 4 else 7
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn((m, k), generator=g, device="cuda", dtype=torch.float64)
B = torch.randn((n, k), generator=g, device="cuda", dtype=torch.float64)
C = be.alloc_matrix(m, n).zero_()
PA, PB = be.OzakiPlanes(m, k, S, 1024), be.OzakiPlanes(n, k, S, 1024)
PA.split(A)
PB.split(B)
del A, B
from linpde_gp_b200._lib import lib
cl = int(sys.argv[5]) if len(sys.argv) > 5 else 2
assert lib.lpgp_set_option(5, cl) == 0  # LPGP_OPT_OZAKI_CLUSTER
pl = int(sys.argv[6]) if len(sys.argv) > 6 else 1
assert lib.lpgp_set_option(6, pl) == 0  # LPGP_OPT_OZAKI_PAIR_LEVELS
cp = int(sys.argv[7]) if len(sys.argv) > 7 else 0
assert lib.lpgp_set_option(7, cp) == 0  # LPGP_OPT_OZAKI_CTA_PAIR
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    be.ozaki_gemm_nt(PA, PB, C, k, -1.0, 1.0)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    pairs = S * (S + 1) // 2
    print(f"ozaki gemm (cluster {cl}, levels per pass {1 + pl}, cta pair {cp}) {m}x{n}x{k} S={S}: {t:.2f} ms  {2.0 * m * n * k / t * 1e-9:.1f} TFLOP/s-equivalent  "
          f"{2.0 * m * n * k * pairs / t * 1e-12:.2f} INT8 POP/s")
