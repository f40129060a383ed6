"""cuBLAS DGEMM / DSYRK-like / DPOTRF / DTRSM reference rates on the box (library numbers, context for roofline)."""
import json, time, torch
dev = "cuda:0"
res = {}
def tm(f, reps=5):
    f(); f(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for n in (4096, 8192, 16384):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty_like(a)
    ms = tm(lambda: torch.matmul(a, b, out=c))
    res[f"dgemm_{n}"] = 2 * n**3 / ms * 1e-9
    print(f"cuBLAS DGEMM {n}^3: {res[f'dgemm_{n}']:.2f} TFLOP/s ({ms:.2f} ms)", flush=True)
# sustained
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev); c = torch.empty_like(a)
torch.cuda.synchronize(); t0 = time.time(); k = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
while time.time() - t0 < 4.0:
    for _ in range(10): torch.matmul(a, b, out=c)
    k += 10; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
res["dgemm_8192_sustained"] = 2 * n**3 * k / e0.elapsed_time(e1) * 1e-9
print(f"cuBLAS DGEMM 8192^3 sustained 4 s: {res['dgemm_8192_sustained']:.2f} TFLOP/s", flush=True)
# K=512 rank update shape (what the Cholesky trailing update looks like)
for (m, k_) in ((16384, 512), (32768, 512), (32768, 256)):
    p = torch.randn(m, k_, dtype=torch.float64, device=dev); cc = torch.zeros(m, m, dtype=torch.float64, device=dev)
    ms = tm(lambda: torch.addmm(cc, p, p.t(), beta=1.0, alpha=-1.0, out=cc), reps=3)
    res[f"gemm_rank{k_}_m{m}"] = 2 * m * m * k_ / ms * 1e-9
    print(f"cuBLAS C-=P P^T m={m} k={k_}: {res[f'gemm_rank{k_}_m{m}']:.2f} TFLOP/s (full square)", flush=True)
    del p, cc
for n in (16384, 32768):
    x = torch.randn(n, n, dtype=torch.float64, device=dev)
    g = x @ x.t() / n + torch.eye(n, dtype=torch.float64, device=dev) * 2; del x
    ms = tm(lambda: torch.linalg.cholesky(g), reps=2)
    res[f"potrf_{n}"] = n**3 / 3 / ms * 1e-9
    print(f"cuSOLVER(torch) cholesky n={n}: {res[f'potrf_{n}']:.2f} TFLOP/s ({ms:.1f} ms)", flush=True)
    L = torch.linalg.cholesky(g); del g
    B = torch.randn(n, 8192, dtype=torch.float64, device=dev)
    ms = tm(lambda: torch.linalg.solve_triangular(L, B, upper=False), reps=2)
    res[f"trsm_{n}x8192"] = n * n * 8192 / ms * 1e-9
    print(f"cuBLAS trsm n={n} nrhs=8192: {res[f'trsm_{n}x8192']:.2f} TFLOP/s ({ms:.1f} ms)", flush=True)
    del L, B
json.dump(res, open("gpurun_out/fp64_cublas.json", "w"), indent=1)
