"""One device-level step of the bench workload at a profiler-friendly size (no CPU baseline, no timing):
   ncu ... python tools/profile_step.py [npde] [nbc_edge] [grid]"""
import sys
sys.path.insert(0, ".")
import torch
import bench

npde = int(sys.argv[1]) if len(sys.argv) > 1 else 15360
nbc = int(sys.argv[2]) if len(sys.argv) > 2 else 256
grid = int(sys.argv[3]) if len(sys.argv) > 3 else 128
prob = bench.make_problem(npde, nbc, grid)
ds = bench.DeviceSolve(prob, 0, 1)
m, v = ds.step()
torch.cuda.synchronize()
print("step done", float(m.abs().max()), float(v.min()), float(v.max()))
