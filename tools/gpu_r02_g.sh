#!/bin/bash
# round-2 GPU check G: plane-reuse emulated GEMM (v2) correctness + A/B timing; Kronecker structured solve tests
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_ozaki.py -x -q > gpurun_out/g_oz.log 2>&1; echo "rc=$?" >> gpurun_out/g_oz.log
for v in 1 2; do for S in 7 6; do timeout -s KILL 120 python tools/profile_ozaki.py 32768 1024 32768 $S $v; done; done > gpurun_out/g_ab.log 2>&1
timeout -s KILL 300 python -m pytest tests/test_gpu_linops.py -x -q -k kronecker > gpurun_out/g_kron.log 2>&1; echo "rc=$?" >> gpurun_out/g_kron.log
tail -30 gpurun_out/g_oz.log; cat gpurun_out/g_ab.log; tail -30 gpurun_out/g_kron.log
