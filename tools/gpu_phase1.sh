#!/bin/bash
# phase-1 GPU check: kernel/API parity tests, kernel timings (lookahead A/B), ncu full capture of the hot kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -x -q 2>&1 | tail -15 > gpurun_out/p1_tests.log
timeout 600 python tools/perf_kernels.py gemm potrf potrf64 trsm > gpurun_out/p1_perf.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -o gpurun_out/prof_kernels_r01d -f python tools/profile_kernels.py > gpurun_out/p1_ncu.log 2>&1
tail -5 gpurun_out/p1_tests.log; cat gpurun_out/p1_perf.log
