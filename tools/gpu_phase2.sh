#!/bin/bash
# phase-2 GPU check: Kronecker path tests + full kernel/API suites, kron/gram timings, ncu of the kron kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -x -q 2>&1 | tail -25 > gpurun_out/p2_tests.log
timeout 300 python tools/perf_kernels.py kron gram > gpurun_out/p2_perf.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kron_sum -c 3 -o gpurun_out/prof_kron_r01 -f python tools/perf_kernels.py kron > gpurun_out/p2_ncu.log 2>&1
tail -25 gpurun_out/p2_tests.log; cat gpurun_out/p2_perf.log
