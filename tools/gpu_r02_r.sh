#!/bin/bash
# round-2 GPU check R: cluster-of-2 (B-tile multicast) emulated GEMM: correctness, then A/B timing
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_gpu_ozaki.py -x -q > gpurun_out/r_oz.log 2>&1; echo "rc=$?" >> gpurun_out/r_oz.log
for cl in 1 2; do timeout -s KILL 100 python tools/profile_ozaki.py 32768 1024 32768 7 $cl; done > gpurun_out/r_ab.log 2>&1
timeout -s KILL 100 python tools/profile_ozaki.py 4224 1024 8192 7 2 >> gpurun_out/r_ab.log 2>&1
tail -25 gpurun_out/r_oz.log; cat gpurun_out/r_ab.log
