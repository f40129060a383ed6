#!/bin/bash
# round-2 GPU check B: radial Matern family + at-size goldens + first contact of the tcgen05 (Ozaki) kernels
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_ozaki.py -x -q -k "split" > gpurun_out/b_oz_split.log 2>&1; echo "rc=$?" >> gpurun_out/b_oz_split.log
timeout -s KILL 120 python -m pytest tests/test_gpu_ozaki.py -x -q -k "gemm_matches and 128-128-1024" > gpurun_out/b_oz_gemm1.log 2>&1; echo "rc=$?" >> gpurun_out/b_oz_gemm1.log
timeout -s KILL 300 python -m pytest tests/test_gpu_ozaki.py -q > gpurun_out/b_oz_all.log 2>&1; echo "rc=$?" >> gpurun_out/b_oz_all.log
timeout -s KILL 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_ozaki.py > gpurun_out/b_tests.log 2>&1; echo "rc=$?" >> gpurun_out/b_tests.log
tail -5 gpurun_out/b_oz_split.log; tail -30 gpurun_out/b_oz_gemm1.log; tail -40 gpurun_out/b_oz_all.log; tail -30 gpurun_out/b_tests.log
