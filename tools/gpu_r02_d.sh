#!/bin/bash
# round-2 GPU check D: full -m gpu suite, ncu of the emulated GEMM, budgeted bench lines with the DMMA and the emulated variance solve
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -x -q -m gpu > gpurun_out/d_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_tests.log
timeout -s KILL 300 python tools/profile_ozaki.py 16384 1024 16384 7 > gpurun_out/d_ozgemm.log 2>&1
timeout -s KILL 300 python tools/profile_ozaki.py 32768 1024 32768 7 >> gpurun_out/d_ozgemm.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_gemm -s 1 -c 1 -f -o gpurun_out/prof_ozaki_r02d python tools/profile_ozaki.py 16384 1024 16384 7 > gpurun_out/d_ncu.log 2>&1
LPGP_OZAKI_SLICES=7 timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --budget-s 150 > gpurun_out/d_bench_oz7.json 2> gpurun_out/d_bench_oz7.err; echo "bench rc=$?" >> gpurun_out/d_bench_oz7.err
tail -5 gpurun_out/d_tests.log; cat gpurun_out/d_ozgemm.log; tail -3 gpurun_out/d_ncu.log; cat gpurun_out/d_bench_oz7.json; tail -5 gpurun_out/d_bench_oz7.err
