#!/bin/bash
# phase-9 GPU check at HEAD: -m gpu suite, smoke(), default bench line + reference arm, ncu launch list of bench.py itself
# at a reduced size (N = 16,384, 128^2 grid: one warm-up + one timed step), kernel timings
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15 > gpurun_out/p9_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/p9_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/p9_bench.json 2> gpurun_out/p9_bench.err
timeout 400 python bench.py --impl reference > gpurun_out/p9_bench_ref.json 2> gpurun_out/p9_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01d.csv python bench.py --npde 15360 --nbc-edge 256 --grid 128 --steps 1 --warmup 1 > gpurun_out/p9_launch_bench.json 2> gpurun_out/p9_launch.err
timeout 300 python tools/perf_kernels.py gemm trsm potrf > gpurun_out/p9_perf.log 2>&1
tail -15 gpurun_out/p9_tests.log; tail -2 gpurun_out/p9_smoke.log; cat gpurun_out/p9_bench.json; tail -3 gpurun_out/p9_bench.err; cat gpurun_out/p9_bench_ref.json; cat gpurun_out/p9_perf.log
