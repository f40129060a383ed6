#!/bin/bash
# phase-4 GPU check at HEAD: -m gpu suite, smoke(), default bench line + reference arm, ncu launch list of one step,
# ncu --set full of the hot kernels (new Gram kernel)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15 > gpurun_out/p4_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/p4_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/p4_bench.json 2> gpurun_out/p4_bench.err
timeout 400 python bench.py --impl reference > gpurun_out/p4_bench_ref.json 2> gpurun_out/p4_bench_ref.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01c.csv python tools/profile_step.py 15360 256 128 > gpurun_out/p4_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gram_sep|post_mean_sep' -c 3 -o gpurun_out/prof_gram_r01e -f python tools/profile_kernels.py > gpurun_out/p4_ncu.log 2>&1
timeout 200 python tools/perf_kernels.py gram > gpurun_out/p4_perf.log 2>&1
tail -15 gpurun_out/p4_tests.log; tail -2 gpurun_out/p4_smoke.log; cat gpurun_out/p4_bench.json; tail -3 gpurun_out/p4_bench.err; cat gpurun_out/p4_bench_ref.json; cat gpurun_out/p4_perf.log
