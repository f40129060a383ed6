#!/bin/bash
# round-2 GPU check O (1 GPU): opt-in 6-plane variance solve at C4; configs[1] (N = 16,384) bench line
mkdir -p gpurun_out
LPGP_OZAKI_SLICES=6 timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --budget-s 130 --cpu-budget-s 3 > gpurun_out/o_bench_s6.json 2> gpurun_out/o_bench_s6.err; echo "rc=$?" >> gpurun_out/o_bench_s6.err
timeout 300 python bench.py --gpus 1 --npde 15360 --nbc-edge 256 --steps 20 --warmup 5 --cpu-budget-s 3 > gpurun_out/o_bench_c2.json 2> gpurun_out/o_bench_c2.err; echo "rc=$?" >> gpurun_out/o_bench_c2.err
cut -c1-400 gpurun_out/o_bench_s6.json; tail -2 gpurun_out/o_bench_s6.err; cut -c1-400 gpurun_out/o_bench_c2.json; tail -2 gpurun_out/o_bench_c2.err
