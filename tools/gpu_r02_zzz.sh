#!/bin/bash
# round-2 GPU check ZZZ: bench.py at HEAD on a reduced configuration (N = 16,384, 256^2 grid): the bench code path after the last host-side changes
mkdir -p gpurun_out
timeout 52 python bench.py --npde 15872 --nbc-edge 128 --grid 256 --steps 3 --warmup 3 --cpu-budget-s 1 > gpurun_out/zzz_bench_small.json 2> gpurun_out/zzz_bench_small.err; echo "rc=$?" >> gpurun_out/zzz_bench_small.err
tail -n 2 gpurun_out/zzz_bench_small.err | cut -c1-200; cut -c1-700 gpurun_out/zzz_bench_small.json
