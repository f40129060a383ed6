#!/bin/bash
# round-2 GPU check N (final, 1 GPU): full -m gpu suite + smoke at HEAD, K-block 512 vs 1024 timing, short budgeted bench line
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/n_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/n_smoke.log
timeout -s KILL 300 python tools/ozaki_perf.py 32768 32768 > gpurun_out/n_oz_kblock.log 2>&1
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --budget-s 170 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?" >> gpurun_out/n_bench.err
tail -6 gpurun_out/n_tests.log; tail -2 gpurun_out/n_smoke.log; cat gpurun_out/n_oz_kblock.log; cut -c1-700 gpurun_out/n_bench.json; tail -2 gpurun_out/n_bench.err
