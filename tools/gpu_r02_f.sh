#!/bin/bash
# round-2 GPU check F: full -m gpu suite with the emulated variance solve as default, smoke, bench with the driver's arguments
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -x -q -m gpu > gpurun_out/f_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_smoke.log
SECONDS=0
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$? wall=${SECONDS}s" >> gpurun_out/f_bench.err
tail -5 gpurun_out/f_tests.log; tail -3 gpurun_out/f_smoke.log; cat gpurun_out/f_bench.json; tail -5 gpurun_out/f_bench.err
