#!/bin/bash
# round-2 GPU check A: -m gpu suite, smoke(), a short budgeted bench line (driver-style arguments)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --budget-s 200 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
tail -8 gpurun_out/a_tests.log; tail -2 gpurun_out/a_smoke.log; cat gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
