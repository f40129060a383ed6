#!/bin/bash
# round-2 GPU check J: full -m gpu suite + smoke at HEAD, reference arm, ncu launch list of bench.py at a reduced size
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/j_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/j_smoke.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j_bench_ref.json 2> gpurun_out/j_bench_ref.err; echo "ref rc=$?" >> gpurun_out/j_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02j.csv python bench.py --npde 15872 --nbc-edge 128 --grid 256 --steps 1 --warmup 1 --cpu-budget-s 2 > gpurun_out/j_launch_bench.json 2> gpurun_out/j_launch.err
tail -6 gpurun_out/j_tests.log; tail -2 gpurun_out/j_smoke.log; cut -c1-600 gpurun_out/j_bench_ref.json; python tools/launch_summary.py gpurun_out/launches_r02j.csv | head -24
