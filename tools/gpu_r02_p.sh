#!/bin/bash
# round-2 GPU check P (final): full -m gpu suite + smoke at HEAD
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/p_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/p_smoke.log
tail -8 gpurun_out/p_tests.log; tail -2 gpurun_out/p_smoke.log
