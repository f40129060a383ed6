#!/bin/bash
# phase-6 GPU check: condition-gated refinement of the panel solves (tests + potrf timings A/B)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -30 > gpurun_out/p6_tests.log
timeout 600 python tools/perf_kernels.py potrf potrf64 > gpurun_out/p6_perf.log 2>&1
tail -30 gpurun_out/p6_tests.log; cat gpurun_out/p6_perf.log
