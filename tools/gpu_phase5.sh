#!/bin/bash
# phase-5 GPU check: refined panel solves (tests + potrf timings A/B), block linops tests
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -30 > gpurun_out/p5_tests.log
timeout 600 python tools/perf_kernels.py potrf potrf64 trsm > gpurun_out/p5_perf.log 2>&1
tail -30 gpurun_out/p5_tests.log; cat gpurun_out/p5_perf.log
