#!/bin/bash
# phase-8 GPU check: fused single-RHS substitution (tests + potrs timings), reference heat test
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -30 > gpurun_out/p8_tests.log
timeout 600 python tools/perf_kernels.py potrf potrf64 trsm > gpurun_out/p8_perf.log 2>&1
tail -30 gpurun_out/p8_tests.log | cut -c1-300; grep "potrs\|lookahead pipeline\]" gpurun_out/p8_perf.log
