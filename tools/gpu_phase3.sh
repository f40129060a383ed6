#!/bin/bash
# phase-3 GPU check: the whole -m gpu suite at HEAD, smoke(), the default 1-GPU bench line and the reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15 > gpurun_out/p3_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/p3_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/p3_bench.json 2> gpurun_out/p3_bench.err
tail -15 gpurun_out/p3_tests.log; tail -2 gpurun_out/p3_smoke.log; cat gpurun_out/p3_bench.json; tail -3 gpurun_out/p3_bench.err
