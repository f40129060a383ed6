#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_ozaki.py -q > gpurun_out/c_oz_all.log 2>&1; echo "rc=$?" >> gpurun_out/c_oz_all.log
timeout -s KILL 600 python tools/ozaki_perf.py 16384 16384 > gpurun_out/c_oz_perf16k.log 2>&1
timeout -s KILL 900 python tools/ozaki_perf.py 32768 32768 > gpurun_out/c_oz_perf32k.log 2>&1
tail -30 gpurun_out/c_oz_all.log; cat gpurun_out/c_oz_perf16k.log gpurun_out/c_oz_perf32k.log
