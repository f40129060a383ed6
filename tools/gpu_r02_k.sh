#!/bin/bash
# round-2 GPU check K (2 GPUs): Kronecker-factor / projection tests, extended smoke, multi-GPU pytest (emulated streamed solve)
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_projections.py tests/test_gpu_linops.py -x -q -k "kronecker or gridded or projection" > gpurun_out/k_tests.log 2>&1; echo "rc=$?" >> gpurun_out/k_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/k_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/k_smoke.log
timeout -s KILL 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/k_multi.log 2>&1; echo "rc=$?" >> gpurun_out/k_multi.log
tail -30 gpurun_out/k_tests.log; tail -3 gpurun_out/k_smoke.log; tail -30 gpurun_out/k_multi.log
