#!/bin/bash
# round-2 GPU check E: NN GEMM / multi-RHS potrs / L2 projections
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm_nn or trsm_rln or potrs or trsm_and" > gpurun_out/e_nn.log 2>&1; echo "rc=$?" >> gpurun_out/e_nn.log
timeout -s KILL 600 python -m pytest tests/test_gpu_projections.py -q > gpurun_out/e_proj.log 2>&1; echo "rc=$?" >> gpurun_out/e_proj.log
tail -40 gpurun_out/e_nn.log; tail -60 gpurun_out/e_proj.log
