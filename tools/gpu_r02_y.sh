#!/bin/bash
# round-2 GPU check Y (2 GPUs): multi-GPU pytest at HEAD (CTA-pair emulated GEMM under torchrun: replicated and streamed variance)
mkdir -p gpurun_out
timeout -s KILL 100 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/y_multi.log 2>&1; echo "rc=$?" >> gpurun_out/y_multi.log
tail -5 gpurun_out/y_multi.log
