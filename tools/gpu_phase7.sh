#!/bin/bash
# phase-7 GPU check (2 GPUs): distributed == single-GPU posterior, distributed factor timings, 2-GPU bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/dist_check.py > gpurun_out/p7_check.log 2>&1
timeout 300 $TR tools/dist_bench.py 65536 1024 > gpurun_out/p7_distbench.log 2>&1
timeout 600 $TR bench.py --gpus 2 > gpurun_out/p7_bench2.json 2> gpurun_out/p7_bench2.err
tail -5 gpurun_out/p7_check.log; tail -6 gpurun_out/p7_distbench.log; cat gpurun_out/p7_bench2.json; tail -3 gpurun_out/p7_bench2.err
