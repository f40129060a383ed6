#!/bin/bash
# round-2 GPU check M (8 GPUs): 1-D block-row cyclic vs 2-D block-cyclic Cholesky, N = 64k and 128k
mkdir -p gpurun_out
timeout -s KILL 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/dist_bench2d.py 65536 512 2:4 4:2 8:1 > gpurun_out/m_2d_64k.log 2>&1
timeout -s KILL 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/dist_bench2d.py 131072 512 2:4 > gpurun_out/m_2d_128k.log 2>&1
grep -v "^\[\|^\*\|NCCL\|OMP\|^$" gpurun_out/m_2d_64k.log | tail -12; grep -v "^\[\|^\*\|NCCL\|OMP\|^$" gpurun_out/m_2d_128k.log | tail -12
