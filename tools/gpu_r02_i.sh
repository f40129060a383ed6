#!/bin/bash
# round-2 GPU check I (8 GPUs): timeline of the distributed Cholesky at N = 64k / 128k, multi-GPU pytest, bench with the driver's arguments at 8 and 4 GPUs
mkdir -p gpurun_out
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/dist_bench.py 65536 512 256 > gpurun_out/i_distbench_8_64k.log 2>&1
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/dist_bench.py 131072 512 > gpurun_out/i_distbench_8_128k.log 2>&1
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/i_bench_8.json 2> gpurun_out/i_bench_8.err; echo "bench rc=$? wall=${SECONDS}s" >> gpurun_out/i_bench_8.err
SECONDS=0
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/i_bench_4.json 2> gpurun_out/i_bench_4.err; echo "bench rc=$? wall=${SECONDS}s" >> gpurun_out/i_bench_4.err
timeout -s KILL 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/i_multi.log 2>&1; echo "rc=$?" >> gpurun_out/i_multi.log
grep -v "^\[\|^\*\|NCCL\|OMP" gpurun_out/i_distbench_8_64k.log | tail -8; grep -v "^\[\|^\*\|NCCL\|OMP" gpurun_out/i_distbench_8_128k.log | tail -6
cat gpurun_out/i_bench_8.json | cut -c1-1800; tail -2 gpurun_out/i_bench_8.err; cat gpurun_out/i_bench_4.json | cut -c1-300; tail -2 gpurun_out/i_bench_4.err; tail -3 gpurun_out/i_multi.log
