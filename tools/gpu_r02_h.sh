#!/bin/bash
# round-2 GPU check H (2 GPUs): multi-GPU pytest, per-stage timeline of the distributed Cholesky, bench with the driver's arguments
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout -s KILL 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/h_multi.log 2>&1; echo "rc=$?" >> gpurun_out/h_multi.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/dist_bench.py 65536 1024 512 > gpurun_out/h_distbench_${NG}.log 2>&1
SECONDS=0
timeout 870 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/h_bench_${NG}.json 2> gpurun_out/h_bench_${NG}.err; echo "bench rc=$? wall=${SECONDS}s" >> gpurun_out/h_bench_${NG}.err
tail -5 gpurun_out/h_multi.log; grep -v "^\[" gpurun_out/h_distbench_${NG}.log | tail -8; cat gpurun_out/h_bench_${NG}.json; tail -3 gpurun_out/h_bench_${NG}.err
