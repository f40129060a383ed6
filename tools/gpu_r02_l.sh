#!/bin/bash
# round-2 GPU check L (8 GPUs): BASELINE.json configs[4], N = 131,072 (137 GB Gram), factor left distributed, emulated streamed variance
mkdir -p gpurun_out
for NB in 1024 512; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953${NB:0:1} bench.py --gpus 8 --npde 126976 --nbc-edge 1024 --steps 2 --warmup 1 --nb $NB > gpurun_out/l_bench_c5_nb$NB.json 2> gpurun_out/l_bench_c5_nb$NB.err; echo "rc=$?" >> gpurun_out/l_bench_c5_nb$NB.err
done
for NB in 1024 512; do cut -c1-1500 gpurun_out/l_bench_c5_nb$NB.json; tail -4 gpurun_out/l_bench_c5_nb$NB.err; done
