#!/bin/bash
# round-2 GPU check S: sustained effect of the cluster-of-2 emulated GEMM inside the bench step (A/B)
mkdir -p gpurun_out
for cl in 1 2; do
LPGP_OZAKI_CLUSTER=$cl timeout 300 python bench.py --gpus 1 --steps 4 --warmup 2 --budget-s 125 --cpu-budget-s 2 > gpurun_out/s_bench_cl$cl.json 2> gpurun_out/s_bench_cl$cl.err; echo "rc=$?" >> gpurun_out/s_bench_cl$cl.err
done
for cl in 1 2; do python - <<PY
import json
d=json.loads(open('gpurun_out/s_bench_cl$cl.json').read().strip().splitlines()[-1])
print('cluster $cl: value',round(d['value'],3),'var',round(d['phases_ms']['var'],1),'kernel s',round(d['roofline']['kernel_seconds_per_step'],3),'TOPS',round(d['roofline']['achieved']),'clocks',d['clocks']['sm_mhz'],'steps',d['steps'])
PY
done
