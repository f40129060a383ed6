#!/bin/bash
# round-2 GPU check Z (2 GPUs): bench.py at HEAD launched the way the driver launches it (short: 1 warm-up + 2 timed steps)
mkdir -p gpurun_out
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --budget-s 45 --cpu-budget-s 1 > gpurun_out/z_bench_2.json 2> gpurun_out/z_bench_2.err; echo "rc=$?" >> gpurun_out/z_bench_2.err
tail -2 gpurun_out/z_bench_2.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/z_bench_2.json').read().strip().splitlines()[-1])
print('2 GPUs: value',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'phases',{k:round(v,1) for k,v in d['phases_ms'].items()},'TOPS',round(d['roofline']['achieved']),'clocks',d['clocks']['sm_mhz'],'steps',d['steps'],d['warmup'])
PY
