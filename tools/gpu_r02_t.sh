#!/bin/bash
# round-2 GPU check T: 2 x 2 cluster patches (A and B multicast): correctness, burst timing, sustained A/B inside the bench step
mkdir -p gpurun_out
LPGP_OZAKI_CLUSTER=4 timeout -s KILL 240 python -m pytest tests/test_gpu_ozaki.py -x -q > gpurun_out/t_oz.log 2>&1; echo "rc=$?" >> gpurun_out/t_oz.log
for cl in 2 4; do timeout -s KILL 100 python tools/profile_ozaki.py 32768 1024 32768 7 $cl; done > gpurun_out/t_ab.log 2>&1
timeout -s KILL 100 python tools/profile_ozaki.py 4224 1152 8192 7 4 >> gpurun_out/t_ab.log 2>&1
tail -5 gpurun_out/t_oz.log; cat gpurun_out/t_ab.log
if grep -q "rc=0" gpurun_out/t_oz.log; then
for cl in 2 4; do
LPGP_OZAKI_CLUSTER=$cl timeout 300 python bench.py --gpus 1 --steps 4 --warmup 2 --budget-s 125 --cpu-budget-s 2 > gpurun_out/t_bench_cl$cl.json 2> gpurun_out/t_bench_cl$cl.err; echo "rc=$?" >> gpurun_out/t_bench_cl$cl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/t_bench_cl$cl.json').read().strip().splitlines()[-1])
print('cluster $cl: value',round(d['value'],3),'var',round(d['phases_ms']['var'],1),'kernel s',round(d['roofline']['kernel_seconds_per_step'],3),'TOPS',round(d['roofline']['achieved']),'clocks',d['clocks']['sm_mhz'],'steps',d['steps'])
PY
done
fi
