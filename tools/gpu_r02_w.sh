#!/bin/bash
# round-2 GPU check W: CTA-pair (tcgen05 cta_group::2) variant of the emulated GEMM -- correctness first (short timeout: a
# protocol error would hang), then burst / in-step A/B; ncu --set full of the kernel(s) at HEAD
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1, device='cuda'); print('warm')" > gpurun_out/w_warm.log 2>&1
timeout -s KILL 90 python -m pytest tests/test_gpu_ozaki.py -x -q -k cta_pair > gpurun_out/w_pair.log 2>&1; echo "rc=$?" >> gpurun_out/w_pair.log
tail -30 gpurun_out/w_pair.log | cut -c1-400
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:ozaki_gemm -s 1 -c 1 -f -o gpurun_out/prof_ozaki_r02w python tools/profile_ozaki.py 16384 1024 16384 7 2 1 0 > gpurun_out/w_ncu.log 2>&1
tail -3 gpurun_out/w_ncu.log
if grep -q "rc=0" gpurun_out/w_pair.log; then
for cp in 0 1; do timeout -s KILL 100 python tools/profile_ozaki.py 32768 1024 32768 7 2 1 $cp; done > gpurun_out/w_ab.log 2>&1
cat gpurun_out/w_ab.log
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:ozaki_gemm -s 1 -c 1 -f -o gpurun_out/prof_ozaki_r02w_pair python tools/profile_ozaki.py 16384 1024 16384 7 2 1 1 > gpurun_out/w_ncu_pair.log 2>&1
LPGP_OZAKI_CTA_PAIR=1 timeout 300 python bench.py --gpus 1 --steps 2 --warmup 1 --budget-s 80 --cpu-budget-s 2 > gpurun_out/w_bench_pair.json 2> gpurun_out/w_bench_pair.err; echo "rc=$?" >> gpurun_out/w_bench_pair.err
python - <<PY
import json
d=json.loads(open('gpurun_out/w_bench_pair.json').read().strip().splitlines()[-1])
print('cta pair: value',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'var',round(d['phases_ms']['var'],1),'kernel s',round(d['roofline']['kernel_seconds_per_step'],3),'TOPS',round(d['roofline']['achieved']),'frac',round(d['roofline']['frac'],3),'clocks',d['clocks']['sm_mhz'],'steps',d['steps'], 'checks', d['checks'])
PY
fi
