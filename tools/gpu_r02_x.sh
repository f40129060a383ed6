#!/bin/bash
# round-2 GPU check X: CTA-pair kernel as the default -- full -m gpu suite, smoke, bench line at HEAD
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -x -q -m gpu > gpurun_out/x_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_tests.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/x_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/x_smoke.log
tail -3 gpurun_out/x_tests.log; tail -2 gpurun_out/x_smoke.log
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --budget-s 120 --cpu-budget-s 2 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err; echo "rc=$?" >> gpurun_out/x_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/x_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'var',round(d['phases_ms']['var'],1),'kernel s',round(d['roofline']['kernel_seconds_per_step'],3),'TOPS',round(d['roofline']['achieved']),'frac',round(d['roofline']['frac'],3),'clocks',d['clocks']['sm_mhz'],'steps',d['steps'],d['warmup'])
PY
