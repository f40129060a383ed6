#!/bin/bash
# round-2 GPU check V: paired digit levels in the emulated GEMM -- correctness (both orders bit-identical), burst timing A/B,
# then (only if green) the full -m gpu suite, smoke and a bench line at HEAD
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_ozaki.py -x -q > gpurun_out/v_oz.log 2>&1; echo "rc=$?" >> gpurun_out/v_oz.log
tail -4 gpurun_out/v_oz.log
if grep -q "rc=0" gpurun_out/v_oz.log; then for pl in 1 0; do timeout -s KILL 100 python tools/profile_ozaki.py 32768 1024 32768 7 2 $pl; done > gpurun_out/v_ab.log 2>&1; fi
if grep -q "rc=0" gpurun_out/v_oz.log; then timeout -s KILL 100 python tools/profile_ozaki.py 16384 1024 4096 7 2 1 >> gpurun_out/v_ab.log 2>&1; fi
cat gpurun_out/v_ab.log
if grep -q "rc=0" gpurun_out/v_oz.log; then
timeout -s KILL 600 python -m pytest tests -x -q -m gpu > gpurun_out/v_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_tests.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/v_smoke.log
tail -3 gpurun_out/v_tests.log; tail -2 gpurun_out/v_smoke.log
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 2 --budget-s 110 --cpu-budget-s 2 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "rc=$?" >> gpurun_out/v_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/v_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'var',round(d['phases_ms']['var'],1),'kernel s',round(d['roofline']['kernel_seconds_per_step'],3),'TOPS',round(d['roofline']['achieved']),'frac',round(d['roofline']['frac'],3),'clocks',d['clocks']['sm_mhz'],'steps',d['steps'])
PY
fi
