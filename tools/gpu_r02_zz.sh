#!/bin/bash
# round-2 GPU check ZZ: full -m gpu suite at HEAD (new: matrix-composed functionals, condition_normal_on_observations)
mkdir -p gpurun_out
timeout -s KILL 85 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/zz_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/zz_tests.log
grep -E "passed|failed|FAILED|Error|rc=" gpurun_out/zz_tests.log | head -20
