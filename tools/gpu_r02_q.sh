#!/bin/bash
# round-2 GPU check Q: is the emulated GEMM latency- or bandwidth-bound?  Same kernel with 6 (shipped), 4 and 3 pipeline stages
mkdir -p gpurun_out
cp linpde_gp_b200/lib/liblpgp.so /tmp/liblpgp_6.so
echo "stages 6" > gpurun_out/q_stages.log; timeout -s KILL 120 python tools/profile_ozaki.py 32768 1024 32768 7 >> gpurun_out/q_stages.log 2>&1
for st in 4 3; do cp tools/_exp/liblpgp_st$st.so linpde_gp_b200/lib/liblpgp.so; echo "stages $st" >> gpurun_out/q_stages.log; timeout -s KILL 120 python tools/profile_ozaki.py 32768 1024 32768 7 >> gpurun_out/q_stages.log 2>&1; done
cp /tmp/liblpgp_6.so linpde_gp_b200/lib/liblpgp.so
cat gpurun_out/q_stages.log
