#!/bin/bash
# round-2 GPU check U: ncu --set full of the clustered emulated GEMM (one launch), ncu launch list of bench.py at a reduced size at HEAD
mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:ozaki_gemm -s 1 -c 1 -f -o gpurun_out/prof_ozaki_r02u python tools/profile_ozaki.py 16384 1024 16384 7 2 > gpurun_out/u_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02u.csv python bench.py --npde 15872 --nbc-edge 128 --grid 256 --steps 1 --warmup 1 --cpu-budget-s 2 > gpurun_out/u_launch_bench.json 2> gpurun_out/u_launch.err
tail -4 gpurun_out/u_ncu.log; python tools/launch_summary.py gpurun_out/launches_r02u.csv | head -12
