#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/plain_s42.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_trace' -s 1 -c 2 -o gpurun_out/prof_s42_c2m python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s42.log 2>&1
echo "rc=$?"; cat gpurun_out/plain_s42.log
