#!/bin/bash
# hole-based ray queues (no atomics in shade), pooled/own triangle tests, fused launch
mkdir -p gpurun_out
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== c2"
timeout 900 python tools/sweep.py --workload c2 --pooled 0,1 --fused 0,1 --reps 3 2>&1 | tee gpurun_out/sweep_s12_c2.log
timeout 900 python tools/sweep.py --workload c2 --pooled 0 --shade-occ 3,4 --prefetch 0,1 --reps 3 2>&1 | tee -a gpurun_out/sweep_s12_c2.log
echo "== c1 c4 c3"
for w in c1 c4 c3; do timeout 1200 python tools/sweep.py --workload $w --pooled 0,1 --reps 2 2>&1 | tail -2 | tee -a gpurun_out/sweep_s12_other.log; done
echo "== ncu full (c2s): shade + trace"
timeout 600 python tools/profile_run.py --workload c2s --reps 1 > gpurun_out/plain_s12.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_trace' -s 2 -c 4 -o gpurun_out/prof_s12 python tools/profile_run.py --workload c2s --reps 1 > gpurun_out/ncu_s12.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_s12.log; cat gpurun_out/plain_s12.log
