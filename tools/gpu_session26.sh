#!/bin/bash
for w in c3 c2; do
timeout 1200 python tools/sweep.py --workload $w --pipelines 1,2 --pool 8388608,16777216,33554432,67108864 --reps 2 2>&1 | tail -8 | cut -c1-175 | tee -a gpurun_out/sweep_s26.log
done
timeout 1200 python tools/sweep.py --workload c3 --refill 16,20,28 --reps 2 2>&1 | tail -3 | cut -c1-175 | tee -a gpurun_out/sweep_s26.log
timeout 1200 python tools/sweep.py --workload c3 --chunk 32,64,256 --reps 2 2>&1 | tail -3 | cut -c1-175 | tee -a gpurun_out/sweep_s26.log
