#!/bin/bash
mkdir -p gpurun_out
P=67108864
timeout 1200 python tools/sweep.py --workload c3 --pool $P --count --reps 2 2>&1 | tail -2 | tee -a gpurun_out/sweep_s15.log
timeout 1200 python tools/sweep.py --workload c3 --pool $P --chunk 32,64 --reps 2 2>&1 | tail -2 | tee -a gpurun_out/sweep_s15.log
timeout 1200 python tools/sweep.py --workload c3 --pool $P --refill 16,28 --reps 2 2>&1 | tail -2 | tee -a gpurun_out/sweep_s15.log
timeout 1200 python tools/sweep.py --workload c3 --pool $P --pooled 0 --reps 2 2>&1 | tail -1 | tee -a gpurun_out/sweep_s15.log
for r in 8 32 64; do timeout 1200 python tools/sweep.py --workload c3 --pool $P --radius $r --count --reps 2 2>&1 | tail -2 | tee -a gpurun_out/sweep_s15.log; done
for l in 1 2; do timeout 1200 python tools/sweep.py --workload c3 --pool $P --leaf $l --count --reps 2 2>&1 | tail -2 | tee -a gpurun_out/sweep_s15.log; done
timeout 1200 python tools/sweep.py --workload c2 --pool $P --pooled 0 --reps 2 2>&1 | tail -1 | tee -a gpurun_out/sweep_s15.log
timeout 1200 python tools/sweep.py --workload c2 --pool 33554432 --pooled 0 --leaf 1 --count --reps 2 2>&1 | tail -2 | tee -a gpurun_out/sweep_s15.log
timeout 1200 python tools/sweep.py --workload c2 --pool 33554432 --pooled 0 --radius 64 --count --reps 2 2>&1 | tail -2 | tee -a gpurun_out/sweep_s15.log
