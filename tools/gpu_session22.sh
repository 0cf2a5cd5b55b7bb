#!/bin/bash
mkdir -p gpurun_out
for fs in 0 16 112; do timeout 900 python tools/sweep.py --workload c3 --first-sample $fs --reps 3 2>&1 | tail -1; done
timeout 900 python tools/sweep.py --workload c3 --first-sample 0 --device 1 --reps 3 2>&1 | tail -1
timeout 900 python tools/sweep.py --workload c3 --first-sample 16 --device 1 --reps 3 2>&1 | tail -1
