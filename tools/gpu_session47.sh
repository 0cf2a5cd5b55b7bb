#!/bin/bash
for rep in 1 2; do
echo -n "c2 prev 1pipe: "; RTB_LIB=$PWD/tools/_exp/librtb_prev.so timeout 900 python tools/sweep.py --workload c2 --pipelines 1 --reps 3 2>&1 | tail -1 | cut -c60-170
echo -n "c2 now  1pipe: "; timeout 900 python tools/sweep.py --workload c2 --pipelines 1 --reps 3 2>&1 | tail -1 | cut -c60-170
done
