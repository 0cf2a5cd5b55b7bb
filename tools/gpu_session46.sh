#!/bin/bash
for st in 0 4096 2101248 35655680; do echo -n "c2 stagger $st 1pipe: "; RTB_STAGGER=$st timeout 900 python tools/sweep.py --workload c2 --pipelines 1 --reps 3 2>&1 | tail -1 | cut -c60-170; done
for st in 0 2101248; do echo -n "c2 stagger $st 2pipes: "; RTB_STAGGER=$st timeout 900 python tools/sweep.py --workload c2 --reps 3 2>&1 | tail -1 | cut -c60-170; done
for st in 0 2101248; do echo -n "c3 stagger $st 2pipes: "; RTB_STAGGER=$st timeout 900 python tools/sweep.py --workload c3 --reps 2 2>&1 | tail -1 | cut -c60-170; done
echo -n "c2 prev 1pipe: "; RTB_LIB=$PWD/tools/_exp/librtb_prev.so timeout 900 python tools/sweep.py --workload c2 --pipelines 1 --reps 3 2>&1 | tail -1 | cut -c60-170
