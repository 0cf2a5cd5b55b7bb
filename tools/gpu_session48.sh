#!/bin/bash
for rep in 1 2; do
for l in prev now; do
echo -n "c2 $l 1pipe: "; if [ $l = now ]; then unset RTB_LIB; else export RTB_LIB=$PWD/tools/_exp/librtb_$l.so; fi; timeout 900 python tools/sweep.py --workload c2 --pipelines 1 --reps 3 2>&1 | tail -1 | cut -c60-170
done; done
