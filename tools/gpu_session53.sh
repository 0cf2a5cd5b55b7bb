#!/bin/bash
for w in c2 c3 c4; do timeout 900 python tools/sweep.py --workload $w --pipelines 1,2,3,4 --reps 3 2>&1 | tail -4 | cut -c1-150; done
timeout 900 python tools/sweep.py --workload c2 --pipelines 3,4 --pool 67108864 --reps 3 2>&1 | tail -2 | cut -c1-150
