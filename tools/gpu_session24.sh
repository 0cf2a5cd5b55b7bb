#!/bin/bash
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for fs in 0 8 16 96; do echo -n "fs $fs: "; timeout 900 python tools/sweep.py --workload c3 --first-sample $fs --reps 2 --count 2>&1 | tail -2 | cut -c60-200; done
timeout 900 python tools/sweep.py --workload c3 --pipelines 1,2 --pooled 0,1 --reps 2 2>&1 | tail -4 | cut -c1-150
for w in c1 c2 c4; do timeout 900 python tools/sweep.py --workload $w --reps 3 --count 2>&1 | tail -2 | cut -c1-200; done
