#!/bin/bash
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c1 c2 c4 c3; do timeout 900 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c1-170; done
timeout 900 python tools/sweep.py --workload c4 --pipelines 1 --reps 3 2>&1 | tail -1 | cut -c1-170
