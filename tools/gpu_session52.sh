#!/bin/bash
for rep in 1 2; do for w in c4 c2; do timeout 900 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c60-170; done; done
