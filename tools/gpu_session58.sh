#!/bin/bash
for rep in 1 2; do for w in c2 c3; do for pf in 1 3; do echo -n "$w prefetch $pf: "; timeout 900 python tools/sweep.py --workload $w --prefetch $pf --reps 3 2>&1 | tail -1 | cut -c60-170; done; done; done
