#!/bin/bash
for w in c2 c4 c3; do for sp in 0 1; do echo -n "$w shade_prefetch $sp: "; RTB_SHADE_PREFETCH=$sp timeout 900 python tools/sweep.py --workload $w --pipelines 1 --reps 3 2>&1 | tail -1 | cut -c60-190; done; done
for sp in 0 1; do echo -n "c2 2 pipes shade_prefetch $sp: "; RTB_SHADE_PREFETCH=$sp timeout 900 python tools/sweep.py --workload c2 --reps 3 2>&1 | tail -1 | cut -c60-190; done
