#!/bin/bash
for rep in 1 2; do for m in 3 2; do echo -n "c2 shade minb $m 1pipe: "; RTB_SHADE_MINB=$m timeout 900 python tools/sweep.py --workload c2 --pipelines 1 --reps 3 2>&1 | tail -1 | cut -c60-170; done; done
for m in 3 2; do echo -n "c2 shade minb $m 2pipes: "; RTB_SHADE_MINB=$m timeout 900 python tools/sweep.py --workload c2 --reps 3 2>&1 | tail -1 | cut -c60-170; done
