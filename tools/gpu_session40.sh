#!/bin/bash
for c in none 25 28 35 50; do echo -n "c2 carveout $c: "; if [ $c = none ]; then unset RTB_CARVEOUT; else export RTB_CARVEOUT=$c; fi; timeout 900 python tools/sweep.py --workload c2 --pipelines 1 --reps 3 2>&1 | tail -1 | cut -c60-170; done
for c in none 25 50; do echo -n "c3 carveout $c: "; if [ $c = none ]; then unset RTB_CARVEOUT; else export RTB_CARVEOUT=$c; fi; timeout 900 python tools/sweep.py --workload c3 --pipelines 1 --reps 3 2>&1 | tail -1 | cut -c60-170; done
