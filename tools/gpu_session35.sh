#!/bin/bash
for w in c2 c3; do for ts in 2 3 4; do echo -n "$w own tri_step $ts: "; RTB_TRI_STEP=$ts timeout 900 python tools/sweep.py --workload $w --pooled 0 --reps 3 2>&1 | tail -1 | cut -c60-170; done; done
for w in c2 c3; do for rf in 20 28; do echo -n "$w own tri_step 2 refill $rf: "; RTB_TRI_STEP=2 timeout 900 python tools/sweep.py --workload $w --pooled 0 --refill $rf --reps 3 2>&1 | tail -1 | cut -c60-170; done; done
