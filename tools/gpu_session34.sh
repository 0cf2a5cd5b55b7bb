#!/bin/bash
echo "== parity tri_step=1"; RTB_TRI_STEP=1 RTB_POOLED=0 timeout 2400 python -m pytest tests/test_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
for w in c2 c4 c1; do for ts in 0 1 2; do echo -n "$w tri_step $ts: "; RTB_TRI_STEP=$ts timeout 900 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c60-170; done; done
for ts in 0 1 2; do echo -n "c3 own tri_step $ts: "; RTB_TRI_STEP=$ts timeout 900 python tools/sweep.py --workload c3 --pooled 0 --reps 2 2>&1 | tail -1 | cut -c60-170; done
echo -n "c3 pooled: "; timeout 900 python tools/sweep.py --workload c3 --pooled 1 --reps 2 2>&1 | tail -1 | cut -c60-170
