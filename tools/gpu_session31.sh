#!/bin/bash
echo "== parity"; timeout 2400 python -m pytest tests/test_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
for w in c3 c2; do for l2 in 0 1; do echo -n "$w l2persist $l2: "; RTB_L2_PERSIST=$l2 timeout 900 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c75-200; done; done
