#!/bin/bash
echo "== parity"; timeout 2400 python -m pytest tests/test_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
for w in c2 c4 c3; do echo -n "$w 1 pipe: "; timeout 900 python tools/sweep.py --workload $w --pipelines 1 --reps 3 2>&1 | tail -1 | cut -c60-190; done
for w in c1 c2 c4 c3; do echo -n "$w 2 pipes: "; timeout 900 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c60-190; done
