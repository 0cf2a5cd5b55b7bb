#!/bin/bash
echo "== parity"; timeout 2400 python -m pytest tests/test_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
for rep in 1; do for w in c2 c3 c4; do
echo -n "$w prev 1pipe: "; RTB_LIB=$PWD/tools/_exp/librtb_prev.so timeout 900 python tools/sweep.py --workload $w --pipelines 1 --reps 3 2>&1 | tail -1 | cut -c60-170
echo -n "$w now  1pipe: "; timeout 900 python tools/sweep.py --workload $w --pipelines 1 --reps 3 2>&1 | tail -1 | cut -c60-170
echo -n "$w prev 2pipes: "; RTB_LIB=$PWD/tools/_exp/librtb_prev.so timeout 900 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c60-170
echo -n "$w now  2pipes: "; timeout 900 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c60-170
done; done
