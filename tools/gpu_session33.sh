#!/bin/bash
echo "== parity with the PRMT build"; RTB_LIB=$PWD/tools/_exp/librtb_prmt.so timeout 2400 python -m pytest tests/test_parity.py tests/test_golden.py tests/test_scale.py -m gpu -x -q 2>&1 | tail -3
for w in c2 c3 c4 c1; do
echo -n "$w i2f : "; timeout 900 python tools/sweep.py --workload $w --reps 3 --count 2>&1 | tail -2 | cut -c60-230 | tr '\n' ' '; echo
echo -n "$w prmt: "; RTB_LIB=$PWD/tools/_exp/librtb_prmt.so timeout 900 python tools/sweep.py --workload $w --reps 3 --count 2>&1 | tail -2 | cut -c60-230 | tr '\n' ' '; echo
done
