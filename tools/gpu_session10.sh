#!/bin/bash
mkdir -p gpurun_out
echo "== parity + scale tests"; timeout 2400 python -m pytest tests/test_parity.py tests/test_golden.py tests/test_scale.py -m gpu -x -q 2>&1 | tail -8
echo "== refill sweep c2"
timeout 900 python tools/sweep.py --workload c2 --refill 8,16,20,24,28,32 --reps 3 2>&1 | tee gpurun_out/sweep_refill2_c2.log
echo "== c1 c3 c4"
for w in c1 c3 c4; do timeout 1200 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1; done
timeout 900 python tools/sweep.py --workload c3 --refill 8,16,28 --reps 2 2>&1 | tail -3
