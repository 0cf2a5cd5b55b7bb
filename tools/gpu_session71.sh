#!/bin/bash
# A/B: triangle phase only when at least RTB_TRI_VOTE lanes hold triangles (0 = every step)
mkdir -p gpurun_out
for w in c3 c2 c4; do
for v in 0 4 8 12 16 0; do
echo "== $w vote $v"; RTB_TRI_VOTE=$v timeout 600 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c1-200
done; done 2>&1 | tee gpurun_out/sweep_s71.log
echo "== parity with vote 8"; RTB_TRI_VOTE=8 timeout 900 python -m pytest tests/test_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
