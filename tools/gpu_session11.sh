#!/bin/bash
# v2 traversal kernels (pooled triangle tests, packed meta decode, fast 1/d, fused extend+shadow launch)
mkdir -p gpurun_out
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== c2 variants"
timeout 900 python tools/sweep.py --workload c2 --ve 0,4 --vs 0,4 --fused 0 --prefetch 0 --reps 3 2>&1 | tee gpurun_out/sweep_v2_c2.log
timeout 900 python tools/sweep.py --workload c2 --fused 0,1 --prefetch 0,1 --reps 3 2>&1 | tee -a gpurun_out/sweep_v2_c2.log
echo "== c1 c3 c4"
for w in c1 c4 c3; do timeout 1200 python tools/sweep.py --workload $w --ve 0,4 --vs 0,4 --fused 0 --reps 2 2>&1 | tail -2 | tee -a gpurun_out/sweep_v2_other.log; done
for w in c1 c4 c3; do timeout 1200 python tools/sweep.py --workload $w --reps 2 2>&1 | tail -1 | tee -a gpurun_out/sweep_v2_other.log; done
