#!/bin/bash
# 4 vs 6 vs 8 wavefronts (trace launches of 8 / W blocks of 128 threads per SM)
mkdir -p gpurun_out
for w in c2 c4 c1; do
for p in 4 8 6 4 8; do
echo "== $w pipelines=$p"; timeout 600 python tools/sweep.py --workload $w --reps 3 --pipelines $p 2>&1 | tail -1 | cut -c1-200
done; done 2>&1 | tee gpurun_out/sweep_s82.log
