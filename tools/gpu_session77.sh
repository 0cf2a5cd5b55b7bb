#!/bin/bash
# half-occupancy trace launches as the default: all GPU tests, then 2 / 3 / 4 wavefronts
mkdir -p gpurun_out
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c2 c4 c1 c3; do
for p in 2 3 4 2; do
echo "== $w pipelines=$p"; timeout 600 python tools/sweep.py --workload $w --reps 3 --pipelines $p 2>&1 | tail -1 | cut -c1-200
done; done 2>&1 | tee gpurun_out/sweep_s77.log
