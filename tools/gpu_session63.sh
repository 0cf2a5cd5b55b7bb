#!/bin/bash
# A/B: traversal stack in local memory (default) vs first 8 entries in shared memory (RTB_SMEM_STACK=8)
mkdir -p gpurun_out
for w in c2 c3 c4; do
for s in 0 8 0 8; do
echo "== $w RTB_SMEM_STACK=$s"; RTB_SMEM_STACK=$s timeout 600 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1
done; done 2>&1 | tee gpurun_out/sweep_s63.log
echo "== instancing gpu tests"; timeout 900 python -m pytest tests/test_instancing.py -m gpu -x -q 2>&1 | tail -3
