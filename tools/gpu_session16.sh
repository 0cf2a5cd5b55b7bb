#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 1200 python tools/sweep.py --workload c3 --pooled 1 --nodepf 0,1 --reps 3 2>&1 | tail -2 | tee -a gpurun_out/sweep_s16.log
timeout 1200 python tools/sweep.py --workload c3 --pooled 0 --nodepf 1 --reps 2 2>&1 | tail -1 | tee -a gpurun_out/sweep_s16.log
timeout 1200 python tools/sweep.py --workload c2 --pooled 0 --nodepf 0,1 --reps 3 2>&1 | tail -2 | tee -a gpurun_out/sweep_s16.log
for w in c1 c2 c4 c3; do timeout 1200 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | tee -a gpurun_out/sweep_s16.log; done
