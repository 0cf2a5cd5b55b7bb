#!/bin/bash
mkdir -p gpurun_out
echo "== parity"; timeout 2400 python -m pytest tests/test_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
for w in c2 c3; do
timeout 1200 python tools/sweep.py --workload $w --prefetch 1,2 --chunk 64,128 --reps 3 2>&1 | tail -4 | tee -a gpurun_out/sweep_s19.log
done
timeout 1200 python tools/sweep.py --workload c4 --prefetch 1,2 --reps 3 2>&1 | tail -2 | tee -a gpurun_out/sweep_s19.log
timeout 1200 python tools/sweep.py --workload c1 --prefetch 1,2 --reps 3 2>&1 | tail -2 | tee -a gpurun_out/sweep_s19.log
