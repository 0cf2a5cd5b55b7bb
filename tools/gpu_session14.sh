#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests (parity)"; timeout 2400 python -m pytest tests/test_parity.py tests/test_golden.py tests/test_reference_live.py tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -4
for w in c2 c1 c4; do timeout 1200 python tools/sweep.py --workload $w --pooled 0,1 --pipelines 1,2 --reps 3 2>&1 | tail -4 | tee -a gpurun_out/sweep_s14.log; done
timeout 1200 python tools/sweep.py --workload c2 --pooled 0 --pipelines 1,2 --pool 16777216,33554432 --reps 3 2>&1 | tail -4 | tee -a gpurun_out/sweep_s14.log
timeout 1200 python tools/sweep.py --workload c3 --pooled 1 --pipelines 1,2 --pool 8388608,33554432,67108864 --reps 2 2>&1 | tail -6 | tee -a gpurun_out/sweep_s14.log
