#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests (parity)"; timeout 2400 python -m pytest tests/test_parity.py tests/test_golden.py tests/test_reference_live.py -m gpu -x -q 2>&1 | tail -4
for w in c2 c1 c4 c3; do timeout 1200 python tools/sweep.py --workload $w --pooled 0,1 --reps 3 2>&1 | tail -2 | tee -a gpurun_out/sweep_s13.log; done
timeout 900 python tools/sweep.py --workload c3 --pool 33554432 --reps 2 2>&1 | tail -1 | tee -a gpurun_out/sweep_s13.log
timeout 900 python tools/sweep.py --workload c2 --refill 16,20,28 --reps 2 2>&1 | tail -3 | tee -a gpurun_out/sweep_s13.log
timeout 900 python tools/sweep.py --workload c2 --pool 4194304,16777216 --reps 2 2>&1 | tail -2 | tee -a gpurun_out/sweep_s13.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s13.json 2> gpurun_out/bench_s13.err; tail -c 3000 gpurun_out/bench_s13.json; tail -3 gpurun_out/bench_s13.err
