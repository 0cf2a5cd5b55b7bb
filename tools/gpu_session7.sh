#!/bin/bash
mkdir -p gpurun_out
echo "== pool sweep c2"; timeout 900 python tools/sweep.py --workload c2 --pool 8388608,16777216,33554432 --reps 3 2>&1 | tee gpurun_out/sweep_pool_c2.log
echo "== workloads at defaults"
for w in c1 c3 c4; do timeout 1200 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | tee gpurun_out/work_$w.log; done
timeout 1200 python tools/sweep.py --workload c3 --pool 16777216,33554432 --reps 2 2>&1 | tail -2 | tee gpurun_out/work_c3_pool.log
echo "== bench ours"; timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?"; cat gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
