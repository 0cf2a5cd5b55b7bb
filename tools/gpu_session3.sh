#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== refill sweep (c2)"
for r in 1 8 16 20 24 28 32; do echo -n "refill $r: "; RTB_REFILL=$r timeout 600 python tools/profile_run.py --workload c2 --reps 2 | tail -1; done
echo "== pool sweep (c2, default refill)"
for p in 2097152 4194304 8388608 16777216; do echo -n "pool $p: "; RTB_POOL=$p timeout 600 python tools/profile_run.py --workload c2 --reps 2 | tail -1; done
echo "== c1 / c4s / flat"
timeout 600 python tools/profile_run.py --workload c1 --reps 2 | tail -1
timeout 600 python tools/profile_run.py --workload c4s --reps 2 | tail -1
timeout 600 python tools/profile_run.py --workload c2 --reps 2 --flags 4 | tail -1
echo "== bench ours"; timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?"; cat gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
echo "== ncu full (big launches)"
timeout 600 python tools/profile_run.py --workload c2s --reps 1 > gpurun_out/plain_c2s_b.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_shade' -s 3 -c 6 -o gpurun_out/prof_big_s3 python tools/profile_run.py --workload c2s --reps 1 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full.log
