#!/bin/bash
# GPU session: tests, timings, bench, ncu captures of the big launches.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== profile runs"
for w in c1 c2 c4s; do timeout 600 python tools/profile_run.py --workload $w --reps 3 > gpurun_out/prun_$w.log 2>&1; tail -3 gpurun_out/prun_$w.log; done
timeout 600 python tools/profile_run.py --workload c2 --reps 2 --flags 4 > gpurun_out/prun_c2_flat.log 2>&1; tail -2 gpurun_out/prun_c2_flat.log
for pool in 1048576 4194304 8388608; do timeout 600 python tools/profile_run.py --workload c2 --reps 2 --pool $pool > gpurun_out/prun_c2_pool$pool.log 2>&1; echo "pool $pool"; tail -1 gpurun_out/prun_c2_pool$pool.log; done
echo "== bench ours"; timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?"; cat gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
echo "== ncu launches c2s"
timeout 600 python tools/profile_run.py --workload c2s --reps 1 > gpurun_out/plain_c2s.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_c2s.csv python tools/profile_run.py --workload c2s --reps 1 > gpurun_out/ncu_launches.log 2>&1
echo "rc=$?"
echo "== ncu full extend/shadow/shade (big launches)"
timeout 600 python tools/profile_run.py --workload c2s --reps 1 > gpurun_out/plain_c2s_b.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_shade' -s 3 -c 6 -o gpurun_out/prof_big_r1 python tools/profile_run.py --workload c2s --reps 1 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out | head -40
