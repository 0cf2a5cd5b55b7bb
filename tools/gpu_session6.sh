#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== workloads"
for w in c1 c2 c3 c4; do timeout 1200 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | tee gpurun_out/work_$w.log; done
echo "== bench reference"; timeout 1500 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json
echo "== bench ours"; timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?"; cat gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
echo "== ncu launch list (c2s)"
timeout 600 python tools/profile_run.py --workload c2s --reps 1 > gpurun_out/plain_c2s.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_c2s.csv python tools/profile_run.py --workload c2s --reps 1 > gpurun_out/ncu_launches.log 2>&1
echo "rc=$?"
echo "== ncu dram bytes of every k_extend launch of one C2 render"
timeout 600 python tools/profile_run.py --workload c2 --reps 1 > gpurun_out/plain_c2.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:'k_extend|k_shadow' --csv --log-file gpurun_out/dram_c2.csv python tools/profile_run.py --workload c2 --reps 1 > gpurun_out/ncu_dram.log 2>&1
echo "rc=$?"
echo "== ncu full (big launches)"
timeout 600 python tools/profile_run.py --workload c2s --reps 1 > gpurun_out/plain_c2s_b.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_shade' -s 3 -c 6 -o gpurun_out/prof_big_s6 python tools/profile_run.py --workload c2s --reps 1 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full.log
