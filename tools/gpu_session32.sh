#!/bin/bash
# final evidence of the round for profiles/: tests, bench (ours + reference arm), launch list, ncu --set full
mkdir -p gpurun_out
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== bench N=1"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s32.json 2> gpurun_out/bench_s32.err; tail -c 3300 gpurun_out/bench_s32.json; tail -3 gpurun_out/bench_s32.err
echo "== bench reference"; timeout 1500 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_s32_ref.json 2> gpurun_out/bench_s32_ref.err; tail -c 900 gpurun_out/bench_s32_ref.json
echo "== launch list c2m"
timeout 600 python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/plain_s32_c2m.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_s32_c2m.csv python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s32_l.log 2>&1
echo "rc=$?"; cat gpurun_out/plain_s32_c2m.log
echo "== ncu full c2m"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_trace' -s 2 -c 4 -o gpurun_out/prof_s32_c2m python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s32_f.log 2>&1
echo "rc=$?"
echo "== dram c2m"
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:'k_trace|k_shade' -c 200 --csv --log-file gpurun_out/dram_s32_c2m.csv python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s32_d.log 2>&1
echo "rc=$?"
