#!/bin/bash
# evidence for profiles/: launch list, ncu --set full, DRAM bytes per launch, bench (1 and 2 GPUs come separately)
mkdir -p gpurun_out
echo "== bench N=1"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s17.json 2> gpurun_out/bench_s17.err; tail -c 2500 gpurun_out/bench_s17.json; tail -3 gpurun_out/bench_s17.err
echo "== launch list (c2m, single pipeline so that launches are serial anyway under ncu)"
timeout 600 python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/plain_s17_c2m.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_s17_c2m.csv python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s17_l.log 2>&1
echo "rc=$?"; cat gpurun_out/plain_s17_c2m.log
echo "== dram bytes of every k_trace / k_shade launch (c2m)"
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:'k_trace|k_shade' -c 200 --csv --log-file gpurun_out/dram_s17_c2m.csv python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s17_d.log 2>&1
echo "rc=$?"
echo "== ncu full: c2m big launches (shade, trace) and c3s"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_trace' -s 2 -c 4 -o gpurun_out/prof_s17_c2m python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s17_f.log 2>&1
echo "rc=$?"
timeout 600 python tools/profile_run.py --workload c3s --reps 1 > gpurun_out/plain_s17_c3s.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_trace' -s 0 -c 3 -o gpurun_out/prof_s17_c3s python tools/profile_run.py --workload c3s --reps 1 > gpurun_out/ncu_s17_g.log 2>&1
echo "rc=$?"; cat gpurun_out/plain_s17_c3s.log
