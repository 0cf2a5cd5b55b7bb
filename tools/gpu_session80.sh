#!/bin/bash
# per-kernel evidence with ONE wavefront (full-occupancy launches, the configuration bench.py's roofline times): launch list + ncu full
mkdir -p gpurun_out
export RTB_PIPELINES=1
timeout 600 python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/plain_s80_c2m.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_s80_c2m.csv python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s80_l.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/plain_s80_c2m.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_trace' -s 2 -c 4 -o gpurun_out/prof_s80_c2m python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s80_f.log 2>&1
echo "rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:'k_trace|k_shade' -c 200 --csv --log-file gpurun_out/dram_s80_c2m.csv python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s80_d.log 2>&1
echo "rc=$?"
