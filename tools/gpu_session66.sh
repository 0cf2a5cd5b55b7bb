#!/bin/bash
# evidence refresh for the final kernels of this session (paired triangle tests): launch list, ncu full c2m / c3s, DRAM + L2 bytes
mkdir -p gpurun_out
echo "== launch list c2m"
timeout 600 python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/plain_s66_c2m.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_s66_c2m.csv python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s66_l.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/plain_s66_c2m.log
echo "== ncu full c2m"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_trace' -s 2 -c 4 -o gpurun_out/prof_s66_c2m python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s66_f.log 2>&1
echo "rc=$?"
echo "== dram c2m"
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:'k_trace|k_shade' -c 200 --csv --log-file gpurun_out/dram_s66_c2m.csv python tools/profile_run.py --workload c2m --reps 1 > gpurun_out/ncu_s66_d.log 2>&1
echo "rc=$?"
echo "== ncu full c3s"
timeout 600 python tools/profile_run.py --workload c3s --reps 1 > gpurun_out/plain_s66_c3s.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_trace' -s 1 -c 2 -o gpurun_out/prof_s66_c3s python tools/profile_run.py --workload c3s --reps 1 > gpurun_out/ncu_s66_g.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/plain_s66_c3s.log
