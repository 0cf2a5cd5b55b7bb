#!/bin/bash
mkdir -p gpurun_out
echo "== scale tests"; timeout 1700 python -m pytest tests/test_scale.py -m gpu -x -q 2>&1 | tail -8
echo "== shade occupancy 3 vs 4 (c2)"
RTB_SHADE_OCC=3 timeout 600 python tools/sweep.py --workload c2 --reps 3 | tail -1
RTB_SHADE_OCC=4 timeout 600 python tools/sweep.py --workload c2 --reps 3 | tail -1
echo "== ncu full k_extend/k_shadow on the 10M scene (c3s)"
timeout 900 python tools/profile_run.py --workload c3s --reps 1 > gpurun_out/plain_c3s.log 2>&1 && \
timeout 1700 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow' -s 1 -c 4 -o gpurun_out/prof_c3s python tools/profile_run.py --workload c3s --reps 1 > gpurun_out/ncu_c3s.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_c3s.log; cat gpurun_out/plain_c3s.log
