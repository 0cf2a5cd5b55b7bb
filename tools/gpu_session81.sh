#!/bin/bash
# C3 as instances: wavefronts / trace blocks per SM
mkdir -p gpurun_out
for cfg in "4 0" "2 0" "2 8" "4 3"; do set -- $cfg
echo "== c3is pipelines=$1 trace_blocks=$2"; RTB_PIPELINES=$1 RTB_TRACE_BLOCKS=$2 timeout 300 python tools/profile_run.py --workload c3is --reps 3 2>&1 | tail -2
done 2>&1 | tee gpurun_out/sweep_s81.log
