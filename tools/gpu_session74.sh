#!/bin/bash
# A/B: trace kernel capped at 3 blocks per SM so that a 128-thread shade block of the OTHER wavefront fits beside it
mkdir -p gpurun_out
for w in c2 c3; do
for cfg in "librtb.so 4" "librtb_sb128.so 3" "librtb.so 3" "librtb_sb128.so 4" "librtb.so 4" "librtb_sb128.so 3"; do
set -- $cfg
echo "== $w $1 trace_blocks=$2"; RTB_TRACE_BLOCKS=$2 RTB_LIB=$PWD/rtcuda_b200/$1 timeout 600 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c1-200
done; done 2>&1 | tee gpurun_out/sweep_s74.log
