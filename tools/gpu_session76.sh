#!/bin/bash
# A/B: how many 128-thread trace blocks per SM beside 128-thread shade blocks
mkdir -p gpurun_out
for w in c2 c4 c3 c1; do
for cfg in "librtb.so 4" "librtb_tb128_sb128.so 3" "librtb_tb128_sb128.so 4" "librtb_tb128_sb128.so 5" "librtb_tb128_sb128.so 6" "librtb_tb128_sb128.so 8" "librtb_tb128_sb64.so 5" "librtb_tb128_sb128.so 5"; do
set -- $cfg
echo "== $w $1 trace_blocks=$2"; RTB_TRACE_BLOCKS=$2 RTB_LIB=$PWD/rtcuda_b200/$1 timeout 600 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c1-200
done; done 2>&1 | tee gpurun_out/sweep_s76.log
