#!/bin/bash
# A/B matrix: trace blocks per SM (cap) x trace block size x shade block size
mkdir -p gpurun_out
for w in c2 c4 c3; do
for cfg in "librtb.so 4" "librtb_sb128.so 3" "librtb_sb64.so 3" "librtb_tb128_sb128.so 6" "librtb_tb128_sb128.so 5" "librtb_tb128_sb64.so 7" "librtb_tb128_sb64.so 6" "librtb_sb128.so 3"; do
set -- $cfg
echo "== $w $1 trace_blocks=$2"; RTB_TRACE_BLOCKS=$2 RTB_LIB=$PWD/rtcuda_b200/$1 timeout 600 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c1-200
done; done 2>&1 | tee gpurun_out/sweep_s75.log
