#!/bin/bash
# A/B: second triangle test of the pair behind a branch (skipped when no lane of the warp has a second triangle)
mkdir -p gpurun_out
for w in c3 c2 c4; do
for lib in librtb.so librtb_pb.so librtb.so librtb_pb.so; do
echo "== $w $lib x"; RTB_LIB=$PWD/rtcuda_b200/$lib timeout 600 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c1-200
done; done 2>&1 | tee gpurun_out/sweep_s72.log
