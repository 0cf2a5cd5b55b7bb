#!/bin/bash
# A/B: block size of the persistent trace kernels (256 default; 128, 64: same 32 warps per SM, finer hand-over of SM slots)
mkdir -p gpurun_out
for w in c2 c3 c4; do
for lib in librtb.so librtb_tb128.so librtb_tb64.so librtb.so librtb_tb128.so librtb_tb64.so; do
echo "== $w $lib"; RTB_LIB=$PWD/rtcuda_b200/$lib timeout 600 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c1-200
done; done 2>&1 | tee gpurun_out/sweep_s70.log
