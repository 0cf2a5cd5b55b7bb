#!/bin/bash
# A/B: two triangle tests of a lane's list issued together (RTB_TRI_PAIR build) vs the sequential loop
mkdir -p gpurun_out
for w in c2 c3 c4; do
for lib in librtb.so librtb_pair.so librtb.so librtb_pair.so; do
echo "== $w $lib"; RTB_LIB=$PWD/rtcuda_b200/$lib timeout 600 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c1-200
done; done 2>&1 | tee gpurun_out/sweep_s64.log
echo "== parity with the pair build"; RTB_LIB=$PWD/rtcuda_b200/librtb_pair.so timeout 900 python -m pytest tests/test_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
