#!/bin/bash
# Alternating A/B of library builds on one box: tools/sweep.py (best of 3 renders) for every workload, twice over all
# builds.  usage: tools/ab.sh out.log "c2 c4 c3" build_a.so build_b.so [...]
# (a variant build: nvcc ... -D<MACRO> -shared rtcuda_b200/csrc/rtb_cuda.cu ... -o tools/_exp/librtb_x.so; RTB_LIB selects it)
out=$1; wl=$2; shift 2
for rep in 1 2; do
  for lib in "$@"; do
    for w in $wl; do
      echo "== $lib $w" >> "$out"
      RTB_LIB=$lib python tools/sweep.py --workload $w --pipelines 0 --reps 3 2>&1 | grep pipes | cut -c1-170 >> "$out"
    done
  done
done
