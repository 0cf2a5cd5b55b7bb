#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== perf sanity"; for w in c2 c3; do timeout 900 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c1-170; done
echo "== builder quality"; timeout 600 python tools/builder_quality.py --scene s1 2>&1 | tee gpurun_out/builder_quality_s1.log
timeout 1500 python tools/builder_quality.py --scene s2 --grid 12 --rays 200000 2>&1 | tee gpurun_out/builder_quality_s2.log
