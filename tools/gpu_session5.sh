#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (parity only)"; timeout 900 python -m pytest tests/test_parity.py -m gpu -x -q 2>&1 | tail -2
echo "== variants c2: ve x vs"
timeout 1200 python tools/sweep.py --workload c2 --refill 24 --steps 1 --chunk 128 --ve 0,1,2,3 --vs 0,1,2,3 2>&1 | tee gpurun_out/sweep_variants_c2.log
echo "== variants c3s"
timeout 1200 python tools/sweep.py --workload c3s --refill 24 --steps 1 --chunk 128 --ve 0,1,2,3 --vs 0 2>&1 | tee gpurun_out/sweep_variants_c3s.log
echo "== variants c4s"
timeout 1200 python tools/sweep.py --workload c4s --refill 24 --steps 1 --chunk 128 --ve 0,2 --vs 0,2 2>&1 | tee gpurun_out/sweep_variants_c4s.log
