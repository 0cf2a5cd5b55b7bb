#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (parity only)"; timeout 900 python -m pytest tests/test_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
echo "== sweep c2"
timeout 1200 python tools/sweep.py --workload c2 --refill 8,16,24 --steps 1,2,4,8,16 --chunk 128 2>&1 | tee gpurun_out/sweep_c2_a.log
timeout 600 python tools/sweep.py --workload c2 --refill 16 --steps 4 --chunk 32,64,256,1024 2>&1 | tee gpurun_out/sweep_c2_b.log
echo "== c3s (10M triangles, 4K)"
timeout 1200 python tools/sweep.py --workload c3s --refill 16 --steps 4 --chunk 128 --reps 2 2>&1 | tee gpurun_out/sweep_c3s.log
echo "== c4s"
timeout 600 python tools/sweep.py --workload c4s --refill 16 --steps 4 --chunk 128 2>&1 | tee gpurun_out/sweep_c4s.log
