#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_s18_n2.json 2> gpurun_out/bench_s18_n2.err; tail -c 1500 gpurun_out/bench_s18_n2.json; tail -3 gpurun_out/bench_s18_n2.err
echo "== bench N=1"
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s18_n1.json 2> gpurun_out/bench_s18_n1.err; tail -c 600 gpurun_out/bench_s18_n1.json
echo "== bench reference"
timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_s18_ref.json 2> gpurun_out/bench_s18_ref.err; tail -c 1200 gpurun_out/bench_s18_ref.json
