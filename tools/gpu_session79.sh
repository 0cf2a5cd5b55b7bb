#!/bin/bash
# N=2 with the new defaults (driver's command line)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_s79_n2.json 2> gpurun_out/bench_s79_n2.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s79_n2.json").read().strip().splitlines()[-1])
print("n2 value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), j["per_rank"], j["clocks"])
PY
