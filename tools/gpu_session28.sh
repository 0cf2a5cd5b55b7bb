#!/bin/bash
# C5 strong scaling (10 M triangles, 4K, 1024 spp in total) at 8/4/2/1 GPUs, and the C3 weak-scaling line at 8
mkdir -p gpurun_out
run() { # N workload tag steps warmup
  if [ "$1" = "1" ]; then timeout 1200 python bench.py --gpus 1 --steps $4 --warmup $5 --workload $2 --no-cpu-baseline > gpurun_out/bench_s28_$3.json 2> gpurun_out/bench_s28_$3.err
  else timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $1 --steps $4 --warmup $5 --workload $2 --no-cpu-baseline > gpurun_out/bench_s28_$3.json 2> gpurun_out/bench_s28_$3.err; fi
  echo "== $3 rc=$?"; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_s28_$3.json").read().strip().splitlines()[-1])
    print("$3", "value", round(j["value"],1), "ms_per_step", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), "n", j["n_gpus"], j.get("per_rank"))
except Exception as e:
    print("$3 failed", e); print(open("gpurun_out/bench_s28_$3.err").read()[-1500:])
PY
}
run 8 c5 c5_n8 2 2
run 8 c3 c3_n8 3 3
run 4 c5 c5_n4 2 1
run 2 c5 c5_n2 2 1
run 1 c5 c5_n1 1 1
