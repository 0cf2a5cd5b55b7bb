#!/bin/bash
# 8-GPU weak scaling: C2 (bench default) and C3 (10 M triangles, 4K)
mkdir -p gpurun_out
run() { # N workload tag
  if [ "$1" = "1" ]; then timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --workload $2 --no-cpu-baseline > gpurun_out/bench_s20_$3.json 2> gpurun_out/bench_s20_$3.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $1 --steps 3 --warmup 3 --workload $2 --no-cpu-baseline > gpurun_out/bench_s20_$3.json 2> gpurun_out/bench_s20_$3.err; fi
  echo "== $3 rc=$?"; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_s20_$3.json").read().strip().splitlines()[-1])
    print("$3", "value", round(j["value"],1), "ms_per_step", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), "n", j["n_gpus"])
except Exception as e:
    print("$3 failed", e); print(open("gpurun_out/bench_s20_$3.err").read()[-1500:])
PY
}
run 8 c2 c2_n8
run 8 c3 c3_n8
run 1 c3 c3_n1
run 4 c2 c2_n4
