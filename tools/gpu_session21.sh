#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --workload c3 --no-cpu-baseline > gpurun_out/bench_s21_c3_n$N.json 2> gpurun_out/bench_s21_c3_n$N.err
echo rc=$?
python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s21_c3_n$N.json").read().strip().splitlines()[-1])
print("value", round(j["value"],1), "ms_per_step", round(j["ms_per_step"],2), j["per_rank"])
PY
grep -i "NVLS\|P2P\|via\|Channel 00" gpurun_out/bench_s21_c3_n$N.err | head -12
