#!/bin/bash
# round-end rehearsal: the driver's own command lines (N=2 under torchrun, both arms)
mkdir -p gpurun_out
echo "== reference arm N=1"
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_s69_ref_n1.json 2> gpurun_out/bench_s69_ref_n1.err; tail -c 900 gpurun_out/bench_s69_ref_n1.json; echo
echo "== ours N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_s69_n2.json 2> gpurun_out/bench_s69_n2.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s69_n2.json").read().strip().splitlines()[-1])
print("n2 value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), j["per_rank"], j["clocks"])
PY
echo "== reference arm N=2 (rank 0 only)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_s69_ref_n2.json 2> gpurun_out/bench_s69_ref_n2.err; echo "rc=$?"; tail -c 400 gpurun_out/bench_s69_ref_n2.json; echo
