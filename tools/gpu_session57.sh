#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_s57_n2.json 2> gpurun_out/bench_s57_n2.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s57_n2.json").read().strip().splitlines()[-1])
print("n2 value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), j["per_rank"], j["clocks"])
PY
echo "== bench N=1"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s57_n1.json 2> gpurun_out/bench_s57_n1.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s57_n1.json").read().strip().splitlines()[-1])
print("n1 value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],3), j["clocks"], j["cpu_baseline"]["value"])
PY
