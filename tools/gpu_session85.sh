#!/bin/bash
# bench.py with the SURVEY 8d floor-model fraction in `roofline`
mkdir -p gpurun_out
timeout 250 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s85.json 2> gpurun_out/bench_s85.err; echo "rc=$?"
python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s85.json").read().strip().splitlines()[-1])
print(round(j["value"],1), j["roofline"]["floor_model"])
PY
