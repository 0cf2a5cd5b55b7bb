#!/bin/bash
# bench.py with the one-thread CPU baseline added
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_s84_n1.json 2> gpurun_out/bench_s84_n1.err; echo "rc=$?"; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s84_n1.json").read().strip().splitlines()[-1])
print("n1 value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],3), "pipes", j["pipelines"], j["clocks"], j["cpu_baseline"])
PY
