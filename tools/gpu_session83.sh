#!/bin/bash
# final state of the session: all GPU tests, smoke, default bench line
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench N=1"
timeout 600 python bench.py > gpurun_out/bench_s83_n1.json 2> gpurun_out/bench_s83_n1.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s83_n1.json").read().strip().splitlines()[-1])
print("n1 value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],3), "pipes", j["pipelines"], "launches", j["gpu_launches"], j["clocks"], j["cpu_baseline"]["value"])
PY
