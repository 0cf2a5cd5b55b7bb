#!/bin/bash
# paired triangle tests as the default (flat + INST kernels): all GPU tests, C2 bench line, C3 / C3i bench lines
mkdir -p gpurun_out
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench N=1"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s65_n1.json 2> gpurun_out/bench_s65_n1.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s65_n1.json").read().strip().splitlines()[-1])
print("n1 value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],3), j["clocks"], j["cpu_baseline"]["value"])
PY
for w in c3 c3i; do
timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s65_$w.json 2> gpurun_out/bench_s65_$w.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s65_$w.json").read().strip().splitlines()[-1])
r=j["roofline"]
print("$w value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e ms", round(j["e2e"]["ms_per_step"],1), "frac", round(r["frac"],3))
PY
done
