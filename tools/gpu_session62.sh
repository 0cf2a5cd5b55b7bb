#!/bin/bash
# validation after the two-level BVH: all GPU tests, smoke, bench N=1 (C2), C3i again (SAH-optimal top tree), ncu of the INST kernel
mkdir -p gpurun_out
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench N=1"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s62_n1.json 2> gpurun_out/bench_s62_n1.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s62_n1.json").read().strip().splitlines()[-1])
print("n1 value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],3), j["clocks"], j["cpu_baseline"]["value"])
PY
echo "== bench c3i"
timeout 600 python bench.py --workload c3i --steps 3 --warmup 3 > gpurun_out/bench_s62_c3i.json 2> gpurun_out/bench_s62_c3i.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s62_c3i.json").read().strip().splitlines()[-1])
r=j["roofline"]
print("c3i value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e ms", round(j["e2e"]["ms_per_step"],1), "nodes", j["bvh_nodes"], "e", r["extend_nodes_per_ray"], r["extend_tris_per_ray"], "s", r["shadow_nodes_per_ray"], r["shadow_tris_per_ray"])
PY
echo "== ncu full c3is"
timeout 600 python tools/profile_run.py --workload c3is --reps 1 > gpurun_out/plain_s62_c3is.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace' -s 1 -c 2 -o gpurun_out/prof_s62_c3is python tools/profile_run.py --workload c3is --reps 1 > gpurun_out/ncu_s62_f.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/plain_s62_c3is.log
