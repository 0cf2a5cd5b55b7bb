#!/bin/bash
# two-level BVH: GPU tests of the instanced path, C3 as instances (bench), C3 flat beside it
mkdir -p gpurun_out
echo "== instancing gpu tests"; timeout 900 python -m pytest tests/test_instancing.py -m gpu -x -q 2>&1 | tail -15
echo "== bench c3i"
timeout 600 python bench.py --workload c3i --steps 3 --warmup 3 > gpurun_out/bench_s61_c3i.json 2> gpurun_out/bench_s61_c3i.err; tail -3 gpurun_out/bench_s61_c3i.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s61_c3i.json").read().strip().splitlines()[-1])
r=j["roofline"]
print("c3i value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), "e2e ms", round(j["e2e"]["ms_per_step"],1), "build", j["bvh_build_ms"], "nodes", j["bvh_nodes"])
print("nodes/tris per extend", r["extend_nodes_per_ray"], r["extend_tris_per_ray"], "shadow", r["shadow_nodes_per_ray"], r["shadow_tris_per_ray"], "trace share", r["kernel_share_of_step"], "single", r["single_pipeline_ms_per_step"])
PY
echo "== bench c3 flat"
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s61_c3.json 2> gpurun_out/bench_s61_c3.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s61_c3.json").read().strip().splitlines()[-1])
print("c3 value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), "e2e ms", round(j["e2e"]["ms_per_step"],1))
PY
