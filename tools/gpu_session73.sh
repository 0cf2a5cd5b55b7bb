#!/bin/bash
# static triangles held by the top tree: instancing tests, C3i bench line
mkdir -p gpurun_out
echo "== instancing + multi gpu tests"; timeout 900 python -m pytest tests/test_instancing.py tests/test_scale.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --workload c3i --steps 3 --warmup 3 > gpurun_out/bench_s73_c3i.json 2> gpurun_out/bench_s73_c3i.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s73_c3i.json").read().strip().splitlines()[-1])
r=j["roofline"]
print("c3i value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e ms", round(j["e2e"]["ms_per_step"],1), "nodes", j["bvh_nodes"], "e", round(r["extend_nodes_per_ray"],2), round(r["extend_tris_per_ray"],2), "s", round(r["shadow_nodes_per_ray"],2), round(r["shadow_tris_per_ray"],2))
PY
