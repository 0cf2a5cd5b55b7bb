#!/bin/bash
# single-launch PLOC tail: tests that exercise the builder, build times with / without it, C2 bench line (e2e includes the build)
mkdir -p gpurun_out
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for t in 0 1 0 1; do
echo "== RTB_PLOC_TAIL=$t"; RTB_PLOC_TAIL=$t python - <<PY
import sys, time; sys.path.insert(0, ".")
from rtcuda_b200 import capi
L = capi.Lib(); ctx = L.context(0); v, f = L.load_mesh()
for kind, grid, inst in ((1, 0, False), (3, 12, True), (3, 4, False)):
    hs = L.host_scene_instanced(kind, v, f, grid=grid) if inst else L.host_scene(kind, v, f, grid=grid)
    best = 1e9; wall = 1e9
    for _ in range(6):
        t0 = time.perf_counter(); sc = ctx.scene(hs.idesc if inst else hs.desc); wall = min(wall, (time.perf_counter() - t0) * 1e3)
        st = sc.stats(); best = min(best, st.build_ms); sc.close()
    print(f"kind {kind} grid {grid} instanced {inst}: {st.num_triangles} tris, build {best:.2f} ms (events), scene_create {wall:.2f} ms (wall), {st.ploc_iterations} ploc rounds, {st.num_nodes} nodes, sah {st.sah_cost:.3f}")
PY
done 2>&1 | tee gpurun_out/build_s67.log
echo "== bench N=1"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s67_n1.json 2> gpurun_out/bench_s67_n1.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_s67_n1.json").read().strip().splitlines()[-1])
print("n1 value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), "e2e ms", round(j["e2e"]["ms_per_step"],2), "build", j["bvh_build_ms"])
PY
