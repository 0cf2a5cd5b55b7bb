"""Small fixed workload for ncu: builds the default scene and renders it
`--reps` times through the C ABI (no torch import).  Used for the launch list
and the `--set full` capture committed under profiles/.

    python tools/profile_run.py --workload c1 --reps 2
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtcuda_b200 import capi  # noqa: E402

W = {"c1": (1, 0, 600, 600, 10, 10), "c2": (1, 0, 1920, 1080, 64, 8), "c2s": (1, 0, 1920, 1080, 4, 8), "c2m": (1, 0, 1920, 1080, 16, 8),
     "c3s": (3, 12, 3840, 2160, 1, 8), "c4s": (2, 0, 1920, 1080, 4, 16),
     "c3is": (3, 12, 3840, 2160, 1, 8)}  # c3s given as instances (two-level BVH)

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c1")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--pool", type=int, default=0)
a = ap.parse_args()
kind, grid, w, h, spp, depth = W[a.workload]
L = capi.Lib()
ctx = L.context(0)
if a.workload == "c3is":
    hs = L.host_scene_instanced(kind, *L.load_mesh(), grid=grid)
    sc = ctx.scene(hs.idesc)
else:
    hs = L.host_scene(kind, *L.load_mesh(), grid=grid)
    sc = ctx.scene(hs.desc)
bs = sc.stats()
print(f"scene: {bs.num_triangles} tris, {bs.num_nodes} nodes, build {bs.build_ms:.2f} ms, sah {bs.sah_cost:.2f}, "
      f"ploc iters {bs.ploc_iterations}, levels {bs.collapse_levels}")
cam = hs.camera(w / h)
p = capi.render_params(L, width=w, height=h, spp=spp, max_bounces=depth, flags=a.flags, pool_size=a.pool)
for r in range(a.reps):
    img, st = sc.render(cam, p)
    rays = st.extend_rays + st.shadow_rays
    print(f"rep {r}: {st.ms_total:.2f} ms total, {st.pipelines} pipeline(s), trace {st.ms_extend + st.ms_shadow:.2f} ms ({st.extend_rays} extend + "
          f"{st.shadow_rays} shadow rays), shade+generate+control {st.ms_shade:.2f} ms, {st.iterations} iterations, {st.kernel_launches} launches, "
          f"{rays / st.ms_total * 1e-3:.1f} Mrays/s, mean {img.mean():.4f}")
