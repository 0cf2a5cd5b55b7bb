"""Smallest end-to-end invocation for compute-sanitizer: default scene, 96x96 primary rays, a 48x48x2spp render with each
traversal schedule (stepped / pooled, fused / separate launches, one / two wavefronts) and the beyond-the-reference flags.

    compute-sanitizer --tool memcheck python tools/sanitize_run.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtcuda_b200 import capi  # noqa: E402

L = capi.Lib()
hs = L.host_scene(capi.RTB_SCENE_S1_MIXED, *L.load_mesh())
cam = hs.camera(1.0)
for env in ({}, {"RTB_POOLED": "1"}, {"RTB_FUSED": "0", "RTB_TRI_STEP": "0"}, {"RTB_PIPELINES": "1"}):
    os.environ.update(env)
    ctx = L.context(0)
    sc = ctx.scene(hs.desc)
    hits = sc.trace_closest(L.primary_rays(cam, 96, 96))
    for flags, env_l in ((0, (0, 0, 0)), (capi.RTB_RENDER_TRUE_MIS | capi.RTB_RENDER_RR_TERMINATE, (0.2, 0.3, 0.4))):
        p = capi.render_params(L, width=288, height=288, spp=2, max_bounces=8, flags=flags, env_L=env_l, rr_start=1)
        img, st = sc.render(cam, p)
        print(env, flags, "pipelines", st.pipelines, "rays", st.extend_rays + st.shadow_rays, "mean", float(img.mean()), "hits", int((hits["prim"] >= 0).sum()))
    sc.close()
    for k in env:
        os.environ.pop(k)
print("sanitize_run: done")
