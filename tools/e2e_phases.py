"""Wall-clock phases of the end-to-end step of bench.py (rtb_scene_create from host arrays, rtb_render to a host buffer):
scene create (upload + BVH build), render call, destroy, per iteration.

    python tools/e2e_phases.py --workload c2 --iters 8 [--option nn_tiled=0]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtcuda_b200 import capi  # noqa: E402

W = {"c2": (1, 0, 1920, 1080, 64, 8), "c3": (3, 12, 3840, 2160, 16, 8), "c4": (2, 0, 1920, 1080, 64, 16), "c1": (1, 0, 600, 600, 10, 10)}
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--iters", type=int, default=8)
ap.add_argument("--option", action="append", default=[], help="context option name=value (rtb_context_set_option)")
ap.add_argument("--no-render", action="store_true")
a = ap.parse_args()
kind, grid, w, h, spp, depth = W[a.workload]
L = capi.Lib()
hs = L.host_scene(kind, *L.load_mesh(), grid=grid)
cam = hs.camera(w / h)
ctx = L.context(0)
for o in a.option:
    k, v = o.split("=")
    ctx.set_option(k, int(v))
p = capi.render_params(L, width=w, height=h, spp=spp, max_bounces=depth)
for it in range(a.iters):
    t0 = time.perf_counter()
    sc = ctx.scene(hs.desc)
    t1 = time.perf_counter()
    bs = sc.stats()
    msg = f"{a.workload} it{it}: create {1e3 * (t1 - t0):.2f} ms (build_ms {bs.build_ms:.2f}, ploc rounds {bs.ploc_iterations}, nodes {bs.num_nodes}, sah {bs.sah_cost:.3f})"
    if not a.no_render:
        t2 = time.perf_counter()
        img, st = sc.render(cam, p)
        t3 = time.perf_counter()
        msg += f", render call {1e3 * (t3 - t2):.1f} (ms_total {st.ms_total:.1f})"
    t4 = time.perf_counter()
    sc.close()
    msg += f", destroy {1e3 * (time.perf_counter() - t4):.1f}"
    print(msg, flush=True)
