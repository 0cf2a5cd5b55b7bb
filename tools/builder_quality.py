"""quality of the GPU builder (PLOC -> 8-wide compressed BVH) against the reference's host builder
(full-sweep SAH binary BVH, bvh.cuh:30-219, as restated by the oracle) on the same scene and the same rays.

    python tools/builder_quality.py --scene s1            # bunny + Cornell box, 69,463 triangles
    python tools/builder_quality.py --scene s2 --grid 12  # 144-bunny field, 10,000,956 triangles (oracle build: minutes)

Prints, per ray set: nodes fetched and triangles tested per ray and the bytes they stand for
(reference layout: 64 B per sibling-pair visit + 72 B per triangle test through the Primitive indirection, SURVEY 8d;
ours: 80 B per node + 48 B per triangle).  The reference numbers include its unclamped slab test (aabb_intersector.cuh:24-36).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from rtcuda_b200 import capi  # noqa: E402
from oracle import binding  # noqa: E402  (test infrastructure: this tool is a measurement, not product code)
from conftest import random_rays  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="s1", choices=["s1", "s2"])
ap.add_argument("--grid", type=int, default=12)
ap.add_argument("--rays", type=int, default=500000)
a = ap.parse_args()
L = capi.Lib()
ctx = L.context(0)
hs = L.host_scene(capi.RTB_SCENE_S1 if a.scene == "s1" else capi.RTB_SCENE_S2, *L.load_mesh(), grid=a.grid if a.scene == "s2" else 0)
sc = ctx.scene(hs.desc)
bs = sc.stats()
t0 = time.perf_counter()
osc = binding.Oracle().scene(hs.desc)
t_ref = time.perf_counter() - t0
ref_nodes, ref_depth = osc.bvh_stats()
print(f"scene {a.scene}: {bs.num_triangles} triangles")
print(f"reference builder (host, full-sweep SAH, one core): {ref_nodes} binary nodes ({ref_nodes * 32 / 1e6:.1f} MB), depth {ref_depth}, {t_ref:.1f} s")
print(f"GPU builder (PLOC radius 16 -> BVH8): {bs.num_nodes} 8-wide nodes ({bs.node_bytes / 1e6:.1f} MB), {bs.collapse_levels} levels, "
      f"SAH {bs.sah_cost:.2f}, {bs.build_ms:.1f} ms")
w, h = (960, 540)
cam = hs.camera(w / h)
sets = {"primary rays 960x540": L.primary_rays(cam, w, h), f"{a.rays} incoherent rays": random_rays(a.rays, seed=3)}
print("| ray set | reference: pair visits / triangle tests / bytes per ray | ours: nodes / triangle tests / bytes per ray |")
print("|---|---|---|")
for name, rays in sets.items():
    rn, rt = osc.trace_counts(rays)
    on, ot = sc.trace_counts(rays)
    ha, hb = sc.trace_closest(rays), osc.trace_closest(rays, capi.HIT_DTYPE)
    diff = ha["prim"] != hb["prim"]
    ties = int((diff & (ha["t"].view(np.uint32) == hb["t"].view(np.uint32))).sum())  # two triangles at the very same t (a shared edge)
    print(f"| {name} | {rn:.2f} / {rt:.2f} / {rn * 64 + rt * 72:.0f} B | {on:.2f} / {ot:.2f} / {on * 80 + ot * 48:.0f} B | "
          f"hit ids differ on {int(diff.sum())} rays, {ties} of them exact-t ties between two triangles (legal, DESIGN.md 5)")
