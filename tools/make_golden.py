"""Generate the golden fixtures of tests/golden/ by running the UNMODIFIED
reference (oracle/_ref/ref_harness = the reference's own CUDA code built from
/root/reference, see oracle/Makefile) on a B200.

    gpurun -- python tools/make_golden.py gpurun_out/golden

The reference ships no golden vectors, so these outputs of the
reference itself are the pins for the oracle and for the product:
  s1_hits.npz      closest-hit (prim, t, u, v) of Bvh::traverse (bvh.cuh:251) for 300x300
                   pixel-centre camera rays + 60,000 random rays on the default scene
  s1_any.npz       any-hit results (bvh.cuh:306) for 60,000 shadow rays with excluded triangle
  s1_ref_128.npz   mean radiance of the reference's estimator, 128x128, depth 10,
                   16 passes x 1024 spp with seeds 1000.. (its stage kernels, render.cuh:84-328)
  s1mixed_ref_128.npz  same for the MATTE/MIRROR/GLASS scene (config C4), depth 16
  ref_bvh.json     node count / depth of the reference's SAH build on both scenes
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from rtcuda_b200 import capi  # noqa: E402
from oracle import binding  # noqa: E402
from conftest import random_rays  # noqa: E402


def harness(scene_file, *cmd):
    out = subprocess.run([binding.REF_HARNESS, scene_file] + [str(c) for c in cmd], capture_output=True, text=True, timeout=3000)
    if out.returncode != 0:
        raise RuntimeError(out.stderr[-2000:])
    return [json.loads(l[5:]) for l in out.stdout.splitlines() if l.startswith("JSON ")]


def shadow_rays(n, nt, seed):
    rng = np.random.default_rng(seed)
    rays = random_rays(n, seed=seed + 1)
    rays["origin"] = (rng.random((n, 3)).astype(np.float32) * np.float32(0.9) + np.float32(0.05)) * np.float32([1, 1, -1])
    target = np.array([0.5, 0.999, -0.5], np.float32) + (rng.random((n, 3)).astype(np.float32) - 0.5) * np.float32([0.2, 0, 0.2])
    d = target - rays["origin"]
    dist = np.linalg.norm(d, axis=1).astype(np.float32)
    rays["dir"] = (d / dist[:, None]).astype(np.float32)
    rays["tmax"] = dist * np.float32(1.0005)
    excl = rng.integers(nt - 2, nt, n).astype(np.int32)
    return rays, excl


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    L = capi.Lib()
    verts, faces = L.load_mesh()
    info = {}
    with tempfile.TemporaryDirectory() as td:
        for name, kind, depth in (("s1", capi.RTB_SCENE_S1, 10), ("s1mixed", capi.RTB_SCENE_S1_MIXED, 16)):
            hs = L.host_scene(kind, verts, faces)
            sf = os.path.join(td, name + ".rtbs")
            hs.save(sf)
            if name == "s1":
                cam = hs.camera(1.0)
                rays = np.concatenate([L.primary_rays(cam, 300, 300), random_rays(60000, seed=7)])
                rf, hf = os.path.join(td, "rays.bin"), os.path.join(td, "hits.bin")
                rays.tofile(rf)
                js = harness(sf, "trace", rf, hf)
                hits = np.fromfile(hf, dtype=capi.HIT_DTYPE)
                np.savez_compressed(os.path.join(outdir, "s1_hits.npz"), rays=rays, hits=hits)
                info["s1_trace"] = js[-1]
                srays, excl = shadow_rays(60000, hs.desc.num_triangles, 21)
                ef, of = os.path.join(td, "excl.bin"), os.path.join(td, "occ.bin")
                srays.tofile(rf); excl.tofile(ef)
                harness(sf, "any", rf, ef, of)
                np.savez_compressed(os.path.join(outdir, "s1_any.npz"), rays=srays, excluded=excl,
                                    occluded=np.fromfile(of, dtype=np.uint8))
            passes, spp = 16, 1024
            sumf = os.path.join(td, "sum.f32")
            js = harness(sf, "loop", 128, 128, spp, depth, passes, 1000, sumf)
            mean = (np.fromfile(sumf, dtype=np.float32).reshape(128, 128, 3).astype(np.float64) / (passes * spp)).astype(np.float32)
            np.savez_compressed(os.path.join(outdir, f"{name}_ref_128.npz"), mean_radiance=mean, spp=passes * spp, depth=depth)
            info[name + "_bvh"] = js[0]
            info[name + "_loop"] = {k: v for k, v in js[-1].items() if k not in ("pass_ms", "pass_rays")}
    with open(os.path.join(outdir, "ref_bvh.json"), "w") as f:
        json.dump(info, f, indent=1)
    print(json.dumps(info))


if __name__ == "__main__":
    main()
