#!/bin/bash
# shade kernels with MUFU-form 1/x, sqrt, 1/sqrt, x/y (default) vs the IEEE-rounded forms (RTB_EXACT_SHADE build)
mkdir -p gpurun_out
echo "== gpu tests (default build)"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c2 c3 c4; do
for lib in librtb_exact.so librtb.so librtb_exact.so librtb.so; do
echo "== $w $lib"; RTB_LIB=$PWD/rtcuda_b200/$lib timeout 600 python tools/sweep.py --workload $w --reps 3 2>&1 | tail -1 | cut -c1-200
done; done 2>&1 | tee gpurun_out/sweep_s68.log
echo "== image error vs oracle, both builds"
for lib in librtb_exact.so librtb.so; do RTB_LIB=$PWD/rtcuda_b200/$lib python - <<PY
import sys; sys.path.insert(0, ".")
import numpy as np
from rtcuda_b200 import capi
from oracle import binding
L = capi.Lib(); ctx = L.context(0); v, f = L.load_mesh()
for kind, depth in ((1, 8), (2, 16), (4, 8)):
    hs = L.host_scene(kind, v, f); sc = ctx.scene(hs.desc); orc = binding.Oracle().scene(hs.desc); cam = hs.camera(1.0)
    p = capi.render_params(L, width=128, height=128, spp=8, max_bounces=depth)
    img, _ = sc.render(cam, p); ref, _, _ = orc.render(cam, p)
    print("$lib", "scene kind", kind, "mean rel err vs oracle %.3e" % (np.abs(img.astype(np.float64) - ref).mean() / np.abs(ref).mean()))
PY
done 2>&1 | tee gpurun_out/shade_err_s68.log
