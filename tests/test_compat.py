"""The source-level drop-in (include/rtcuda_compat.cuh): a main.cu-style host
program — device arrays made by hand, Primitives holding device pointers,
Bvh(triangles, primitives), Scene, Camera, render() — must give the same image
as the flat-description path of the C ABI on the same scene."""
import os
import subprocess

import numpy as np
import pytest

from rtcuda_b200 import capi
from conftest import ROOT, mean_rel_err

EXE = os.path.join(ROOT, "tests", "compat", "compat_main")


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [capi.RTB_SCENE_S1, capi.RTB_SCENE_S1_MIXED])
def test_reference_style_program_matches_flat_api(gpu, bunny, tmp_path, kind):
    assert os.path.exists(EXE), "tests/compat/compat_main not built (run __graft_entry__.build())"
    hs = gpu.host_scene(kind, *bunny)
    sf, out = str(tmp_path / "s.rtbs"), str(tmp_path / "img.f32")
    hs.save(sf)
    w, h, spp, depth = 160, 120, 8, 10
    r = subprocess.run([EXE, sf, str(w), str(h), str(spp), str(depth), out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    img = np.fromfile(out, dtype=np.float32).reshape(h, w, 3)
    sc = gpu.context(0).scene(hs.desc)
    ref, _ = sc.render(hs.camera(w / h), capi.render_params(gpu, width=w, height=h, spp=spp, max_bounces=depth))
    assert mean_rel_err(img, ref) <= 1e-5


@pytest.mark.gpu
def test_cpp_caller_on_every_gpu_of_the_box(gpu, bunny, tmp_path):
    """a C++ program through rtb_render_multi (all GPUs of the box) and the main.cu-style program with RTB_DEVICES
    spanning them: both must equal the single-GPU image to 1e-5 (float sums in another order)"""
    import torch
    ndev = torch.cuda.device_count()
    hs = gpu.host_scene(capi.RTB_SCENE_S1, *bunny)
    sf = str(tmp_path / "s.rtbs")
    hs.save(sf)
    w, h, spp, depth = 160, 120, 8, 10
    sc = gpu.context(0).scene(hs.desc)
    ref, _ = sc.render(hs.camera(w / h), capi.render_params(gpu, width=w, height=h, spp=spp, max_bounces=depth))
    out = str(tmp_path / "multi.f32")
    r = subprocess.run([EXE, sf, str(w), str(h), str(spp), str(depth), out, "multi"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    print(r.stdout.strip())
    assert f"{ndev} GPU(s)" in r.stdout
    assert mean_rel_err(np.fromfile(out, dtype=np.float32).reshape(h, w, 3), ref) <= 1e-5
    out2 = str(tmp_path / "devices.f32")
    env = dict(os.environ, RTB_DEVICES=",".join(str(i) for i in range(ndev)))
    r = subprocess.run([EXE, sf, str(w), str(h), str(spp), str(depth), out2], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert mean_rel_err(np.fromfile(out2, dtype=np.float32).reshape(h, w, 3), ref) <= 1e-5


def test_compat_header_compiles_for_sm100a():
    """no GPU needed: the drop-in header builds with nvcc for sm_100a"""
    r = subprocess.run(["nvcc", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I" + os.path.join(ROOT, "include"),
                        "-c", os.path.join(ROOT, "tests/compat/compat_main.cu"), "-o", os.devnull], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
