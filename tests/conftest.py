import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from rtcuda_b200 import capi  # noqa: E402

EMU_LIB = os.path.join(ROOT, "tests", "emu", "librtb_emu.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def build_emu():
    """host build of the kernel bodies (test infrastructure, see tests/emu/emu_backend.cpp)"""
    srcs = [os.path.join(ROOT, "tests/emu/emu_backend.cpp"), os.path.join(ROOT, "rtcuda_b200/csrc/host/host_scene.cpp"),
            os.path.join(ROOT, "rtcuda_b200/csrc/host/host_util.cpp")]
    deps = srcs + [os.path.join(ROOT, "rtcuda_b200/csrc", f) for f in os.listdir(os.path.join(ROOT, "rtcuda_b200/csrc"))
                   if f.endswith(".h")] + [os.path.join(ROOT, "include/rtb.h")]
    if os.path.exists(EMU_LIB) and all(os.path.getmtime(EMU_LIB) >= os.path.getmtime(d) for d in deps):
        return EMU_LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden",
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "rtcuda_b200/csrc")] + srcs +
                          ["-o", EMU_LIB])
    return EMU_LIB


@pytest.fixture(scope="session")
def emu():
    return capi.Lib(build_emu())


@pytest.fixture(scope="session")
def gpu():
    """the shipped CUDA library through its C ABI; no fallback of any kind"""
    L = capi.Lib()  # raises if librtb.so is missing
    return L


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    return binding.Oracle()


@pytest.fixture(scope="session")
def bunny(emu):
    return emu.load_mesh()


def mean_rel_err(a, b):
    """mean relative error used for image parity: mean|a-b| / mean|b|"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).mean() / max(np.abs(b).mean(), 1e-30))


def small_scene_arrays(seed=0, n=200, two_lights=True):
    """random triangle soup inside the unit box with a floor and an emitter"""
    rng = np.random.default_rng(seed)
    c = rng.random((n, 1, 3)).astype(np.float32) * np.float32(0.8) + np.float32(0.1)
    v = (c + (rng.random((n, 3, 3)).astype(np.float32) - np.float32(0.5)) * np.float32(0.25)).astype(np.float32)
    v[..., 2] -= np.float32(1.0)
    floor = np.array([[[0, 0, 0], [1, 0, 0], [1, 0, -1]], [[0, 0, 0], [0, 0, -1], [1, 0, -1]]], np.float32)
    light = np.array([[[0.3, 0.99, -0.3], [0.7, 0.99, -0.3], [0.7, 0.99, -0.7]],
                      [[0.3, 0.99, -0.3], [0.3, 0.99, -0.7], [0.7, 0.99, -0.7]]], np.float32)
    verts = np.concatenate([v, floor, light]).reshape(-1, 9)
    nt = len(verts)
    mat = (np.arange(nt) % 3).astype(np.int32)
    mat[-2:] = 0
    lid = np.full(nt, -1, np.int32)
    lid[-2] = 0
    if two_lights:
        lid[-1] = 1
    return verts, mat, lid


def make_desc(verts, mat_ids, light_ids, materials, lights):
    """build a capi.SceneDesc from numpy arrays; returns (desc, keepalive)"""
    import ctypes as C
    verts = np.ascontiguousarray(verts, np.float32)
    mat_ids = np.ascontiguousarray(mat_ids, np.int32)
    light_ids = np.ascontiguousarray(light_ids, np.int32)
    M = (capi.Material * len(materials))(*materials)
    Ls = (capi.Light * max(len(lights), 1))(*lights)
    d = capi.SceneDesc()
    d.num_triangles = len(mat_ids)
    d.vertices = verts.ctypes.data_as(C.c_void_p).value if len(mat_ids) else None
    d.material_ids = mat_ids.ctypes.data_as(C.c_void_p).value if len(mat_ids) else None
    d.light_ids = light_ids.ctypes.data_as(C.c_void_p).value if len(mat_ids) else None
    d.num_materials = len(materials)
    d.materials = C.cast(M, C.c_void_p).value
    d.num_lights = len(lights)
    d.lights = C.cast(Ls, C.c_void_p).value if lights else None
    return d, (verts, mat_ids, light_ids, M, Ls)


def std_materials():
    def m(t, r, g, b, ior=0.0):
        x = capi.Material(); x.albedo[0] = r; x.albedo[1] = g; x.albedo[2] = b; x.ior = ior; x.type = t
        return x
    return [m(0, 0.7, 0.6, 0.5), m(1, 0.9, 0.9, 0.9), m(2, 0, 0, 0, 1.5)]


def area_light(tri, L=10.0):
    l = capi.Light(); l.type = 1; l.triangle = tri; l.L[0] = l.L[1] = l.L[2] = L
    return l


def random_rays(n, seed=1, tmax=None):
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=capi.RAY_DTYPE)
    rays["origin"] = rng.random((n, 3)).astype(np.float32) * np.float32(1.4) - np.float32(0.2)
    rays["origin"][:, 2] -= np.float32(0.9)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    rays["dir"] = d
    rays["tmax"] = np.float32(3.0e38) if tmax is None else tmax
    return rays
