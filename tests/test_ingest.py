"""Reference-pointer ingest (rtb_scene_create_from_primitives / rtb_scene_attach_lights): the reference hands
Bvh::Bvh a host vector of Primitive{Triangle*, Material*, Light*} holding DEVICE pointers (primitive.cuh:4-12,
bvh.cuh:17, main.cu:141-151) and fills Scene{bvh, num_lights, d_lights} afterwards (main.cu:156).  On the host build the
"device" arrays are host arrays, so the pointer gather, the deferred lights and the range checks run without a GPU;
tests/test_compat.py runs the same through a compiled main.cu-style program on the GPU."""
import ctypes as C

import numpy as np
import pytest

from rtcuda_b200 import capi
from conftest import area_light, make_desc, mean_rel_err, random_rays, small_scene_arrays, std_materials

TRI = np.dtype([("p0", np.float32, 3), ("e1", np.float32, 3), ("e2", np.float32, 3), ("n", np.float32, 3)])       # triangle.cuh:20
MAT = np.dtype([("albedo", np.float32, 3), ("ior", np.float32), ("type", np.int32)])                                  # material.cuh:10-25
LIGHT = np.dtype([("type", np.int32), ("pos", np.float32, 3), ("d_triangle", np.uint64), ("L", np.float32, 3), ("pad", np.int32)])  # light.cuh:9-28
PRIM = np.dtype([("d_triangle", np.uint64), ("d_mat", np.uint64), ("d_area_light", np.uint64)])                      # primitive.cuh:4-12


def reference_arrays(verts, mat_ids, light_ids, materials, nlights):
    n = len(mat_ids)
    v = verts.reshape(n, 3, 3).astype(np.float32)
    tri = np.zeros(n, TRI)
    tri["p0"] = v[:, 0]; tri["e1"] = v[:, 0] - v[:, 1]; tri["e2"] = v[:, 2] - v[:, 0]
    e1, e2 = tri["e1"], tri["e2"]
    tri["n"] = np.stack([e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1], e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2],
                         e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]], axis=1).astype(np.float32)  # host cross, no FMA (triangle.cuh:7)
    mats = np.zeros(len(materials), MAT)
    for i, m in enumerate(materials):
        mats[i] = ((m.albedo[0], m.albedo[1], m.albedo[2]), m.ior, m.type)
    lights = np.zeros(max(nlights, 1), LIGHT)
    prims = np.zeros(n, PRIM)
    prims["d_triangle"] = tri.ctypes.data + 48 * np.arange(n, dtype=np.uint64)
    prims["d_mat"] = mats.ctypes.data + 20 * mat_ids.astype(np.uint64)
    for i in range(n):
        if light_ids[i] >= 0:
            prims["d_area_light"][i] = lights.ctypes.data + 40 * int(light_ids[i])
            lights[light_ids[i]] = (1, (0, 0, 0), tri.ctypes.data + 48 * i, (10, 10, 10), 0)
    return tri, mats, lights, prims


@pytest.fixture(scope="module")
def setup(emu):
    verts, mat, lid = small_scene_arrays(seed=4, n=300)
    ms = std_materials()
    desc, keep = make_desc(verts, mat, lid, ms, [area_light(len(mat) - 2), area_light(len(mat) - 1)])
    ctx = emu.context(0)
    flat = ctx.scene(desc)
    return emu, ctx, flat, (verts, mat, lid, ms), keep


def from_primitives(L, ctx, prims, tri, mats, lights, nlights):
    h = C.c_void_p()
    rc = L.lib.rtb_scene_create_from_primitives(ctx.h, prims.ctypes.data_as(C.c_void_p), C.c_int64(len(prims)), tri.ctypes.data_as(C.c_void_p),
                                                mats.ctypes.data_as(C.c_void_p), len(mats), lights.ctypes.data_as(C.c_void_p) if nlights > 0 else None,
                                                nlights, None, C.byref(h))
    if rc != 0:
        return rc, None
    sc = capi.Scene.__new__(capi.Scene)
    sc.ctx, sc.L, sc.h = ctx, L, h
    return 0, sc


def test_pointer_ingest_equals_the_flat_description(setup):
    L, ctx, flat, (verts, mat, lid, ms), _ = setup
    tri, mats, lights, prims = reference_arrays(verts, mat, lid, ms, 2)
    rc, sc = from_primitives(L, ctx, prims, tri, mats, lights, 2)
    assert rc == 0
    rays = random_rays(20000, seed=2)
    assert (sc.trace_closest(rays) == flat.trace_closest(rays)).all()
    cam = L.camera_look_at((0.5, 0.5, 1.5), (0.5, 0.5, 0), (0, 1, 0), 40.0, 1.0)
    p = capi.render_params(L, width=48, height=48, spp=4, max_bounces=6)
    a, _ = sc.render(cam, p)
    b, _ = flat.render(cam, p)
    assert (a == b).all()


def test_lights_attached_after_the_build(setup):
    """Bvh::Bvh first, Scene{bvh, num_lights, d_lights} later (main.cu:151-156)"""
    L, ctx, flat, (verts, mat, lid, ms), _ = setup
    tri, mats, lights, prims = reference_arrays(verts, mat, lid, ms, 2)
    rc, sc = from_primitives(L, ctx, prims, tri, mats, lights, -1)
    assert rc == 0
    st = sc.stats()
    assert st.num_nodes > 0 and st.collapse_levels > 0  # Bvh::num_nodes / max_depth right after the constructor
    cam = L.camera_look_at((0.5, 0.5, 1.5), (0.5, 0.5, 0), (0, 1, 0), 40.0, 1.0)
    p = capi.render_params(L, width=48, height=48, spp=4, max_bounces=6)
    dark, _ = sc.render(cam, p)  # no lights yet: nothing emits, nothing is lit
    assert dark.max() == 0.0
    L.check(L.lib.rtb_scene_attach_lights(sc.h, lights.ctypes.data_as(C.c_void_p), 2))
    a, _ = sc.render(cam, p)
    b, _ = flat.render(cam, p)
    assert (a == b).all()
    # a different light array: one light only, twice as bright
    lights2 = lights[:1].copy(); lights2["L"] = 20
    prims2 = prims.copy()
    L.check(L.lib.rtb_scene_attach_lights(sc.h, lights.ctypes.data_as(C.c_void_p), 2))  # idempotent
    c, _ = sc.render(cam, p)
    assert (c == b).all()
    # attaching to a scene that was not created with deferred lights is an error
    assert L.lib.rtb_scene_attach_lights(flat.h, lights.ctypes.data_as(C.c_void_p), 2) == -1


def test_pointers_outside_the_arrays_are_rejected(setup):
    L, ctx, flat, (verts, mat, lid, ms), _ = setup
    tri, mats, lights, prims = reference_arrays(verts, mat, lid, ms, 2)
    bad = prims.copy(); bad["d_mat"][3] += 20 * 100  # beyond d_materials
    rc, _ = from_primitives(L, ctx, bad, tri, mats, lights, 2)
    assert rc == -1 and b"d_mat" in L.lib.rtb_last_error()
    bad = prims.copy(); bad["d_area_light"][5] = lights.ctypes.data + 40 * 7  # beyond d_lights
    rc, _ = from_primitives(L, ctx, bad, tri, mats, lights, 2)
    assert rc == -1 and b"d_area_light" in L.lib.rtb_last_error()
    bad = prims.copy(); bad["d_mat"][0] += 4  # misaligned
    rc, _ = from_primitives(L, ctx, bad, tri, mats, lights, 2)
    assert rc == -1
    lbad = lights.copy(); lbad["d_triangle"][0] = tri.ctypes.data + 48 * (len(tri) + 5)
    rc, _ = from_primitives(L, ctx, prims, tri, mats, lbad, 2)
    assert rc == -1 and b"area light" in L.lib.rtb_last_error()
    rc, sc = from_primitives(L, ctx, bad := prims.copy(), tri, mats, lights, -1)
    assert rc == 0
    bad_l = np.zeros(1, LIGHT)  # the primitives point at OTHER lights than the array attached
    assert L.lib.rtb_scene_attach_lights(sc.h, bad_l.ctypes.data_as(C.c_void_p), 1) == -1
