"""Two-level BVH (rtb_scene_create_instanced) against the FLATTENED scene it stands for.

The reference has no instancing (main.cu:67-86 transforms every vertex on the host), so the check is a chain:
instanced scene == flat scene of rtb_instanced_flatten (this file), flat scene == oracle == reference
(test_parity.py, test_golden.py; the oracle also renders the flattened scene here).  An instance is traversed in
object space: the ray transform rounds, so the bars are those of BASELINE.json's north_star rather than bit-exactness:
hit ids equal except on near-ties, t within 1e-5 (of max(t, scene size 1)), images within 1e-3 mean relative error
with identical per-pixel RNG streams.  Every test runs on the host build of the kernel bodies (tests/emu) and,
with -m gpu, on the shipped CUDA library, where the persistent k_trace<.., INST> kernel does the traversal.
"""
import ctypes as C

import numpy as np
import pytest

from rtcuda_b200 import capi
from conftest import area_light, make_desc, mean_rel_err, random_rays, std_materials

T_TOL = 1e-5       # |t_inst - t_flat| <= T_TOL * max(t, 1): the scenes are of unit size
IMAGE_TOL = 1e-3   # mean relative error, north_star
TIE_FRACTION = 2e-4  # rays allowed to report the other triangle of a near-tie


@pytest.fixture(scope="module", params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def L(request):
    return request.getfixturevalue(request.param)


@pytest.fixture(scope="module")
def ctx(L):
    return L.context(0)


@pytest.fixture(scope="module")
def field(L, ctx, bunny):
    """2 x 2 bunnies + Cornell shell: instanced, its flattening, both on the device"""
    hi = L.host_scene_instanced(capi.RTB_SCENE_S2, *bunny, grid=2)
    flat = L.flatten(hi.idesc)
    return hi, flat, ctx.scene(hi.idesc), ctx.scene(flat.desc)


def assert_hits_close(h, ref, tie_fraction=TIE_FRACTION):
    """same triangle -> t within T_TOL, u / v close.  A different triangle at the SAME distance is a tie (a ray through
    the shared edge of two triangles: pixel-centre rays of the symmetric camera do that on the wall diagonals; which
    one wins depends on the order of the tests, i.e. on the tree).  Anything else may only happen within rounding of
    an edge: there a ray takes the neighbour, or slips through the crack between two triangles — the reference's
    triangle test (triangle.cuh:39-58) is not watertight, in either space."""
    differ = h["prim"] != ref["prim"]
    same = ~differ & (ref["prim"] >= 0)
    dt = np.abs(h["t"][same].astype(np.float64) - ref["t"][same])
    assert (dt <= T_TOL * np.maximum(ref["t"][same], 1.0)).all(), dt.max()
    # u, v: the rounding of the object-space test relative to the SIZE of the triangle (a bunny triangle of the
    # 12 x 12 field is 5e-4 across, less when seen at a grazing angle)
    duv = np.maximum(np.abs(h["u"][same] - ref["u"][same]), np.abs(h["v"][same] - ref["v"][same]))
    assert np.quantile(duv, 0.99) <= 1e-3 and duv.max() <= 5e-2
    both = (h["prim"] >= 0) & (ref["prim"] >= 0)
    tie = differ & both & (np.abs(h["t"].astype(np.float64) - ref["t"]) <= T_TOL * np.maximum(ref["t"], 1.0))
    assert tie.mean() <= 1e-3, f"{tie.sum()} ties of {len(h)}"
    hard = differ & ~tie
    assert hard.mean() <= tie_fraction, f"{hard.sum()} of {len(h)} hits differ"

    def edge(x):
        e = np.minimum(np.minimum(x["u"], x["v"]), 1 - x["u"] - x["v"])
        return np.where(x["prim"] >= 0, e, 1.0)
    assert ((edge(ref[hard]) <= 1e-3) | (edge(h[hard]) <= 1e-3)).all()


def test_flattening_reproduces_the_flat_generator(emu, bunny):
    """the instanced S1 / S2 flatten to exactly the triangles rtb_host_scene_build makes, in the same order"""
    for kind, grid in ((capi.RTB_SCENE_S1, 0), (capi.RTB_SCENE_S2, 2)):
        hi = emu.host_scene_instanced(kind, *bunny, grid=grid)
        hf, hb = emu.host_scene(kind, *bunny, grid=grid), emu.flatten(hi.idesc)  # (the arrays are views: keep the owners)
        a, b = hf.arrays(), hb.arrays()
        assert a["vertices"].shape == b["vertices"].shape
        assert (a["vertices"].view(np.uint32) == b["vertices"].view(np.uint32)).all()
        assert (a["material_ids"] == b["material_ids"]).all() and (a["light_ids"] == b["light_ids"]).all()
        assert [int(l["triangle"]) for l in a["lights"]] == [int(l["triangle"]) for l in b["lights"]]


def test_stats_count_what_is_stored(field):
    hi, flat, si, sf = field
    st = si.stats()
    assert st.num_instances == 5 and st.num_triangles == 69451 + 12 and st.num_flat_triangles == 4 * 69451 + 12 == flat.desc.num_triangles
    assert 1 <= st.num_top_nodes <= 5 and st.num_nodes > st.num_top_nodes
    assert st.node_bytes + st.triangle_bytes < (sf.stats().node_bytes + sf.stats().triangle_bytes) / 3


def test_closest_hits_match_the_flattened_scene(L, field):
    hi, flat, si, sf = field
    rays = np.concatenate([L.primary_rays(flat.camera(16 / 9), 320, 180), random_rays(60000, seed=11)])
    h, ref = si.trace_closest(rays), sf.trace_closest(rays)
    assert (ref["prim"] >= 0).mean() > 0.5
    assert_hits_close(h, ref)


def test_any_hit_and_excluded_ids_use_the_flattened_numbering(field):
    hi, flat, si, sf = field
    rays = random_rays(40000, seed=5)
    ref = sf.trace_closest(rays)
    hit = ref["prim"] >= 0
    assert (si.trace_any(rays) == sf.trace_any(rays)).mean() >= 1 - TIE_FRACTION
    # a ray cut just behind its closest hit is occluded by that triangle alone: excluding it (by its FLATTENED id,
    # which names one instance's copy of the mesh triangle) clears the ray, excluding the same triangle of ANOTHER
    # instance does not
    r = rays[hit].copy()
    r["tmax"] = ref["t"][hit] + np.float32(3 * T_TOL)
    occ = si.trace_any(r)
    assert occ.mean() >= 1 - TIE_FRACTION
    ex = ref["prim"][hit].astype(np.int32)
    cleared = si.trace_any(r, ex)
    assert (cleared == sf.trace_any(r, ex)).mean() >= 1 - 5e-3  # (grazing second hits inside the slack)
    assert cleared.mean() < 0.05
    in_bunny = ex < 4 * 69451
    other = np.where(in_bunny, (ex + 69451) % (4 * 69451), ex).astype(np.int32)
    assert si.trace_any(r[in_bunny], other[in_bunny]).mean() >= 1 - TIE_FRACTION


def test_render_matches_the_flattened_scene(L, field):
    hi, flat, si, sf = field
    cam = flat.camera(16 / 9)
    p = capi.render_params(L, width=96, height=54, spp=4, max_bounces=8)
    a, sa = si.render(cam, p)
    b, sb = sf.render(cam, p)
    assert sa.paths == sb.paths
    assert abs(int(sa.extend_rays) - int(sb.extend_rays)) <= 1e-3 * sb.extend_rays
    assert mean_rel_err(a, b) <= IMAGE_TOL


def test_feature_buffers_match_the_flattened_scene(field):
    hi, flat, si, sf = field
    cam = flat.camera(1.0)
    al, no, de, pr = si.render_aovs(cam, 96, 96)
    al2, no2, de2, pr2 = sf.render_aovs(cam, 96, 96)
    same = pr == pr2
    tie = ~same & (pr >= 0) & (pr2 >= 0) & (np.abs(de - de2) <= T_TOL * np.maximum(de2, 1.0))  # the wall diagonals
    assert (same | tie).mean() >= 1 - 1e-3 and tie.mean() <= 1e-2
    assert np.abs(al - al2)[same].max() == 0
    assert np.abs(no - no2)[same].max() <= 1e-4
    assert np.abs(de - de2)[same].max() <= T_TOL * max(1.0, float(de2.max()))


def general_instanced_scene(seed=3):
    """a triangle soup placed six times — rotations, a NON-uniform scale, a mirror image (negative determinant), per
    instance material overrides of every BSDF type — over a static mesh with a floor and two emitters"""
    rng = np.random.default_rng(seed)
    n = 120
    c = (rng.random((n, 1, 3)).astype(np.float32) - np.float32(0.5)) * np.float32(0.9)
    soup = (c + (rng.random((n, 3, 3)).astype(np.float32) - np.float32(0.5)) * np.float32(0.5)).astype(np.float32)
    static = np.array([[[0, 0, 0], [1, 0, 0], [1, 0, -1]], [[0, 0, 0], [0, 0, -1], [1, 0, -1]],
                       [[0, 0, -1], [1, 0, -1], [1, 1, -1]], [[0, 0, -1], [0, 1, -1], [1, 1, -1]],
                       [[0.3, 0.99, -0.3], [0.7, 0.99, -0.3], [0.7, 0.99, -0.7]],
                       [[0.3, 0.99, -0.3], [0.3, 0.99, -0.7], [0.7, 0.99, -0.7]]], np.float32)
    verts = np.concatenate([soup, static]).reshape(-1, 9)
    nt = len(verts)
    mat = np.zeros(nt, np.int32)
    mat[:n] = np.arange(n) % 3
    lid = np.full(nt, -1, np.int32)
    lid[-2], lid[-1] = 0, 1
    mats = std_materials() + [glossy()]
    lights = [area_light(nt - 2), area_light(nt - 1)]
    desc, keep = make_desc(verts, mat, lid, mats, lights)
    mesh_first = np.array([0, n, nt], np.int64)

    def xf(rot_axis, ang, scale, trans):
        a = np.asarray(rot_axis, np.float64); a /= np.linalg.norm(a)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        M = R @ np.diag(scale)
        return np.concatenate([M, np.asarray(trans, np.float64)[:, None]], axis=1).astype(np.float32).reshape(-1)

    placements = [
        (0, -1, xf((0, 1, 0), 0.7, (0.22, 0.22, 0.22), (0.25, 0.25, -0.3))),
        (0, 1, xf((1, 0, 0), 2.1, (0.2, 0.2, 0.2), (0.7, 0.3, -0.35))),     # all mirror
        (0, 2, xf((1, 1, 0), 1.3, (0.25, 0.12, 0.2), (0.5, 0.6, -0.6))),    # non-uniform scale, all glass
        (0, 0, xf((0, 0, 1), 0.4, (-0.2, 0.2, 0.2), (0.3, 0.7, -0.7))),     # mirror image, all matte
        (0, 3, xf((1, 2, 3), 4.0, (0.18, 0.18, 0.3), (0.75, 0.7, -0.75))),  # all glossy
        (0, -1, xf((0, 1, 0), 0.0, (0.15, 0.15, 0.15), (0.55, 0.2, -0.2))),
        (1, -1, np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)),
    ]
    inst = (capi.Instance * len(placements))()
    for k, (m, ma, x) in enumerate(placements):
        inst[k].mesh = m; inst[k].material = ma
        for j in range(12):
            inst[k].xform[j] = float(x[j])
    idesc = capi.InstancedSceneDesc()
    idesc.geometry = desc
    idesc.num_meshes = 2
    idesc.mesh_first = mesh_first.ctypes.data_as(C.c_void_p).value
    idesc.num_instances = len(placements)
    idesc.instances = C.cast(inst, C.c_void_p).value
    return idesc, (keep, mesh_first, inst, desc)


def glossy():
    m = capi.Material(); m.albedo[0] = 0.8; m.albedo[1] = 0.7; m.albedo[2] = 0.5; m.ior = 30.0; m.type = capi.RTB_GLOSSY
    return m


def test_general_transforms_and_material_overrides_match_the_oracle(L, ctx, oracle):
    """instanced product == flat product == oracle on the flattened scene (rotations, non-uniform scale, a mirror
    image, per-instance materials of all four BSDF types, depth 12 with roulette)"""
    idesc, keep = general_instanced_scene()
    si = ctx.scene(idesc)
    flat = L.flatten(idesc)
    sf = ctx.scene(flat.desc)
    osc = oracle.scene(flat.desc)
    cam = L.camera_look_at((0.5, 0.5, 1.4), (0.5, 0.45, -0.5), (0, 1, 0), 40.0, 1.0)
    rays = np.concatenate([L.primary_rays(cam, 128, 128), random_rays(30000, seed=2)])
    h, ref = si.trace_closest(rays), osc.trace_closest(rays, capi.HIT_DTYPE)
    assert (ref["prim"] >= 0).mean() > 0.3
    assert_hits_close(h, ref, tie_fraction=1e-3)
    flat_h = sf.trace_closest(rays)
    assert (flat_h["prim"] == ref["prim"]).all()
    p = capi.render_params(L, width=64, height=64, spp=8, max_bounces=12, rr_start=2)
    a, sa = si.render(cam, p)
    rimg, _, ost = osc.render(cam, p)
    assert sa.paths == ost[0]
    assert mean_rel_err(a, rimg) <= IMAGE_TOL
    # the beyond-the-reference estimator works on instanced scenes like on flat ones
    p2 = capi.render_params(L, width=48, height=48, spp=4, max_bounces=8, flags=capi.RTB_RENDER_TRUE_MIS | capi.RTB_RENDER_RR_TERMINATE,
                            env_L=(0.2, 0.3, 0.4))
    a2, _ = si.render(cam, p2)
    r2, _, _ = osc.render(cam, p2)
    assert mean_rel_err(a2, r2) <= IMAGE_TOL


def test_descriptions_the_builder_must_reject(L, ctx):
    idesc, keep = general_instanced_scene()
    inst = keep[2]

    def fails(msg):
        with pytest.raises(capi.RtbError, match=msg):
            ctx.scene(idesc)

    saved = inst[0].xform[0], inst[0].xform[4], inst[0].xform[8]
    inst[0].xform[0] = inst[0].xform[4] = inst[0].xform[8] = 0.0  # a zero column: not invertible
    fails("not invertible")
    inst[0].xform[0], inst[0].xform[4], inst[0].xform[8] = saved
    inst[0].mesh = 2
    fails("mesh out of range")
    inst[0].mesh = 0
    inst[1].material = 9
    fails("material out of range")
    inst[1].material = 1
    inst[5].mesh = 1  # the static mesh (emitters) a second time
    fails("emissive triangle")
    inst[5].mesh = 0
    inst[6].xform[3] = 0.25  # the emitters' mesh moved
    fails("emissive triangle")
    inst[6].xform[3] = 0.0
    mf = keep[1]
    mf[1] = 0
    fails("ascend")
    mf[1] = 120
    ctx.scene(idesc)  # intact again


@pytest.mark.gpu
def test_persistent_kernel_matches_one_thread_per_ray(gpu, bunny):
    """k_trace<3, false, INST> (two wavefronts, dynamic fetch, stepped triangle tests) against the plain per-ray loop"""
    hi = gpu.host_scene_instanced(capi.RTB_SCENE_S2, *bunny, grid=3)
    sc = gpu.context(0).scene(hi.idesc)
    cam = hi.camera(16 / 9)
    kw = dict(width=480, height=270, spp=4, max_bounces=8)
    a, sa = sc.render(cam, capi.render_params(gpu, **kw))
    b, sb = sc.render(cam, capi.render_params(gpu, flags=capi.RTB_RENDER_NONPERSISTENT, **kw))
    c, sc_ = sc.render(cam, capi.render_params(gpu, flags=capi.RTB_RENDER_SINGLE_PIPELINE, **kw))
    assert sa.extend_rays == sb.extend_rays == sc_.extend_rays and sa.shadow_rays == sb.shadow_rays == sc_.shadow_rays
    assert sa.fused_trace == 1 and sa.pipelines >= 2
    assert mean_rel_err(a, b) <= 1e-6 and mean_rel_err(c, b) <= 1e-6  # (only the order of the accumulation atomics differs)


@pytest.mark.gpu
def test_full_size_field_as_instances(gpu, bunny):
    """C3 / C5 as instances: 144 bunnies + shell = 10,000,956 flattened triangles held as 69,463; 4K primary hits
    equal those of the flattened 10 M-triangle scene"""
    hi = gpu.host_scene_instanced(capi.RTB_SCENE_S2, *bunny, grid=12)
    ctx = gpu.context(0)
    si = ctx.scene(hi.idesc)
    st = si.stats()
    assert st.num_instances == 145 and st.num_flat_triangles == 144 * 69451 + 12 and st.num_triangles == 69451 + 12
    assert st.node_bytes + st.triangle_bytes < 8 << 20
    flat = gpu.host_scene(capi.RTB_SCENE_S2, *bunny, grid=12)
    sf = ctx.scene(flat.desc)
    cam = flat.camera(16 / 9)
    rays = gpu.primary_rays(cam, 3840, 2160)[::7]
    rays = np.concatenate([rays, random_rays(300000, seed=9)])
    assert_hits_close(si.trace_closest(rays), sf.trace_closest(rays))
    p = capi.render_params(gpu, width=960, height=540, spp=2, max_bounces=8)
    a, sa = si.render(cam, p)
    b, sb = sf.render(cam, p)
    assert sa.paths == sb.paths
    assert mean_rel_err(a, b) <= IMAGE_TOL


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_affine_placements_match_the_flattened_scene(emu, seed):
    """two soups, nine instances with RANDOM invertible matrices (rotation, non-uniform scale, shear, some mirrored),
    random material overrides: hits and a small render equal those of the flattened scene (host build of the kernels)"""
    rng = np.random.default_rng(100 + seed)
    n1, n2 = 60, 90

    def soup(n):
        c = (rng.random((n, 1, 3)).astype(np.float32) - np.float32(0.5))
        return (c + (rng.random((n, 3, 3)).astype(np.float32) - np.float32(0.5)) * np.float32(0.6)).astype(np.float32)
    static = np.array([[[0, 0, 0], [1, 0, 0], [1, 0, -1]], [[0, 0, 0], [0, 0, -1], [1, 0, -1]],
                       [[0.3, 0.99, -0.3], [0.7, 0.99, -0.3], [0.7, 0.99, -0.7]]], np.float32)
    verts = np.concatenate([soup(n1), soup(n2), static]).reshape(-1, 9)
    nt = len(verts)
    mat = (np.arange(nt) % 4).astype(np.int32)
    mat[-3:] = 0
    lid = np.full(nt, -1, np.int32)
    lid[-1] = 0
    desc, keep = make_desc(verts, mat, lid, std_materials() + [glossy()], [area_light(nt - 1)])
    mesh_first = np.array([0, n1, n1 + n2, nt], np.int64)
    placements = []
    for k in range(9):
        M = rng.normal(size=(3, 3)) * 0.15 + np.diag(rng.choice([-1.0, 1.0], 3) * (0.12 + 0.15 * rng.random(3)))
        if abs(np.linalg.det(M)) < 1e-4:
            M += np.eye(3) * 0.2
        t = np.array([0.15 + 0.7 * rng.random(), 0.15 + 0.6 * rng.random(), -0.15 - 0.7 * rng.random()])
        placements.append((int(rng.integers(0, 2)), int(rng.integers(-1, 4)), np.concatenate([M, t[:, None]], axis=1).astype(np.float32).reshape(-1)))
    placements.append((2, -1, np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)))
    inst = (capi.Instance * len(placements))()
    for k, (m, ma, x) in enumerate(placements):
        inst[k].mesh = m; inst[k].material = ma
        for j in range(12):
            inst[k].xform[j] = float(x[j])
    idesc = capi.InstancedSceneDesc()
    idesc.geometry = desc
    idesc.num_meshes = 3
    idesc.mesh_first = mesh_first.ctypes.data_as(C.c_void_p).value
    idesc.num_instances = len(placements)
    idesc.instances = C.cast(inst, C.c_void_p).value
    ctx = emu.context(0)
    si = ctx.scene(idesc)
    flat = emu.flatten(idesc)
    sf = ctx.scene(flat.desc)
    assert si.stats().num_flat_triangles == flat.desc.num_triangles
    cam = emu.camera_look_at((0.5, 0.5, 1.4), (0.5, 0.45, -0.5), (0, 1, 0), 40.0, 1.0)
    rays = np.concatenate([emu.primary_rays(cam, 96, 96), random_rays(20000, seed=seed)])
    assert_hits_close(si.trace_closest(rays), sf.trace_closest(rays), tie_fraction=1e-3)
    occ_i, occ_f = si.trace_any(rays), sf.trace_any(rays)
    assert (occ_i == occ_f).mean() >= 1 - 1e-3
    p = capi.render_params(emu, width=40, height=40, spp=4, max_bounces=8, rr_start=2)
    a, sa = si.render(cam, p)
    b, sb = sf.render(cam, p)
    assert sa.paths == sb.paths
    assert mean_rel_err(a, b) <= IMAGE_TOL
