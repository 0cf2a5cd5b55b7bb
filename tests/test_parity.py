"""Parity of the product kernels against the CPU oracle (oracle/oracle.cpp, a
restatement of the reference's render path), through the C ABI.

Every test runs twice: on the shipped CUDA library (`-m gpu`, B200) and on the
host build of the same kernel bodies (tests/emu, no GPU needed).  Bars:
hit triangle bit-exact, t/u/v bit-exact (stricter than the 1e-5 of
BASELINE.json), occlusion bit-exact, image mean relative error <= 1e-3 with
identical per-pixel RNG streams (BASELINE.json north_star).
"""
import ctypes as C

import numpy as np
import pytest

from rtcuda_b200 import capi
from conftest import (area_light, make_desc, mean_rel_err, random_rays, small_scene_arrays, std_materials)

IMAGE_TOL = 1e-3  # mean relative error, north_star


@pytest.fixture(scope="module", params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def L(request):
    return request.getfixturevalue(request.param)


@pytest.fixture(scope="module")
def ctx(L):
    return L.context(0)


@pytest.fixture(scope="module")
def s1(L, bunny):
    return L.host_scene(capi.RTB_SCENE_S1, *bunny)


@pytest.fixture(scope="module")
def s1_dev(ctx, s1):
    return ctx.scene(s1.desc)


@pytest.fixture(scope="module")
def s1_orc(oracle, s1):
    return oracle.scene(s1.desc)


def assert_hits_equal(hits, ref):
    bad = np.nonzero(hits["prim"] != ref["prim"])[0]
    assert len(bad) == 0, f"{len(bad)} hit-id mismatches, first: {bad[:5]} {hits[bad[:5]]} vs {ref[bad[:5]]}"
    for k in ("t", "u", "v"):
        assert (hits[k].view(np.uint32) == ref[k].view(np.uint32)).all(), f"{k} not bit-exact"


def test_primary_hits_default_scene(L, s1, s1_dev, s1_orc):
    """config C1 camera, pixel-centre rays at 600x600 (main.cu:159-166)"""
    cam = s1.camera(1.0)
    rays = L.primary_rays(cam, 600, 600)
    hits = s1_dev.trace_closest(rays)
    ref = s1_orc.trace_closest(rays, capi.HIT_DTYPE)
    assert (ref["prim"] >= 0).mean() > 0.9
    assert_hits_equal(hits, ref)


def test_incoherent_rays_default_scene(s1_dev, s1_orc):
    rays = random_rays(200000, seed=7)
    assert_hits_equal(s1_dev.trace_closest(rays), s1_orc.trace_closest(rays, capi.HIT_DTYPE))


def test_tmax_is_inclusive_and_limits_hits(s1_dev, s1_orc):
    """0 < t <= tmax (triangle.cuh:49): tmax equal to the hit distance still hits"""
    rays = random_rays(50000, seed=3)
    ref = s1_orc.trace_closest(rays, capi.HIT_DTYPE)
    hit = ref["prim"] >= 0
    r2 = rays[hit].copy()
    r2["tmax"] = ref["t"][hit]
    h2 = s1_dev.trace_closest(r2)
    assert (h2["prim"] == ref["prim"][hit]).all()
    r3 = r2.copy()
    r3["tmax"] = np.nextafter(ref["t"][hit], np.float32(0))
    h3 = s1_dev.trace_closest(r3)
    o3 = s1_orc.trace_closest(r3, capi.HIT_DTYPE)
    assert_hits_equal(h3, o3)
    assert (h3["prim"] != ref["prim"][hit]).all()


def test_any_hit_with_excluded_triangle(s1, s1_dev, s1_orc):
    """shadow rays towards the two emitters, excluding the emitter itself (render.cuh:196-197)"""
    rng = np.random.default_rng(5)
    n = 100000
    rays = random_rays(n, seed=11)
    rays["origin"] = (rng.random((n, 3)).astype(np.float32) * np.float32(0.9) + np.float32(0.05)) * np.float32([1, 1, -1])
    target = np.array([0.5, 0.999, -0.5], np.float32) + (rng.random((n, 3)).astype(np.float32) - 0.5) * np.float32([0.2, 0, 0.2])
    d = target - rays["origin"]
    dist = np.linalg.norm(d, axis=1).astype(np.float32)
    rays["dir"] = (d / dist[:, None]).astype(np.float32)
    rays["tmax"] = dist * np.float32(1.0005)  # reaches the emitter plane, not the ceiling
    nt = s1.desc.num_triangles
    excl = rng.integers(nt - 2, nt, n).astype(np.int32)
    occ = s1_dev.trace_any(rays, excl)
    ref = s1_orc.trace_any(rays, excl)
    assert 0.05 < ref.mean() < 0.95
    assert (occ == ref).all()
    occ2 = s1_dev.trace_any(rays, None)
    ref2 = s1_orc.trace_any(rays, None)
    assert (occ2 == ref2).all() and occ2.sum() >= occ.sum()


def test_axis_aligned_and_degenerate_rays(s1_dev, s1_orc):
    """direction components that are exactly zero (aabb_intersector.cuh:17-19 clamps them)"""
    rays = random_rays(30000, seed=13)
    d = rays["dir"].copy()
    d[0::3, 0] = 0
    d[1::3, 1] = 0
    d[2::3, (0, 2)] = 0
    d[5::7] = np.float32([0, -1, 0])
    nrm = np.linalg.norm(d, axis=1, keepdims=True)
    nrm[nrm == 0] = 1
    rays["dir"] = (d / nrm).astype(np.float32)
    assert_hits_equal(s1_dev.trace_closest(rays), s1_orc.trace_closest(rays, capi.HIT_DTYPE))


def test_non_finite_rays_miss_without_walking_the_tree(s1_dev, s1_orc):
    """a ray with a NaN / infinite component can hit nothing (every comparison of triangle.cuh:39-58 fails) but
    passes every slab test: the kernels must retire it at once instead of visiting all nodes and triangles"""
    rays = random_rays(64, seed=11)
    nan, inf = np.float32("nan"), np.float32("inf")
    rays["dir"][0] = (nan, nan, nan)
    rays["dir"][1] = (nan, 0.3, -0.9)
    rays["origin"][2] = (nan, 0.5, -0.5)
    rays["dir"][3] = (inf, 0.0, 0.0)
    rays["origin"][4] = (inf, 0.0, 0.0)
    hits = s1_dev.trace_closest(rays)
    ref = s1_orc.trace_closest(rays, capi.HIT_DTYPE)
    assert (hits["prim"][:5] == -1).all() and (ref["prim"][:5] == -1).all()
    assert_hits_equal(hits, ref)
    assert (s1_dev.trace_any(rays[:5]) == 0).all()
    nodes, tris = s1_dev.trace_counts(rays[:5])
    assert nodes == 0 and tris == 0


def test_slab_test_pads_each_axis_on_its_own(s1_dev, s1_orc):
    """a ray almost parallel to one axis has a huge (origin - plane) / d on that axis; the conservative pad of the
    slab test must come from each axis alone, or such a ray passes every box whose slab it starts in and walks a
    whole slice of the scene (round 1: seconds per ray on the 10 M-triangle scene).  Same hits as the oracle, and
    no more nodes than the same ray tilted a little."""
    base = np.zeros(4, capi.RAY_DTYPE)
    base["origin"] = (0.45, 0.2, 0.9)
    base["tmax"] = 3.0e38
    for k, dx in enumerate((1e-3, 1e-9, 1e-20, 1e-35)):
        d = np.array((dx, 0.02, -1.0), np.float64)
        base["dir"][k] = (d / np.linalg.norm(d)).astype(np.float32)
    assert_hits_equal(s1_dev.trace_closest(base), s1_orc.trace_closest(base, capi.HIT_DTYPE))
    counts = [s1_dev.trace_counts(base[k:k + 1])[0] for k in range(4)]
    assert max(counts) <= counts[0] + 2, counts


def test_sah_optimal_collapse_is_a_valid_smaller_tree(L, ctx, s1, s1_dev, s1_orc):
    """rtb_build_params.collapse = RTB_COLLAPSE_SAH_OPTIMAL (dynamic programme over the binary tree): fewer 8-wide nodes
    and a lower SAH cost than the default greedy collapse, the same hits — except on exact-t ties (a ray through the
    shared edge of two triangles), where either triangle is a correct answer and the choice follows the test order"""
    bp = capi.BuildParams()
    L.lib.rtb_build_params_default(C.byref(bp))
    assert bp.collapse == 0
    bp.collapse = 1
    sc = ctx.scene(s1.desc, bp)
    a, b = sc.stats(), s1_dev.stats()
    assert a.num_nodes < 0.8 * b.num_nodes and a.sah_cost <= b.sah_cost
    rays = np.concatenate([L.primary_rays(s1.camera(1.0), 300, 300), random_rays(100000, seed=9)])
    hits, ref = sc.trace_closest(rays), s1_orc.trace_closest(rays, capi.HIT_DTYPE)
    diff = hits["prim"] != ref["prim"]
    assert (hits["t"].view(np.uint32) == ref["t"].view(np.uint32)).all()  # t is bit-exact everywhere
    assert diff.sum() <= 1e-3 * len(rays)  # ties only
    occ = sc.trace_any(rays[:50000])
    assert (occ == s1_orc.trace_any(rays[:50000])).all()
    sc.close()


def test_build_schedules_give_the_same_tree(L, s1, s1_dev, s1_orc):
    """The builder runs its PLOC rounds and collapse levels in batches against counts that stay on the device, searches
    nearest neighbours from a shared-memory window and finishes the rounds in one block once the clusters fit it (round
    2).  None of that may change the tree: the one-thread-per-cluster search ("nn_tiled" = 0) and one launch set per
    round to the end ("ploc_tail" = 0) must give the same nodes, rounds, levels, cost and — ties included — hits."""
    ref = s1_dev.stats()
    rays = np.concatenate([L.primary_rays(s1.camera(1.0), 200, 200), random_rays(60000, seed=21)])
    want = s1_orc.trace_closest(rays, capi.HIT_DTYPE)
    for opts in ({"nn_tiled": 0}, {"ploc_tail": 0}, {"nn_tiled": 0, "ploc_tail": 0}):
        c = L.context(0)
        for k, v in opts.items():
            c.set_option(k, v)
        sc = c.scene(s1.desc)
        st = sc.stats()
        assert (st.num_nodes, st.ploc_iterations, st.collapse_levels) == (ref.num_nodes, ref.ploc_iterations, ref.collapse_levels), opts
        assert abs(st.sah_cost - ref.sah_cost) <= 1e-5 * ref.sah_cost  # (a float sum over the nodes in launch order)
        assert_hits_equal(sc.trace_closest(rays), want)
        sc.close()
        del sc, c


def test_identical_triangles_build_one_pair_per_round(L, ctx, oracle):
    """the builder's worst case: every cluster box is the same, area ties go to the lowest position, so a PLOC round
    merges exactly one pair — n - 1 rounds, through the batched rounds, the single-block tail and the collapse — and the
    binary tree is a chain.  200 copies still fit the traversal stack; 2500 do not and come back as an error status
    (never a hang: the rounds and the levels are bounded)"""
    tri = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], np.float32)
    mats = std_materials()

    def desc_of(n):
        return make_desc(np.tile(tri, (n, 1)), np.zeros(n, np.int32), np.full(n, -1, np.int32), mats, [])

    n = 200
    desc, keep = desc_of(n)
    sc = ctx.scene(desc)
    st = sc.stats()
    assert st.num_triangles == n and st.ploc_iterations == n - 1
    rays = np.zeros(2, capi.RAY_DTYPE)
    rays["origin"] = [(0.2, 0.2, 1.0), (2.0, 2.0, 1.0)]
    rays["dir"] = [(0, 0, -1), (0, 0, -1)]
    rays["tmax"] = 1e30
    hits = sc.trace_closest(rays)
    ref = oracle.scene(desc).trace_closest(rays, capi.HIT_DTYPE)
    assert hits["prim"][0] >= 0 and hits["prim"][1] == -1
    assert (hits["t"].view(np.uint32) == ref["t"].view(np.uint32)).all()  # (which of the coincident triangles wins the tie is the tree's choice)
    sc.close()
    desc, keep = desc_of(2500)
    with pytest.raises(capi.RtbError, match="too deep"):
        ctx.scene(desc)


@pytest.mark.parametrize("n", [1019, 1020, 1021, 1400, 2700, 5000])
def test_triangle_counts_around_the_builder_s_batch_and_tail_sizes(L, ctx, oracle, n):
    """n + 4 triangles (soup + floor + emitters) = 1023, 1024 and 1025 — the single-block tail takes over at <= 1024
    clusters — and sizes whose first batch is a fraction of a round, one round and several: hits equal the oracle's"""
    verts, mat, lid = small_scene_arrays(seed=100 + n, n=n)
    desc, keep = make_desc(verts, mat, lid, std_materials(), [area_light(len(verts) - 2), area_light(len(verts) - 1)])
    sc = ctx.scene(desc)
    assert sc.stats().num_triangles == n + 4
    rays = random_rays(20000, seed=n)
    assert_hits_equal(sc.trace_closest(rays), oracle.scene(desc).trace_closest(rays, capi.HIT_DTYPE))
    sc.close()


@pytest.mark.gpu
def test_gpu_builder_and_its_host_emulation_build_the_same_tree(gpu, emu, bunny):
    """same bodies, two drivers (batched launches on device-resident counts / plain loops): same tree statistics"""
    out = []
    for lib in (gpu, emu):
        hs = lib.host_scene(capi.RTB_SCENE_S1, *bunny)
        sc = lib.context(0).scene(hs.desc)
        st = sc.stats()
        out.append((st.num_nodes, st.ploc_iterations, st.collapse_levels, st.sah_cost))
        sc.close()
    assert out[0][:3] == out[1][:3], out
    assert abs(out[0][3] - out[1][3]) <= 1e-5 * out[1][3]


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 9, 33, 200])
def test_small_and_empty_scenes(L, ctx, oracle, n):
    """empty, single-triangle and ragged triangle counts; degenerate (zero-area) triangles included"""
    verts, mat, lid = small_scene_arrays(seed=n, n=200)
    verts, mat, lid = verts[-n:].copy() if n else verts[:0], mat[-n:] if n else mat[:0], np.full(n, -1, np.int32)
    if n >= 9:
        verts[3, 3:6] = verts[3, 0:3]  # zero-area triangle
        verts[4] = verts[5]            # duplicate triangle
    desc, keep = make_desc(verts, mat, lid, std_materials(), [])
    sc = ctx.scene(desc)
    st = sc.stats()
    assert st.num_triangles == n and st.num_nodes >= 1
    rays = random_rays(20000, seed=n + 1)
    hits = sc.trace_closest(rays)
    if n == 0:
        assert (hits["prim"] == -1).all()
        return
    osc = oracle.scene(desc)
    ref = osc.trace_closest(rays, capi.HIT_DTYPE)
    same_t = hits["t"].view(np.uint32) == ref["t"].view(np.uint32)
    assert same_t.all()
    # exact ties between duplicated triangles may resolve to either copy
    diff = hits["prim"] != ref["prim"]
    if diff.any():
        assert n >= 9 and set(hits["prim"][diff]) | set(ref["prim"][diff]) <= {n - 200 + 4 if False else 4, 5} | set(range(n))
        a, b = verts[hits["prim"][diff]], verts[ref["prim"][diff]]
        assert (a == b).all()
    occ = sc.trace_any(rays, None)
    assert (occ == osc.trace_any(rays, None)).all()


def render_pair(L, dev_scene, orc_scene, cam, **kw):
    p = capi.render_params(L, **kw)
    img, st = dev_scene.render(cam, p)
    ref, _, ost = orc_scene.render(cam, p)
    return img, st, ref, ost


def test_image_default_scene(L, s1, s1_dev, s1_orc):
    """config C1 at reduced size: 150x150, 8 spp, depth 10, seed 1, shared per-pixel RNG streams"""
    cam = s1.camera(1.0)
    img, st, ref, ost = render_pair(L, s1_dev, s1_orc, cam, width=150, height=150, spp=8, max_bounces=10)
    assert st.paths == ost[0] == 150 * 150 * 8
    assert abs(int(st.extend_rays) - int(ost[1])) <= 2e-3 * ost[1]
    assert abs(int(st.shadow_rays) - int(ost[2])) <= 2e-3 * ost[2]
    assert np.isfinite(img).all()
    assert mean_rel_err(img, ref) <= IMAGE_TOL


def test_image_mixed_materials_deep_paths(L, bunny, ctx, oracle):
    """config C4 at reduced size: MATTE/MIRROR/GLASS round-robin, depth 16 with Russian roulette"""
    hs = L.host_scene(capi.RTB_SCENE_S1_MIXED, *bunny)
    sc, osc = ctx.scene(hs.desc), oracle.scene(hs.desc)
    cam = hs.camera(4 / 3)
    img, st, ref, ost = render_pair(L, sc, osc, cam, width=128, height=96, spp=8, max_bounces=16)
    assert st.paths == ost[0]
    assert abs(int(st.extend_rays) - int(ost[1])) <= 5e-3 * ost[1]
    assert mean_rel_err(img, ref) <= IMAGE_TOL


def test_image_random_soup_with_point_light(L, ctx, oracle):
    verts, mat, lid = small_scene_arrays(seed=4, n=300)
    pl = capi.Light(); pl.type = 0; pl.pos[0] = 0.5; pl.pos[1] = 0.8; pl.pos[2] = -0.2; pl.L[0] = pl.L[1] = pl.L[2] = 0.5
    nt = len(mat)
    lights = [area_light(nt - 2), area_light(nt - 1), pl]
    desc, keep = make_desc(verts, mat, lid, std_materials(), lights)
    sc, osc = ctx.scene(desc), oracle.scene(desc)
    cam = L.camera_look_at((0.5, 0.5, 1.5), (0.5, 0.5, 0.0), (0, 1, 0), 37.8, 1.0)
    img, st, ref, ost = render_pair(L, sc, osc, cam, width=96, height=96, spp=16, max_bounces=12)
    assert st.paths == ost[0]
    assert mean_rel_err(img, ref) <= IMAGE_TOL


def test_russian_roulette_and_depth_limits(L, s1, s1_dev, s1_orc):
    """rr_start=0 makes roulette fire from the second bounce on (render.cuh:112-124, Quirk A);
    max_bounces=0 leaves only camera-visible emission (render.cuh:98-109)"""
    cam = s1.camera(1.0)
    img, st, ref, ost = render_pair(L, s1_dev, s1_orc, cam, width=96, height=96, spp=8, max_bounces=10, rr_start=0,
                                    rr_threshold=1.0)
    assert mean_rel_err(img, ref) <= IMAGE_TOL
    img0, st0, ref0, ost0 = render_pair(L, s1_dev, s1_orc, cam, width=96, height=96, spp=2, max_bounces=0)
    assert st0.shadow_rays == 0 and st0.extend_rays == 96 * 96 * 2 == ost0[1]
    assert mean_rel_err(img0, ref0) <= 1e-6
    assert (img0 > 0).any() and (img0 == 0).mean() > 0.9


# ---- beyond the reference: flags that are OFF in parity mode ----
@pytest.mark.parametrize("kind,depth", [(capi.RTB_SCENE_S1, 8), (capi.RTB_SCENE_S1_MIXED, 12)])
def test_true_mis_rr_termination_and_environment_match_the_oracle(L, bunny, ctx, oracle, kind, depth):
    """RTB_RENDER_TRUE_MIS | RTB_RENDER_RR_TERMINATE with a constant environment: the product and the oracle implement
    the same corrected estimator (power heuristic on float pdfs, emitters seen after a bounce weighed as the BSDF
    sample, roulette kills end the path, rays leaving the scene pick up env_L)"""
    hs = L.host_scene(kind, *bunny)
    sc, osc = ctx.scene(hs.desc), oracle.scene(hs.desc)
    cam = hs.camera(4 / 3)
    flags = capi.RTB_RENDER_TRUE_MIS | capi.RTB_RENDER_RR_TERMINATE
    img, st, ref, ost = render_pair(L, sc, osc, cam, width=128, height=96, spp=8, max_bounces=depth, flags=flags,
                                    rr_start=1, env_L=(0.3, 0.4, 0.6))
    assert st.paths == ost[0]
    assert abs(int(st.extend_rays) - int(ost[1])) <= 5e-3 * ost[1]
    assert abs(int(st.shadow_rays) - int(ost[2])) <= 5e-3 * ost[2]
    assert np.isfinite(img).all()
    assert mean_rel_err(img, ref) <= IMAGE_TOL
    # the flags change the estimator: not the parity image
    par, _ = sc.render(cam, capi.render_params(L, width=128, height=96, spp=8, max_bounces=depth))
    assert mean_rel_err(img, par) > 10 * IMAGE_TOL


@pytest.mark.parametrize("flags", [0, capi.RTB_RENDER_TRUE_MIS | capi.RTB_RENDER_RR_TERMINATE])
def test_glossy_material_matches_the_oracle(L, bunny, ctx, oracle, flags):
    """RTB_GLOSSY (not in the reference: energy-normalised Phong lobe, `ior` = exponent): glossy bunny (exponent 50) and a
    sharp glossy wall (exponent 400) in the Cornell box, with and without the MIS estimator"""
    hs = L.host_scene(capi.RTB_SCENE_S1_GLOSSY, *bunny)
    types = [int(m["type"]) for m in hs.arrays()["materials"]]
    assert types.count(capi.RTB_GLOSSY) == 2
    sc, osc = ctx.scene(hs.desc), oracle.scene(hs.desc)
    cam = hs.camera(4 / 3)
    img, st, ref, ost = render_pair(L, sc, osc, cam, width=128, height=96, spp=8, max_bounces=8, flags=flags)
    assert st.paths == ost[0]
    assert abs(int(st.extend_rays) - int(ost[1])) <= 5e-3 * ost[1]
    assert abs(int(st.shadow_rays) - int(ost[2])) <= 5e-3 * ost[2]
    assert np.isfinite(img).all() and img.mean() > 0.05
    assert mean_rel_err(img, ref) <= IMAGE_TOL
    sc.close()


def test_glossy_lobe_conserves_energy(L, ctx):
    """white furnace: a glossy floor of albedo 1 under a constant environment of radiance 1 reflects at most what it
    receives (the lobe is normalised; the part of it below the horizon is lost), for a broad and for a sharp lobe"""
    floor = np.array([[[-50, 0, 50], [50, 0, 50], [50, 0, -50]], [[-50, 0, 50], [50, 0, -50], [-50, 0, -50]]], np.float32).reshape(-1, 9)
    for exponent, lo in ((2.0, 0.6), (200.0, 0.6)):  # the Phong lobe's albedo falls like cos(theta) at oblique incidence
        m = capi.Material(); m.type = capi.RTB_GLOSSY; m.ior = exponent
        m.albedo[0] = m.albedo[1] = m.albedo[2] = 1.0
        desc, keep = make_desc(floor, np.zeros(2, np.int32), np.full(2, -1, np.int32), [m], [])
        sc = ctx.scene(desc)
        cam = L.camera_look_at((0.0, 1.0, 0.0), (0.0, 0.0, -1.0), (0, 1, 0), 40.0, 1.0)  # 45 degrees down onto the floor
        img, st = sc.render(cam, capi.render_params(L, width=16, height=16, spp=256, max_bounces=4, env_L=(1.0, 1.0, 1.0)))
        lower = (img[8:] ** 2).mean()  # bottom rows look at the floor
        assert lo <= lower <= 1.02, (exponent, lower)
        sc.close()


def test_true_mis_and_light_sampling_only_agree_in_the_mean(L, oracle):
    """direct lighting of a floor by a large, low, black-bodied emitter: the reference's estimator (weight-1 light
    sampling) and the MIS estimator (light sample + BSDF sample, complete at depth 2) are both unbiased,
    so the floor converges to the same radiance; at depth 1 MIS lacks its BSDF half (56 % of the light here), which
    shows that the test weighs that branch"""
    floor = [[[-3, 0, 3], [3, 0, 3], [3, 0, -3]], [[-3, 0, 3], [3, 0, -3], [-3, 0, -3]]]
    lamp = [[[-1, 0.3, 1], [1, 0.3, 1], [0.0, 0.3, -1]]]
    verts = np.array(floor + lamp, np.float32).reshape(-1, 9)
    grey, black = capi.Material(), capi.Material()
    grey.albedo[0] = grey.albedo[1] = grey.albedo[2] = 0.6
    desc, keep = make_desc(verts, np.array([0, 0, 1], np.int32), np.array([-1, -1, 0], np.int32), [grey, black], [area_light(2, 2.0)])
    osc = oracle.scene(desc)
    cam = L.camera_look_at((0.0, 0.25, 3.5), (0.0, 0.0, 0.0), (0, 1, 0), 50.0, 1.0)
    w = h = 32
    prim = osc.trace_closest(L.primary_rays(cam, w, h), capi.HIT_DTYPE)["prim"].reshape(h, w)
    on_floor = (prim == 0) | (prim == 1)
    mask = on_floor.copy()
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            mask &= np.roll(np.roll(on_floor, dy, 0), dx, 1)  # pixel jitter stays on the floor
    assert mask.sum() > 300

    def floor_mean(**kw):
        tot = 0.0
        for seed in (1, 5):
            _, acc, _ = osc.render(cam, capi.render_params(L, width=w, height=h, spp=1024, seed=seed, **kw), want_accum=True)
            tot += float(acc[mask].mean())
        return tot / 2

    nee = floor_mean(max_bounces=2)
    mis = floor_mean(max_bounces=2, flags=capi.RTB_RENDER_TRUE_MIS)
    mis_half = floor_mean(max_bounces=1, flags=capi.RTB_RENDER_TRUE_MIS)
    assert abs(mis - nee) <= 0.025 * nee, (mis, nee)
    assert mis_half < 0.6 * nee


def test_environment_light_reaches_open_scenes_only_through_misses(L, s1, s1_dev, s1_orc):
    """env_L adds beta * env for every ray that leaves the scene and nothing else: the image is the parity image plus a
    term linear in env_L"""
    cam = s1.camera(1.0)
    kw = dict(width=64, height=64, spp=4, max_bounces=6)
    base, _ = s1_dev.render(cam, capi.render_params(L, **kw))
    e1, _, r1, _ = render_pair(L, s1_dev, s1_orc, cam, env_L=(1.0, 1.0, 1.0), **kw)
    e2, _ = s1_dev.render(cam, capi.render_params(L, env_L=(2.0, 2.0, 2.0), **kw))
    assert mean_rel_err(e1, r1) <= IMAGE_TOL
    lin = lambda x: x.astype(np.float64) ** 2  # images are sqrt(sum / spp)
    assert (lin(e1) >= lin(base) - 1e-6).all() and lin(e1).sum() > lin(base).sum() * 1.01
    assert mean_rel_err(lin(e2) - lin(base), 2 * (lin(e1) - lin(base))) <= 1e-4


def test_render_edge_cases(L, ctx, oracle, s1, s1_dev, s1_orc):
    """empty scene, scene without lights, 1x1 image, one sample, depth 0, image sizes changing between calls on one scene,
    a second context alive at the same time"""
    cam = L.camera_look_at((0.5, 0.5, 1.5), (0.5, 0.5, 0.0), (0, 1, 0), 37.8, 1.0)
    empty, keep0 = make_desc(np.zeros((0, 9), np.float32), np.zeros(0, np.int32), np.zeros(0, np.int32), std_materials(), [])
    sc0 = ctx.scene(empty)
    img, st = sc0.render(cam, capi.render_params(L, width=8, height=4, spp=3, max_bounces=5))
    assert (img == 0).all() and st.paths == 96 and st.extend_rays == 96 and st.shadow_rays == 0
    img, st = sc0.render(cam, capi.render_params(L, width=8, height=4, spp=3, max_bounces=5, env_L=(4.0, 1.0, 0.25)))
    assert np.allclose(img, np.array([2.0, 1.0, 0.5], np.float32))  # sqrt(env): every camera ray leaves the scene
    verts, mat, lid = small_scene_arrays(seed=2, n=40)
    dark, keep1 = make_desc(verts, mat, np.full(len(mat), -1, np.int32), std_materials(), [])
    sc1, osc1 = ctx.scene(dark), oracle.scene(dark)
    img, st, ref, ost = render_pair(L, sc1, osc1, cam, width=33, height=17, spp=2, max_bounces=6)
    assert (img == 0).all() and (ref == 0).all() and st.shadow_rays == 0 and st.extend_rays == ost[1]
    ctx2 = L.context(0)
    sc2 = ctx2.scene(s1.desc)
    for (w, h, spp, depth) in ((1, 1, 1, 1), (1, 1, 5, 0), (7, 3, 1, 12), (64, 2, 2, 3), (2, 64, 2, 3)):
        c = s1.camera(w / h)
        a, sa, r, so = render_pair(L, s1_dev, s1_orc, c, width=w, height=h, spp=spp, max_bounces=depth)
        b, sb = sc2.render(c, capi.render_params(L, width=w, height=h, spp=spp, max_bounces=depth))
        assert sa.paths == so[0] == w * h * spp and sa.extend_rays == sb.extend_rays
        assert mean_rel_err(a, r) <= IMAGE_TOL or np.abs(r).max() == 0
        assert mean_rel_err(a, b) <= 1e-5 or np.abs(b).max() == 0
    sc2.close(); sc1.close(); sc0.close()


def test_primary_hit_feature_buffers(L, s1, s1_dev, s1_orc):
    """rtb_render_aovs: albedo / normal / depth / triangle index of the pixel-centre primary hits, against the oracle's
    hits and the scene arrays"""
    w, h = 96, 64
    cam = s1.camera(w / h)
    al, no, de, pr = s1_dev.render_aovs(cam, w, h)
    rays = L.primary_rays(cam, w, h)
    ref = s1_orc.trace_closest(rays, capi.HIT_DTYPE)
    assert (pr.ravel() == ref["prim"]).all()
    assert (de.ravel().view(np.uint32) == ref["t"].view(np.uint32)).all()
    arr = s1.arrays()
    hit = ref["prim"] >= 0
    V = arr["vertices"].reshape(-1, 3, 3)[ref["prim"][hit]].astype(np.float64)
    n = -np.cross(V[:, 0] - V[:, 1], V[:, 2] - V[:, 0])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    flip = (n * rays["dir"][hit]).sum(1) > 0
    n[flip] *= -1
    assert np.abs(no.reshape(-1, 3)[hit] - n).max() < 1e-5
    assert (no.reshape(-1, 3)[~hit] == 0).all() and (al.reshape(-1, 3)[~hit] == 0).all()
    mats = np.array([[0.65, 0.05, 0.05], [0.12, 0.45, 0.15], [0.73, 0.73, 0.73], [0.62, 0.57, 0.54]], np.float32)  # main.cu:42-45
    assert np.allclose(al.reshape(-1, 3)[hit], mats[arr["material_ids"][ref["prim"][hit]]])


def test_sample_pass_sharding_is_additive(L, s1, s1_dev):
    """the multi-GPU decomposition: samples [0,a) + [a,a+b) == samples [0,a+b) (per-pixel RNG keyed by sample index)"""
    cam = s1.camera(1.0)
    w = h = 64
    full = capi.render_params(L, width=w, height=h, spp=6, max_bounces=6, total_spp=6)
    a = capi.render_params(L, width=w, height=h, spp=2, max_bounces=6, total_spp=6, first_sample=0)
    b = capi.render_params(L, width=w, height=h, spp=4, max_bounces=6, total_spp=6, first_sample=2)
    img_full, _ = s1_dev.render(cam, full)
    ia, _ = s1_dev.render(cam, a)
    ib, _ = s1_dev.render(cam, b)
    # images are sqrt(sum/total): sums add
    assert mean_rel_err(ia.astype(np.float64) ** 2 + ib.astype(np.float64) ** 2, img_full.astype(np.float64) ** 2) <= 1e-5


def test_pool_size_does_not_change_the_image(L, s1, s1_dev):
    cam = s1.camera(1.0)
    p1 = capi.render_params(L, width=80, height=60, spp=4, max_bounces=8, pool_size=1 << 10)
    p2 = capi.render_params(L, width=80, height=60, spp=4, max_bounces=8, pool_size=1 << 15)
    a, sa = s1_dev.render(cam, p1)
    b, sb = s1_dev.render(cam, p2)
    assert sa.extend_rays == sb.extend_rays and sa.shadow_rays == sb.shadow_rays
    assert sa.iterations > sb.iterations
    assert mean_rel_err(a, b) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("env", [
    {"RTB_POOLED": "0", "RTB_PIPELINES": "1", "RTB_FUSED": "0", "RTB_TRI_STEP": "0"},  # each lane walks all its triangles, extend and shadow launches apart
    {"RTB_POOLED": "0", "RTB_PIPELINES": "2", "RTB_FUSED": "1", "RTB_TRI_STEP": "1"},  # one triangle per lane and step, the rest carried over
    {"RTB_POOLED": "1", "RTB_PIPELINES": "1", "RTB_FUSED": "1"},  # pooled triangle tests (shared memory), one trace launch
    {"RTB_POOLED": "1", "RTB_PIPELINES": "2", "RTB_FUSED": "1", "RTB_CHUNK": "32", "RTB_REFILL": "32"},
    {"RTB_POOLED": "0", "RTB_PIPELINES": "2", "RTB_FUSED": "1", "RTB_PREFETCH": "0", "RTB_REFILL": "1"},
])
def test_every_traversal_kernel_variant_matches_the_oracle(gpu, oracle, bunny, monkeypatch, env):
    """the tuning knobs CudaBackend reads when a context is created select different kernels / schedules;
    all of them must give the oracle's image and ray counts (hits are bit-exact, so the counts are equal
    up to the ulp-level differences of the shading arithmetic)"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    hs = gpu.host_scene(capi.RTB_SCENE_S1, *bunny)
    ctx2 = gpu.context(0)
    sc = ctx2.scene(hs.desc)
    cam = hs.camera(1.0)
    img, st, ref, ost = render_pair(gpu, sc, oracle.scene(hs.desc), cam, width=320, height=320, spp=4, max_bounces=10)
    assert st.pipelines == int(env["RTB_PIPELINES"]) and st.fused_trace == int(env["RTB_FUSED"])
    assert st.paths == ost[0]
    assert abs(int(st.extend_rays) - int(ost[1])) <= 2e-3 * ost[1]
    assert abs(int(st.shadow_rays) - int(ost[2])) <= 2e-3 * ost[2]
    assert mean_rel_err(img, ref) <= IMAGE_TOL
    sc.close()


def test_unknown_material_type_is_rejected(L, ctx):
    verts, mat, lid = small_scene_arrays(n=5)
    ms = std_materials()
    ms[1].type = 7
    desc, keep = make_desc(verts, mat, np.full(len(mat), -1, np.int32), ms, [])
    h = C.c_void_p()
    assert L.lib.rtb_scene_create(ctx.h, C.byref(desc), None, C.byref(h)) == -1
    assert b"material type" in L.lib.rtb_last_error()


def test_bad_arguments_return_status_codes(L, ctx, s1_dev):
    lib = L.lib
    assert lib.rtb_scene_create(ctx.h, None, None, None) == -1
    p = capi.render_params(L, width=0, height=10)
    out = np.zeros(30, np.float32)
    assert lib.rtb_render(s1_dev.h, C.byref(capi.Camera()), C.byref(p), out.ctypes.data_as(C.c_void_p), None) == -1
    assert b"bad" in lib.rtb_last_error()
    verts, mat, lid = small_scene_arrays(n=5)
    mat = mat.copy(); mat[0] = 99
    desc, keep = make_desc(verts, mat, lid, std_materials(), [])
    h = C.c_void_p()
    assert lib.rtb_scene_create(ctx.h, C.byref(desc), None, C.byref(h)) == -1
    assert b"material id" in lib.rtb_last_error()


@pytest.mark.parametrize("bad", [np.inf, -np.inf, np.nan, 3e38, -2e30])
def test_non_finite_or_huge_vertices_are_rejected_not_hung(L, ctx, bad):
    """one infinite vertex used to leave PLOC without a mutual pair for ever (the build never returned).
    The builder now checks every vertex on the device and returns RTB_ERR_INVALID; 1e25 still builds."""
    verts, mat, lid = small_scene_arrays(n=50)
    v = verts.copy(); v[17, 4] = bad
    desc, keep = make_desc(v, mat, np.full(len(mat), -1, np.int32), std_materials(), [])
    h = C.c_void_p()
    assert L.lib.rtb_scene_create(ctx.h, C.byref(desc), None, C.byref(h)) == -1
    assert b"not finite" in L.lib.rtb_last_error()
    v[17, 4] = 1e25
    desc, keep = make_desc(v, mat, np.full(len(mat), -1, np.int32), std_materials(), [])
    assert L.lib.rtb_scene_create(ctx.h, C.byref(desc), None, C.byref(h)) == 0
    L.lib.rtb_scene_destroy(h)
