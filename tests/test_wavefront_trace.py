"""The RENDER path's traversal kernel pinned directly.

rtb_trace_closest / rtb_trace_any run one thread per ray.  The renderer does not: its rays are traversed by the
persistent kernel k_trace (dynamic fetch from the extend / shadow queues, refill of idle lanes, stepped or pooled
triangle tests, finish words staged by cp.async, hit records appended per material type) — the replacement of
kernels ch / ah (render.cuh:278-328) around Bvh::traverse (bvh.cuh:251-357).  rtb_trace_wavefront loads caller rays into
those queues, runs one iteration's trace launch and reads the hit records / splats back, so these tests compare
the kernel that renders, in every schedule, bit for bit with
  - the reference's own outputs (tests/golden/s1_hits.npz, s1_any.npz, s2_hits.npz, s2_any.npz),
  - the oracle,
  - the reference run live on the same B200 (oracle/_ref/ref_harness), when it was built.
The host build (tests/emu) runs the same queue plumbing with its sequential extend / shadow bodies.
"""
import os

import numpy as np
import pytest

from rtcuda_b200 import capi
from conftest import GOLDEN, random_rays
from test_parity import assert_hits_equal

# (name, options): every traversal schedule the CUDA backend has.  Defaults: tri_step 2, fused, not pooled.
SCHEDULES = [
    ("default", {}),
    ("tri_step0", {"tri_step": 0}),
    ("tri_step1", {"tri_step": 1}),
    ("tri_step3", {"tri_step": 3}),
    ("pooled", {"pooled": 1}),
    ("smem_stack", {"smem_stack": 1}),
    ("unfused", {"fused": 0}),
    ("unfused_pooled", {"fused": 0, "pooled": 1}),
    ("refill1_chunk32_noprefetch", {"refill": 1, "chunk": 32, "prefetch": 0}),
    ("refill32", {"refill": 32, "chunk": 4096}),
    ("one_block_per_sm", {"trace_blocks": 1}),
]


@pytest.fixture(scope="module", params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def L(request):
    return request.getfixturevalue(request.param)


@pytest.fixture(scope="module")
def s1(L, bunny):
    return L.host_scene(capi.RTB_SCENE_S1, *bunny)


@pytest.fixture(scope="module")
def s1_dev(L, s1):
    return L.context(0).scene(s1.desc)


@pytest.fixture(scope="module")
def s1_orc(oracle, s1):
    return oracle.scene(s1.desc)


def golden(name):
    path = os.path.join(GOLDEN, name)
    assert os.path.exists(path), f"missing fixture {path}: run tools/make_golden.py on a B200"
    return np.load(path)


def test_wavefront_kernel_equals_the_reference_fixtures(s1_dev):
    g, a = golden("s1_hits.npz"), golden("s1_any.npz")
    hits, occ, launches = s1_dev.trace_wavefront(g["rays"], a["rays"], a["excluded"])
    assert launches in (1, 2)
    assert_hits_equal(hits, g["hits"])
    assert (occ == a["occluded"]).all()


def test_wavefront_kernel_equals_the_oracle(L, s1, s1_dev, s1_orc):
    cam = s1.camera(16 / 9)
    rays = np.concatenate([L.primary_rays(cam, 320, 180), random_rays(60000, seed=17)])
    srays = random_rays(50000, seed=23, tmax=np.float32(0.7))
    excl = np.random.default_rng(3).integers(-1, s1.desc.num_triangles, len(srays)).astype(np.int32)
    hits, occ, _ = s1_dev.trace_wavefront(rays, srays, excl)
    assert_hits_equal(hits, s1_orc.trace_closest(rays, capi.HIT_DTYPE))
    ref_occ = s1_orc.trace_any(srays, excl)
    assert 0.05 < ref_occ.mean() < 0.95
    assert (occ == ref_occ).all()
    # either queue alone
    h2, o2, _ = s1_dev.trace_wavefront(rays[:1000])
    assert o2 is None and (h2 == hits[:1000]).all()
    h3, o3, _ = s1_dev.trace_wavefront(None, srays[:1000], excl[:1000])
    assert h3 is None and (o3 == occ[:1000]).all()
    h4, o4, n4 = s1_dev.trace_wavefront()
    assert h4 is None and o4 is None and n4 == 0


def test_wavefront_kernel_agrees_with_the_per_ray_entry_on_edge_cases(s1_dev):
    rays = random_rays(4096, seed=5)
    rays["dir"][::7] = [0, 0, -1]; rays["dir"][1::7] = [1, 0, 0]                      # axis-parallel
    rays["origin"][2::7] = [0.5, 0.5, 5.0]; rays["dir"][2::7] = [0, 0, 1]              # leaves the scene: miss
    rays["origin"][3] = [np.nan, 0, 0]; rays["dir"][4] = [np.inf, 0, 0]; rays["dir"][5] = [0, 0, 0]  # non-finite / null
    hits, _, _ = s1_dev.trace_wavefront(rays)
    ref = s1_dev.trace_closest(rays)
    keep = np.ones(len(rays), bool); keep[5] = False  # a null direction: the per-ray entry and the queue agree on a miss, t is moot
    assert (hits["prim"][keep] == ref["prim"][keep]).all() and (hits["t"][keep].view(np.uint32) == ref["t"][keep].view(np.uint32)).all()
    assert hits["prim"][3] == -1 and hits["prim"][4] == -1 and (hits["prim"][2::7] == -1).all()
    sr = random_rays(4096, seed=6, tmax=np.float32(0.4))
    sr["tmax"][::5] = 0.0; sr["tmax"][1::5] = -1.0; sr["origin"][2] = [np.nan, 0, 0]
    _, occ, _ = s1_dev.trace_wavefront(None, sr, None)
    assert (occ == s1_dev.trace_any(sr, None)).all()
    assert occ[0] == 0 and occ[1] == 0 and occ[2] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name,opts", SCHEDULES, ids=[s[0] for s in SCHEDULES])
def test_every_schedule_of_the_persistent_kernel_is_bit_exact(gpu, bunny, name, opts):
    """RTB_TRI_STEP 0/1/2, pooled triangle tests, shared-memory stack, separate launches, fetch parameters: all
    of them must return the reference's hit records and occlusion bits"""
    hs = gpu.host_scene(capi.RTB_SCENE_S1, *bunny)
    ctx = gpu.context(0)
    for k, v in opts.items():
        ctx.set_option(k, v)
        assert ctx.get_option(k) == v
    sc = ctx.scene(hs.desc)
    g, a = golden("s1_hits.npz"), golden("s1_any.npz")
    hits, occ, launches = sc.trace_wavefront(g["rays"], a["rays"], a["excluded"])
    assert launches == (2 if opts.get("fused", 1) == 0 else 1)
    assert_hits_equal(hits, g["hits"])
    assert (occ == a["occluded"]).all()
    sc.close()


@pytest.mark.gpu
def test_more_rays_than_one_queue_holds(gpu, bunny):
    """4.5 M rays: two passes over the 4 Mi-entry queues; the second pass is partial"""
    hs = gpu.host_scene(capi.RTB_SCENE_S1, *bunny)
    sc = gpu.context(0).scene(hs.desc)
    rays = random_rays(4_500_000, seed=77)
    hits, _, launches = sc.trace_wavefront(rays)
    assert launches == 2
    ref = sc.trace_closest(rays)
    assert (hits == ref).all()
    sc.close()


def test_options_are_validated(L):
    ctx = L.context(0)
    assert L.lib.rtb_context_set_option(ctx.h, b"no_such_option", 1) == -1
    assert L.lib.rtb_context_set_option(ctx.h, b"pool", 3) == -1
    ctx.set_option("pool", 1 << 20)
    assert ctx.get_option("pool") == 1 << 20


# ---------------------------------------------------------------- the 10 M-triangle scene (configs C3 / C5)
def s2_rays(L, hs):
    """the rays of tests/golden/s2_hits.npz, regenerated (the fixture stores the hits and a checksum of the rays):
    every 8th pixel centre of the 3840x2160 camera + 100,000 incoherent rays"""
    cam = hs.camera(3840 / 2160)
    prim = L.primary_rays(cam, 3840, 2160).reshape(2160, 3840)[4::8, 4::8].reshape(-1)
    return np.concatenate([prim, random_rays(100000, seed=41)])


def s2_shadow_rays(n, nt, seed=43):
    """shadow rays towards the emitter of the 12x12 field from points in the box; excluded = one of the two light triangles"""
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=capi.RAY_DTYPE)
    rays["origin"] = (rng.random((n, 3)).astype(np.float32) * np.float32(0.96) + np.float32(0.02)) * np.float32([1, 1, -1])
    target = np.array([0.5, 0.999, -0.5], np.float32) + (rng.random((n, 3)).astype(np.float32) - np.float32(0.5)) * np.float32([0.2, 0, 0.2])
    d = target - rays["origin"]
    dist = np.linalg.norm(d, axis=1).astype(np.float32)
    rays["dir"] = (d / dist[:, None]).astype(np.float32)
    rays["tmax"] = dist * np.float32(1.0005)
    excl = rng.integers(nt - 2, nt, n).astype(np.int32)
    return rays, excl


def ray_checksum(rays):
    return int(np.bitwise_xor.reduce(np.ascontiguousarray(rays).view(np.uint32).astype(np.uint64) * np.arange(1, rays.size * 7 + 1, dtype=np.uint64)))


@pytest.mark.gpu
def test_s2_wavefront_kernel_equals_the_reference_fixture(gpu, bunny):
    """C3 / C5 geometry (144 bunnies, 10,000,956 triangles): the reference's Bvh::traverse on its own 10.9 M-node SAH tree
    (35 s host build, tools/make_golden.py) against k_trace on the GPU-built 8-wide tree — ids, t, u, v and the any-hit bits.
    Exact-t ties between two triangles (rays through the shared edge of two shell triangles) are the one legal
    difference: counted and printed."""
    hs = gpu.host_scene(capi.RTB_SCENE_S2, *bunny, grid=12)
    sc = gpu.context(0).scene(hs.desc)
    g, a = golden("s2_hits.npz"), golden("s2_any.npz")
    rays = s2_rays(gpu, hs)
    srays, excl = s2_shadow_rays(len(a["occluded"]), hs.desc.num_triangles)
    assert ray_checksum(rays) == int(g["ray_checksum"]) and ray_checksum(srays) == int(a["ray_checksum"])
    for entry in ("wavefront", "per_ray"):
        if entry == "wavefront":
            hits, occ, _ = sc.trace_wavefront(rays, srays, excl)
        else:
            hits, occ = sc.trace_closest(rays), sc.trace_any(srays, excl)
        ref = g["hits"]
        diff = np.nonzero(hits["prim"] != ref["prim"])[0]
        ties = diff[hits["t"][diff].view(np.uint32) == ref["t"][diff].view(np.uint32)]
        same = hits["prim"] == ref["prim"]
        print(f"S2 {entry}: {len(rays)} rays, {len(diff)} id differences of which {len(ties)} exact-t ties; any-hit mismatches {(occ != a['occluded']).sum()}")
        assert len(diff) == len(ties) and len(diff) <= 2e-4 * len(rays)
        for k in ("t", "u", "v"):
            assert (hits[k][same].view(np.uint32) == ref[k][same].view(np.uint32)).all(), k
        assert (occ == a["occluded"]).all()
    sc.close()
