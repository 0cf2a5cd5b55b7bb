"""Unit known-answer tests of the device functions on the render path ("unit KATs for
ray/triangle, ray/box, offset_ray_origin, BSDF/light sampling against the host oracle").

rtb_kat_eval runs ONE device function per record on the GPU (`-m gpu`) or in the host build of the
same bodies (tests/emu); the oracle's orc_* helpers are the reference-following restatements
(triangle.cuh:39-58, utility.cuh:31-47, material.cuh:60-109, light.cuh:38-46).  Bars: bit-exact
where only IEEE + - * / sqrt fma are involved (triangle test, origin offset, RNG, area-light
sample); 4 ulp-scale relative tolerance where sincosf / powf differ between libm and the device
(BSDF samples); the quantised slab test must be CONSERVATIVE against an exact (float64) slab test.
"""
import ctypes as C

import numpy as np
import pytest

from rtcuda_b200 import capi


@pytest.fixture(scope="module", params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def L(request):
    return request.getfixturevalue(request.param)


@pytest.fixture(scope="module")
def ctx(L):
    return L.context(0)


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def unit(v):
    v = np.asarray(v, np.float32)
    return (v / np.linalg.norm(v, axis=-1, keepdims=True)).astype(np.float32)


def test_triangle_intersection_bit_exact(ctx, oracle):
    rng = np.random.default_rng(11)
    n = 4000
    tri = rng.random((n, 9)).astype(np.float32)
    o = (rng.random((n, 3)).astype(np.float32) - np.float32(0.5)) * np.float32(3)
    # aim at a point of the triangle's plane neighbourhood so that about half of the rays hit
    bary = rng.random((n, 2)).astype(np.float32) * np.float32(1.4) - np.float32(0.2)
    p0, p1, p2 = tri[:, 0:3], tri[:, 3:6], tri[:, 6:9]
    target = p0 + bary[:, :1] * (p1 - p0) + bary[:, 1:] * (p2 - p0)
    d = unit(target - o)
    tmax = np.where(rng.random(n) < 0.2, np.float32(0.5), np.float32(3e38)).astype(np.float32)
    # hand-made edge cases: ray through a vertex, along an edge, parallel to the plane, behind the origin, tmax == t
    tri[0] = [0, 0, 0, 1, 0, 0, 0, 1, 0]; o[0] = [0, 0, 1]; d[0] = [0, 0, -1]; tmax[0] = 3e38        # through p0
    tri[1] = tri[0]; o[1] = [0.5, 0, 1]; d[1] = [0, 0, -1]; tmax[1] = 3e38                            # on the edge p0-p1
    tri[2] = tri[0]; o[2] = [0.2, 0.2, 1]; d[2] = [1, 0, 0]; tmax[2] = 3e38                           # parallel: 1/0
    tri[3] = tri[0]; o[3] = [0.2, 0.2, -1]; d[3] = [0, 0, -1]; tmax[3] = 3e38                         # behind
    tri[4] = tri[0]; o[4] = [0.2, 0.2, 1]; d[4] = [0, 0, -1]; tmax[4] = 1.0                           # t == tmax: inclusive
    tri[5] = tri[0]; o[5] = [0.2, 0.2, 1]; d[5] = [0, 0, -1]; tmax[5] = np.nextafter(np.float32(1), np.float32(0))  # just short
    tri[6] = [0, 0, 0, 0, 0, 0, 0, 0, 0]; o[6] = [0, 0, 1]; d[6] = [0, 0, -1]; tmax[6] = 3e38        # degenerate: n = 0
    rec = np.concatenate([tri, o, d, tmax[:, None]], axis=1)
    out = ctx.kat(capi.RTB_KAT_TRI_INTERSECT, rec)
    hits = 0
    for i in range(n):
        ray = np.zeros(1, capi.RAY_DTYPE); ray["origin"] = o[i]; ray["dir"] = d[i]; ray["tmax"] = tmax[i]
        hit, tuv = oracle.tri_intersect(tri[i], ray)
        assert bool(out[i, 0]) == hit, i
        if hit:
            hits += 1
            assert (bits(out[i, 1:4]) == bits(tuv)).all(), (i, out[i], tuv)
    assert 0.1 * n < hits < 0.75 * n
    assert out[0, 0] == 1 and out[1, 0] == 1 and out[2, 0] == 0 and out[3, 0] == 0 and out[4, 0] == 1 and out[5, 0] == 0 and out[6, 0] == 0


def test_offset_ray_origin_bit_exact(ctx, oracle):
    rng = np.random.default_rng(5)
    n = 3000
    p = ((rng.random((n, 3)) - 0.5) * 4).astype(np.float32)
    p[: n // 3] *= np.float32(0.01)  # below 1/32: the float-offset branch (utility.cuh:43-45)
    p[0] = [0, 0, 0]; p[1] = [1 / 32, -1 / 32, 0.03125]; p[2] = [-0.0, 1e-30, -1e-30]
    nrm = unit(rng.normal(size=(n, 3)))
    nrm[3] = [1, 0, 0]; nrm[4] = [0, -1, 0]
    out = ctx.kat(capi.RTB_KAT_OFFSET_ORIGIN, np.concatenate([p, nrm], axis=1))
    for i in range(n):
        assert (bits(out[i]) == bits(oracle.offset_ray_origin(p[i], nrm[i]))).all(), i


def test_counter_rng_bit_exact_and_in_range(ctx, oracle):
    rng = np.random.default_rng(9)
    n = 5000
    keys = rng.integers(0, 2 ** 32, size=(n, 4), dtype=np.uint64).astype(np.uint32)
    keys[0] = [1, 0, 0, 0]; keys[1] = [1, 2 ** 31 - 1, 2 ** 24 - 1, 0x80000000 + 255]; keys[2] = [0xffffffff] * 4
    out = ctx.kat(capi.RTB_KAT_RAND4, keys.view(np.float32))
    assert (out > 0).all() and (out <= 1).all()  # (0, 1] like curand_uniform
    for i in range(0, n, 7):
        assert (bits(out[i]) == bits(oracle.rand4(*[int(k) for k in keys[i]]))).all(), i
    assert abs(float(out.mean()) - 0.5) < 0.01 and abs(float(out.var()) - 1 / 12) < 0.005


@pytest.mark.parametrize("mtype", [capi.RTB_MATTE, capi.RTB_MIRROR, capi.RTB_GLASS, capi.RTB_GLOSSY])
def test_bsdf_sampling_matches_the_oracle(L, ctx, oracle, mtype):
    rng = np.random.default_rng(100 + mtype)
    n = 1500
    albedo = rng.random((n, 3)).astype(np.float32)
    ior = (np.float32(1.1) + rng.random(n).astype(np.float32)) if mtype == capi.RTB_GLASS else (np.float32(5) + 200 * rng.random(n).astype(np.float32))
    wo, nrm = unit(rng.normal(size=(n, 3))), unit(rng.normal(size=(n, 3)))
    u = rng.random((n, 2)).astype(np.float32)
    rec = np.zeros((n, 16), np.float32)
    rec[:, 0:3] = albedo; rec[:, 3] = ior; rec[:, 4] = np.full(n, mtype, np.int32).view(np.float32)
    rec[:, 5:8] = wo; rec[:, 8:11] = nrm; rec[:, 11:13] = u
    out = ctx.kat(capi.RTB_KAT_SAMPLE_F, rec)
    exact = L is not None and mtype == capi.RTB_MIRROR  # no libm call on that branch
    worst = 0.0
    for i in range(n):
        m = capi.Material(); m.albedo[0], m.albedo[1], m.albedo[2] = albedo[i]; m.ior = float(ior[i]); m.type = mtype
        f, no, wi, pdf = oracle.sample_f(m, wo[i], nrm[i], float(u[i, 0]), float(u[i, 1]))
        ref = np.concatenate([f, no, wi, [pdf]]).astype(np.float32)
        got = out[i, :10]
        if exact:
            assert (bits(got) == bits(ref)).all(), i
        else:
            scale = np.maximum(np.abs(ref), 1e-3)
            ok = np.isfinite(ref)
            worst = max(worst, float(np.max(np.abs(got[ok] - ref[ok]) / scale[ok])))
    # sincosf / powf: device and libm agree to a few ulp; a glass sample next to the Fresnel threshold may flip branch
    assert worst < (1e-4 if mtype != capi.RTB_GLASS else 5e-3), worst
    # invariants of material.cuh:60-109: the returned normal faces wi (reflection side) except after refraction
    wi, nn = out[:, 6:9], out[:, 3:6]
    if mtype in (capi.RTB_MATTE, capi.RTB_MIRROR):
        assert (np.einsum("ij,ij->i", wi, nn) > -1e-6).all()
        assert np.allclose(np.linalg.norm(wi, axis=1), 1, atol=1e-5)
    if mtype == capi.RTB_MATTE:  # pdf = cos / pi
        assert np.allclose(out[:, 9], np.einsum("ij,ij->i", wi, nn) / np.pi, rtol=1e-5, atol=1e-7)


def test_area_light_sampling_bit_exact(ctx, oracle):
    rng = np.random.default_rng(21)
    n = 2000
    tri = rng.random((n, 9)).astype(np.float32)
    p = (rng.random((n, 3)).astype(np.float32) - np.float32(0.5)) * np.float32(3)
    u = rng.random((n, 2)).astype(np.float32)
    u[0] = [1.0, 1.0]; u[1] = [np.float32(2 ** -24), 1.0]  # the ends of (0, 1]
    rec = np.zeros((n, 16), np.float32)
    rec[:, :9] = tri; rec[:, 9:12] = p; rec[:, 12:14] = u
    out = ctx.kat(capi.RTB_KAT_SAMPLE_LI, rec)
    for i in range(n):
        wi, t, pdf = oracle.sample_li_area(tri[i], p[i], float(u[i, 0]), float(u[i, 1]))
        assert (bits(out[i, :3]) == bits(wi)).all() and bits(out[i, 3]) == bits(t) and bits(out[i, 4]) == bits(pdf), i
    # the sampled point lies on the triangle: P + t * wi has barycentrics in [0, 1]
    q = p + out[:, 3:4] * out[:, :3]
    p0, e1, e2 = tri[:, :3], tri[:, 3:6] - tri[:, :3], tri[:, 6:9] - tri[:, :3]
    A = np.stack([e1, e2], axis=2).astype(np.float64)
    for i in range(0, n, 50):
        b, *_ = np.linalg.lstsq(A[i], (q[i] - p0[i]).astype(np.float64), rcond=None)
        assert b.min() > -1e-4 and b.sum() < 1 + 1e-4


def test_quantised_slab_test_is_conservative(ctx):
    """aabb_intersector.cuh:14-36 accepts iff entry <= exit; the 8-wide node tests QUANTISED child boxes against
    [0, tmax] with a pad.  It may accept more, never less: whenever the exact (float64) box meets the ray segment the
    quantised test must pass; and it must still cull: most rays that clearly miss are rejected."""
    rng = np.random.default_rng(33)
    n = 200000
    plo = (rng.random((n, 3)) * 2 - 1).astype(np.float32)
    ext = (10.0 ** rng.uniform(-3, 0.5, size=(n, 3))).astype(np.float32)
    phi = plo + ext
    a, b = rng.random((n, 3)).astype(np.float32), rng.random((n, 3)).astype(np.float32)
    clo = plo + np.minimum(a, b) * ext
    chi = plo + np.maximum(a, b) * ext
    chi = np.minimum(chi, phi); clo = np.maximum(clo, plo)
    o = (rng.random((n, 3)) * 6 - 3).astype(np.float32)
    target = clo + rng.random((n, 3)).astype(np.float32) * (chi - clo) + (rng.normal(size=(n, 3)) * 0.3 * ext * (rng.random((n, 1)) < 0.5)).astype(np.float32)
    d = unit(target - o)
    axis = rng.random(n) < 0.1  # axis-parallel rays: zero direction components (|d| < eps clamp, aabb_intersector.cuh:17-19)
    d[axis, 0] = 0; d[axis] = unit(d[axis] + np.float32([0, 1e-20, 0]))
    tmax = np.where(rng.random(n) < 0.3, rng.random(n) * 4, 3e38).astype(np.float32)
    rec = np.zeros((n, 20), np.float32)
    rec[:, 0:3] = plo; rec[:, 3:6] = phi; rec[:, 6:9] = clo; rec[:, 9:12] = chi; rec[:, 12:15] = o; rec[:, 15:18] = d; rec[:, 18] = tmax
    got = ctx.kat(capi.RTB_KAT_SLAB, rec)[:, 0] > 0
    # exact slab test in float64 on the UNQUANTISED child box
    o64, d64 = o.astype(np.float64), d.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (clo.astype(np.float64) - o64) / d64
        t2 = (chi.astype(np.float64) - o64) / d64
    inside = (o64 >= clo) & (o64 <= chi)
    par = d64 == 0
    tn = np.where(par, np.where(inside, -np.inf, np.inf), np.minimum(t1, t2)).max(axis=1)
    tf = np.where(par, np.where(inside, np.inf, -np.inf), np.maximum(t1, t2)).min(axis=1)
    exact = (np.maximum(tn, 0) <= np.minimum(tf, tmax.astype(np.float64)))
    assert 0.2 < exact.mean() < 0.9
    missed = exact & ~got
    assert not missed.any(), f"{missed.sum()} boxes culled that the exact test hits, first {np.nonzero(missed)[0][:5]}"
    # looseness: of the boxes the exact test rejects by a clear margin, few pass
    clear = (np.maximum(tn, 0) > np.minimum(tf, tmax.astype(np.float64)) + 0.05 * np.linalg.norm(ext, axis=1) + 1e-3)
    assert (got & clear).sum() <= 0.02 * clear.sum()
