// rtb_core.h — arithmetic shared by every kernel: explicit-rounding float ops,
// 3-vectors, the counter-based RNG and the reference's triangle test.
//
// Everything here is RTB_HD (host + device).  The device build is the
// product; the host build exists only so tests/emu can single-step the very
// same kernel bodies on a machine without a GPU.  All float arithmetic is
// spelled with explicit IEEE operations (no compiler contraction) so that the
// two builds — and the CPU oracle — agree bit for bit wherever libm is not
// involved.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <float.h>

#if defined(__CUDACC__)
#define RTB_HD __host__ __device__ __forceinline__
#define RTB_HD_NOINLINE __host__ __device__
#else
#define RTB_HD inline
#define RTB_HD_NOINLINE
#endif

namespace rtb {

// ---------------------------------------------------------------- scalar ops
#if defined(__CUDA_ARCH__)
RTB_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
RTB_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
RTB_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
RTB_HD float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
RTB_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
RTB_HD float frcp(float a) { return __frcp_rn(a); }  // IEEE 1/x: the sequence nvcc emits for the reference's `1.f / x`
RTB_HD float fsqrt(float a) { return __fsqrt_rn(a); }
// approximate 1/x (MUFU.RCP, <= 1 ulp): only for quantities that feed CONSERVATIVE tests (the slab
// test pads its interval by 16 ulp), never for anything that decides a hit
RTB_HD float frcp_fast(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
RTB_HD int f2i(float a) { return __float_as_int(a); }
RTB_HD float i2f(int a) { return __int_as_float(a); }
RTB_HD uint32_t f2u(float a) { return __float_as_uint(a); }
RTB_HD float u2f(uint32_t a) { return __uint_as_float(a); }
RTB_HD int popc(uint32_t a) { return __popc(a); }
RTB_HD int bfind(uint32_t a) { return 31 - __clz(a); }  // index of highest set bit, -1 if none
#else
// host: compiled with -ffp-contract=off, so + - * / are single IEEE operations
RTB_HD float fadd(float a, float b) { return a + b; }
RTB_HD float fsub(float a, float b) { return a - b; }
RTB_HD float fmul(float a, float b) { return a * b; }
RTB_HD float ffma(float a, float b, float c) { return fmaf(a, b, c); }
RTB_HD float fdiv(float a, float b) { return a / b; }
RTB_HD float frcp(float a) { return 1.0f / a; }
RTB_HD float fsqrt(float a) { return sqrtf(a); }
RTB_HD float frcp_fast(float a) { return 1.0f / a; }
RTB_HD int f2i(float a) { int r; memcpy(&r, &a, 4); return r; }
RTB_HD float i2f(int a) { float r; memcpy(&r, &a, 4); return r; }
RTB_HD uint32_t f2u(float a) { uint32_t r; memcpy(&r, &a, 4); return r; }
RTB_HD float u2f(uint32_t a) { float r; memcpy(&r, &a, 4); return r; }
RTB_HD int popc(uint32_t a) { return __builtin_popcount(a); }
RTB_HD int bfind(uint32_t a) { return a ? 31 - __builtin_clz(a) : -1; }
#endif

constexpr float kPi = 3.14159265358979323846f;      // constant.hpp:4
constexpr float kTwoPi = 6.28318530717958647692f;   // constant.hpp:5
constexpr float kInvPi = 0.31830988618379067153f;   // constant.hpp:6

// 16-byte word: every node / triangle / ray fetch is one 128-bit load
struct alignas(16) Q4 {
    uint32_t x, y, z, w;
};
struct alignas(16) F4 {
    float x, y, z, w;
};
#if defined(__CUDA_ARCH__)
RTB_HD Q4 ldg(const Q4 *p) { uint4 v = __ldg(reinterpret_cast<const uint4 *>(p)); Q4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
RTB_HD F4 ldg(const F4 *p) { float4 v = __ldg(reinterpret_cast<const float4 *>(p)); F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
#else
RTB_HD Q4 ldg(const Q4 *p) { return *p; }
RTB_HD F4 ldg(const F4 *p) { return *p; }
#endif

// ---------------------------------------------------------------- 3-vector
struct V3 {
    float x, y, z;
};
RTB_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RTB_HD V3 v3(float s) { return v3(s, s, s); }
RTB_HD V3 vneg(V3 a) { return v3(-a.x, -a.y, -a.z); }
RTB_HD V3 vadd(V3 a, V3 b) { return v3(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
RTB_HD V3 vsub(V3 a, V3 b) { return v3(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
RTB_HD V3 vmul(V3 a, V3 b) { return v3(fmul(a.x, b.x), fmul(a.y, b.y), fmul(a.z, b.z)); }
RTB_HD V3 vscale(V3 a, float s) { return v3(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)); }
// a + s*b  (one fma per component)
RTB_HD V3 vmad(V3 a, float s, V3 b) { return v3(ffma(s, b.x, a.x), ffma(s, b.y, a.y), ffma(s, b.z, a.z)); }
// dot and cross use the contraction pattern nvcc emits for the reference's
// vec3.cuh:61-69 (SASS of kernel ch, sm_100a): the triangle test must round
// exactly like the reference to pick the same triangle near shared edges.
RTB_HD float vdot(V3 a, V3 b) { return ffma(a.z, b.z, ffma(a.x, b.x, fmul(a.y, b.y))); }
RTB_HD V3 vcross(V3 a, V3 b) {
    return v3(ffma(a.y, b.z, -fmul(a.z, b.y)), ffma(a.z, b.x, -fmul(a.x, b.z)),
              ffma(a.x, b.y, -fmul(a.y, b.x)));
}
// cross without contraction: the reference builds Triangle::n on the HOST
// (triangle.cuh:7 called from main.cu:80), where g++ does not fuse.
RTB_HD V3 vcross_nofma(V3 a, V3 b) {
    return v3(fsub(fmul(a.y, b.z), fmul(a.z, b.y)), fsub(fmul(a.z, b.x), fmul(a.x, b.z)),
              fsub(fmul(a.x, b.y), fmul(a.y, b.x)));
}
RTB_HD float vlen2(V3 a) { return vdot(a, a); }
RTB_HD float vlen(V3 a) { return fsqrt(vlen2(a)); }
RTB_HD V3 vnormalize(V3 a) { return vscale(a, frcp(vlen(a))); }  // Vec3::unit_vector, vec3.cuh:129-132
RTB_HD float vmax(V3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }

// ---------------------------------------------------------------- RNG
// Counter-based generator keyed by (seed, pixel, sample, dimension block):
// PCG-4D (Jarzynski & Olano 2020).  It replaces the reference's per-slot
// cuRAND XORWOW state (render.cuh:68-73) whose streams depend on the slot
// history and cannot be reproduced outside the reference.
// One call yields four 32-bit words.
struct U4 {
    uint32_t x, y, z, w;
};
RTB_HD U4 pcg4d(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    U4 v;
    v.x = a * 1664525u + 1013904223u;
    v.y = b * 1664525u + 1013904223u;
    v.z = c * 1664525u + 1013904223u;
    v.w = d * 1664525u + 1013904223u;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    v.x ^= v.x >> 16; v.y ^= v.y >> 16; v.z ^= v.z >> 16; v.w ^= v.w >> 16;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    return v;
}
// uniform in (0,1] like curand_uniform (24 random bits)
RTB_HD float u01(uint32_t bits) { return fmul((float)((bits >> 8) + 1u), 5.9604644775390625e-08f); }

// Dimension blocks of one path: block 0 = camera jitter; for bounce b
// (1-based shade index) block 2b-1 = Russian roulette, 2b = BSDF + light.
struct Rand4 {
    float a, b, c, d;
};
RTB_HD Rand4 rand4(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t block) {
    U4 r = pcg4d(pixel, sample, block, seed);
    Rand4 o;
    o.a = u01(r.x); o.b = u01(r.y); o.c = u01(r.z); o.d = u01(r.w);
    return o;
}

// ---------------------------------------------------------------- triangle
// 48-byte record {p0, e1 = p0-p1, e2 = p2-p0, n = cross(e1,e2)}: the
// reference's Triangle (triangle.cuh:4-21), stored as three 16-byte words so
// one test is three 128-bit loads.
struct Tri48 {
    float p0x, p0y, p0z, e1x;
    float e1y, e1z, e2x, e2y;
    float e2z, nx, ny, nz;
};
RTB_HD V3 tri_p0(const Tri48 &t) { return v3(t.p0x, t.p0y, t.p0z); }
RTB_HD V3 tri_e1(const Tri48 &t) { return v3(t.e1x, t.e1y, t.e1z); }
RTB_HD V3 tri_e2(const Tri48 &t) { return v3(t.e2x, t.e2y, t.e2z); }
RTB_HD V3 tri_n(const Tri48 &t) { return v3(t.nx, t.ny, t.nz); }

RTB_HD Tri48 tri_from_vertices(V3 p0, V3 p1, V3 p2) {
    V3 e1 = vsub(p0, p1), e2 = vsub(p2, p0), n = vcross_nofma(e1, e2);
    Tri48 t;
    t.p0x = p0.x; t.p0y = p0.y; t.p0z = p0.z;
    t.e1x = e1.x; t.e1y = e1.y; t.e1z = e1.z;
    t.e2x = e2.x; t.e2y = e2.y; t.e2z = e2.z;
    t.nx = n.x; t.ny = n.y; t.nz = n.z;
    return t;
}

// Triangle::intersect, triangle.cuh:39-58, operation for operation as nvcc
// compiles it for sm_100a.  Accepts iff u>=0, v>=0, u+v<=1 and 0 < t <= tmax.
// tri_candidate is the part that does not depend on tmax (the barycentric
// test and t); it returns t, or -1 when the barycentric test rejects, so that
// a warp can pool the candidates of all its rays and leave the order-dependent
// accept rule `0 < t <= tmax` to the ray's own lane.
RTB_HD float tri_candidate(const Tri48 &tr, V3 o, V3 d, float &u_out, float &v_out) {
    V3 c = vsub(tri_p0(tr), o);
    V3 r = vcross(d, c);
    V3 n = tri_n(tr);
    float inv_det = frcp(vdot(d, n));
    float u = fmul(inv_det, vdot(tri_e2(tr), r));
    float v = fmul(inv_det, vdot(tri_e1(tr), r));
    u_out = u; v_out = v;
    if (u >= 0.0f && v >= 0.0f && fadd(u, v) <= 1.0f) return fmul(inv_det, vdot(c, n));
    return -1.0f;
}
RTB_HD bool tri_intersect(const Tri48 &tr, V3 o, V3 d, float tmax, float &t_out, float &u_out,
                          float &v_out) {
    float u, v;
    const float t = tri_candidate(tr, o, d, u, v);
    if (0.0f < t && t <= tmax) {
        t_out = t; u_out = u; v_out = v;
        return true;
    }
    return false;
}

// Wächter–Binder origin offset, utility.cuh:31-47
RTB_HD V3 offset_ray_origin(V3 p, V3 n) {
    const float int_scale = 256.f, float_scale = 1.f / 65536.f, origin = 1.f / 32.f;
    int ox = (int)fmul(int_scale, n.x), oy = (int)fmul(int_scale, n.y), oz = (int)fmul(int_scale, n.z);
    float px = i2f(f2i(p.x) + (p.x < 0 ? -ox : ox));
    float py = i2f(f2i(p.y) + (p.y < 0 ? -oy : oy));
    float pz = i2f(f2i(p.z) + (p.z < 0 ? -oz : oz));
    return v3(fabsf(p.x) < origin ? ffma(float_scale, n.x, p.x) : px,
              fabsf(p.y) < origin ? ffma(float_scale, n.y, p.y) : py,
              fabsf(p.z) < origin ? ffma(float_scale, n.z, p.z) : pz);
}

}  // namespace rtb
