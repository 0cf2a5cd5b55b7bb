// rtb_bvh8.h — compressed 8-wide BVH: node format, quantisation and the
// per-ray traversal loop (closest-hit and any-hit).
//
// Replaces the reference's 32-byte binary Bvh::Node (bvh.cuh:5-14), the slab
// test AABBIntersector (aabb_intersector.cuh:4-36), the 29-int local stack
// DeviceStack (device_stack.cuh:4-11) and both Bvh::traverse overloads
// (bvh.cuh:251-357).  Layout follows the compressed wide BVH of Ylitie,
// Karras & Laine (HPG 2017): 80-byte nodes = five 128-bit words, child boxes
// quantised to 8 bits per plane relative to the node origin with per-axis
// power-of-two scales, children of one node stored contiguously, triangles of
// one node stored contiguously (<= 24), octant-ordered traversal through the
// slot permutation `slot ^ octinv`.
//
//   word0: origin.xyz (f32) | ex | ey | ez | imask          (biased exponents)
//   word1: child_base (u32) | tri_base (u32) | meta[0..7]
//   word2: qlo_x[0..7] | qlo_y[0..7]
//   word3: qlo_z[0..7] | qhi_x[0..7]
//   word4: qhi_y[0..7] | qhi_z[0..7]
//
//   meta[s] = 0                         empty slot
//           = 0x20 | (24 + s)           inner child in slot s
//           = unary(count) << 5 | off   leaf child: `count` (1..3) triangles
//                                       starting at tri_base + off
#pragma once
#include "rtb_core.h"

namespace rtb {

constexpr int kNodeWords = 5;       // 80 bytes
constexpr float kSahNodeCost = 1.0f;
constexpr float kSahTriCost = 0.3f;

struct Bvh8View {
    const Q4 *nodes;     // kNodeWords per node
    const F4 *tris;      // 3 words per triangle, leaf order
    const int32_t *prim; // leaf order -> caller's triangle index
    int32_t num_nodes;
    int32_t num_tris;
    // two-level scenes (rtb_scene_create_instanced), null / 0 otherwise: the tree rooted at node 0 is then the one
    // over the INSTANCES (its leaf entries index `top_inst`), every mesh has its own tree further up the same node
    // array (object space, indices already absolute), and `inst` holds kInstWords words per instance
    const F4 *inst;
    const int32_t *top_inst;  // leaf order of the top tree -> instance
    int32_t num_inst;
};

// Instance record, 8 x 16 bytes: rows of the world->object matrix (w = translation), rows of the object->world
// matrix, {root node of the mesh's tree, material word (index | type << 24) or -1, first id of the instance in the
// flattened numbering, first caller index of the mesh} as raw ints, one spare word.
constexpr int kInstWords = 8;
struct InstInfo {
    int32_t root, material, flat_first, mesh_first;
};
RTB_HD InstInfo inst_info(const F4 *inst, int i) {
    const F4 w = ldg(inst + (size_t)i * kInstWords + 6);
    InstInfo r;
    r.root = f2i(w.x); r.material = f2i(w.y); r.flat_first = f2i(w.z); r.mesh_first = f2i(w.w);
    return r;
}
RTB_HD V3 xform_point(const F4 &r0, const F4 &r1, const F4 &r2, V3 p) {
    return v3(ffma(r0.z, p.z, ffma(r0.y, p.y, ffma(r0.x, p.x, r0.w))), ffma(r1.z, p.z, ffma(r1.y, p.y, ffma(r1.x, p.x, r1.w))),
              ffma(r2.z, p.z, ffma(r2.y, p.y, ffma(r2.x, p.x, r2.w))));
}
RTB_HD V3 xform_vector(const F4 &r0, const F4 &r1, const F4 &r2, V3 d) {
    return v3(ffma(r0.z, d.z, ffma(r0.y, d.y, fmul(r0.x, d.x))), ffma(r1.z, d.z, ffma(r1.y, d.y, fmul(r1.x, d.x))),
              ffma(r2.z, d.z, ffma(r2.y, d.y, fmul(r2.x, d.x))));
}

struct HitRec {
    float t, u, v;
    int32_t tri;  // leaf-order triangle index, -1 = miss
};

struct TraceCounters {
    uint32_t nodes, tris;
};

RTB_HD Tri48 load_tri(const F4 *tris, int idx) {
    F4 a = ldg(tris + 3 * (size_t)idx), b = ldg(tris + 3 * (size_t)idx + 1), c = ldg(tris + 3 * (size_t)idx + 2);
    Tri48 t;
    t.p0x = a.x; t.p0y = a.y; t.p0z = a.z; t.e1x = a.w;
    t.e1y = b.x; t.e1z = b.y; t.e2x = b.z; t.e2y = b.w;
    t.e2z = c.x; t.nx = c.y; t.ny = c.z; t.nz = c.w;
    return t;
}

// id of a hit in the caller's numbering: the triangle list of a flat scene, the flattened list of an instanced one
RTB_HD int hit_prim(const Bvh8View &B, int tri, int inst) {
    const int p = B.prim[tri];
    if (inst < 0) return p;
    const InstInfo in = inst_info(B.inst, inst);
    return in.flat_first + (p - in.mesh_first);
}
// A triangle of an instance in WORLD space: its vertices through the object->world matrix, then the record the
// reference's host constructor derives from vertices (triangle.cuh:6-7) — what the flattened scene stores, up to the
// rounding of the transform.  inst < 0: the stored record.
RTB_HD Tri48 load_tri_world(const Bvh8View &B, int idx, int inst) {
    const Tri48 t = load_tri(B.tris, idx);
    if (inst < 0) return t;
    const F4 *m = B.inst + (size_t)inst * kInstWords + 3;
    const F4 r0 = ldg(m), r1 = ldg(m + 1), r2 = ldg(m + 2);
    const V3 p0 = tri_p0(t);
    return tri_from_vertices(xform_point(r0, r1, r2, p0), xform_point(r0, r1, r2, vsub(p0, tri_e1(t))),
                             xform_point(r0, r1, r2, vadd(p0, tri_e2(t))));
}

// ------------------------------------------------------------ quantisation
// smallest biased exponent e with 255 * 2^(e-127) >= extent (with margin)
RTB_HD uint32_t quant_exponent(float extent) {
    float f = fmul(fmul(extent, 1.0f / 255.0f), 1.000001f);
    uint32_t b = f2u(f);
    uint32_t e = (b >> 23) & 0xffu;
    if (b & 0x7fffffu) e += 1;
    if (e < 40u) e = 40u;
    if (e > 235u) e = 235u;  // the traversal adds 15 to it (node_frame)
    return e;
}
RTB_HD uint32_t quant_lo(float lo, float p, uint32_t e) {
    float scale = u2f(e << 23), inv = u2f((254u - e) << 23);
    float q = floorf(fmul(fsub(lo, p), inv));
    q = fminf(fmaxf(q, 0.f), 255.f);
    if (q > 0.f && ffma(q, scale, p) > lo) q -= 1.f;
    return (uint32_t)q;
}
RTB_HD uint32_t quant_hi(float hi, float p, uint32_t e) {
    float scale = u2f(e << 23), inv = u2f((254u - e) << 23);
    float q = ceilf(fmul(fsub(hi, p), inv));
    q = fminf(fmaxf(q, 0.f), 255.f);
    if (q < 255.f && ffma(q, scale, p) < hi) q += 1.f;
    return (uint32_t)q;
}

// ------------------------------------------------------------ ray setup
struct RaySetup {
    V3 o, d, idir;
    uint32_t octinv;  // bit a set iff d_a >= 0
};
RTB_HD RaySetup ray_setup(V3 o, V3 d) {
    RaySetup r;
    r.o = o; r.d = d;
    const float tiny = 8.271806125530277e-25f;  // 2^-80: keeps q*2^e*idir finite
    float dx = fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x;
    float dy = fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y;
    float dz = fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z;
    // 1/d only feeds the conservative slab test, whose interval is padded by 16 ulp: the 1-ulp
    // MUFU.RCP result is enough (an IEEE 1/x costs ~10 instructions per axis)
    r.idir = v3(frcp_fast(dx), frcp_fast(dy), frcp_fast(dz));
    r.octinv = (dx >= 0.f ? 1u : 0u) | (dy >= 0.f ? 2u : 0u) | (dz >= 0.f ? 4u : 0u);
    return r;
}

RTB_HD uint32_t byte_of(uint32_t w, int i) { return (w >> (8 * i)) & 0xffu; }

// Slab-test the 8 quantised child boxes of one node against the ray segment
// [0, tmax]; returns the hit mask: inner children in bits 24..31 at position
// 24 + (slot ^ octinv) (so the highest set bit is the nearest child in octant
// order), triangles of leaf children in bits 0..23.
//
// Plane distance along one axis: t(q) = (origin_node + q*2^e - origin_ray) / d
// = q*a + c with a = 2^e/d and c = (origin_node - origin_ray)/d: one
// integer->float conversion (I2F.U8 with a byte selector) and one FMA per
// plane.  (A byte permute into the mantissa instead of the conversion was
// measured 1.4-2x slower on the B200: profiles/r1/r1_variants.md.)
//
// The test is CONSERVATIVE: it never culls a box the exact reference triangle
// test (triangle.cuh:39-58) could still hit inside.  Error of the computed t:
// c carries two roundings and the 1-ulp reciprocal (|dc| <= 2^-22 |c|), a is
// exact up to the reciprocal, the FMA rounds once; so per AXIS
//   |t - t_exact| <= 2^-22 |c_axis| + 2^-21 |t|.
// Both parts are folded into the constant of the FMA — near planes use c - pad,
// far planes c + pad with pad = 3 * 2^-21 (|c| + 256 |a|) (node_frame: twice the
// absolute bound plus twice the relative one, |t| <= |c| + 255 |a|) — so the
// final comparison is plain tn <= tf.  (Until round 2 the relative part sat in
// the comparison, tn <= tf * (1 + 2^-19): one FMUL per child.)  The pad of an axis
// depends on that axis alone: a ray almost parallel to one axis (|c| huge
// there) must not loosen the test on the other two.  (Round 1 first used one
// pad 2^-21 max|c| for all three axes: such rays then passed every slab test
// and walked whole slices of a 10 M-triangle scene; see profiles/r1/README.md.)
//
// Byte -> float.  I2F.U8 runs on the quarter-rate XU pipe: 48 of them per node made XU the busiest pipe of k_trace
// (ncu r1: XU 52 %, FMA 21 %).  RTB_NODE_CVT_HALF (device builds): two bytes at a time are placed under the exponent
// byte 0x64 of a half2 (one PRMT on the ALU pipe: 0x6400 | q = 1024 + q exactly) and widened with HADD2.F32 on the FMA
// pipe; the 1024 is taken out of the constant, t = (1024 + q) a + (c - 1024 a).  The computed a cancels exactly between
// the two terms, so the only new error is the rounding of (c - 1024 a), at most 2^-24 (|c| + 1024 |a|): the pads below
// grow by 2^-13 |a| — a ten-thousandth of one quantisation step.
#if defined(__CUDA_ARCH__) && !defined(RTB_NODE_CVT_I2F)
#define RTB_NODE_CVT_HALF 1
#endif
#if defined(RTB_NODE_REL_PAD)
#define RTB_NODE_TF(tf) fmul(tf, 1.0000019f)
#else
#define RTB_NODE_TF(tf) (tf)
#endif
struct NodeFrame {
    float ax, ay, az;     // 2^e / d
    float nx, ny, nz;     // c lowered: for the entry planes
    float fx, fy, fz;     // c raised: for the exit planes
};
RTB_HD NodeFrame node_frame(const Q4 &n0, const RaySetup &r) {
    NodeFrame f;
    f.ax = fmul(u2f(byte_of(n0.w, 0) << 23), r.idir.x);
    f.ay = fmul(u2f(byte_of(n0.w, 1) << 23), r.idir.y);
    f.az = fmul(u2f(byte_of(n0.w, 2) << 23), r.idir.z);
    const float cx = fmul(fsub(u2f(n0.x), r.o.x), r.idir.x);
    const float cy = fmul(fsub(u2f(n0.y), r.o.y), r.idir.y);
    const float cz = fmul(fsub(u2f(n0.z), r.o.z), r.idir.z);
    // pad = 3 * 2^-21 (|c| + 256 |a|): 2^-21 (|c| + 256 |a|) for the absolute part (twice the bound above, the 256 |a|
    // covering the rounding of c - 1024 a) plus 2^-20 (|c| + 256 |a|) >= 2 * 2^-21 |t| for the relative part, since
    // |t| = |q a + c| <= |c| + 255 |a|.  Round 1 kept the relative part in the comparison (tn <= tf * (1 + 2^-19)): one
    // FMUL per child and node; folded into the constants it costs nothing per child and at most a third of a
    // quantisation step of extra box for a node a thousand times smaller than its distance.
#if defined(RTB_NODE_REL_PAD)
    const float k = 4.76837158203125e-07f;  // 2^-21 (A/B: relative part in the comparison)
#else
    const float k = 1.430511474609375e-06f;  // 3 * 2^-21
#endif
    const float px = fmul(ffma(fabsf(f.ax), 256.f, fabsf(cx)), k);
    const float py = fmul(ffma(fabsf(f.ay), 256.f, fabsf(cy)), k);
    const float pz = fmul(ffma(fabsf(f.az), 256.f, fabsf(cz)), k);
#if defined(RTB_NODE_CVT_HALF)
    const float bx = ffma(-1024.f, f.ax, cx), by = ffma(-1024.f, f.ay, cy), bz = ffma(-1024.f, f.az, cz);  // base = c - 1024 a
#else
    const float bx = cx, by = cy, bz = cz;
#endif
    f.nx = fsub(bx, px); f.fx = fadd(bx, px);
    f.ny = fsub(by, py); f.fy = fadd(by, py);
    f.nz = fsub(bz, pz); f.fz = fadd(bz, pz);
    return f;
}
#if defined(RTB_NODE_CVT_HALF)
// bytes 2h and 2h+1 of w as the half2 {1024 + b(2h), 1024 + b(2h+1)}
template <int H>
__device__ __forceinline__ uint32_t byte_pair_half2(uint32_t w) { return __byte_perm(w, 0x64646464u, H ? 0x4342 : 0x4140); }
// 1024 + byte J of the quadruple whose two pairs are (lo, hi), as a float
template <int J>
__device__ __forceinline__ float biased_byte(uint32_t lo, uint32_t hi) {
    const uint32_t p = (J & 2) ? hi : lo;
    float r;
    if (J & 1) asm("{.reg .b16 l, h; mov.b32 {l, h}, %1; cvt.f32.f16 %0, h;}" : "=f"(r) : "r"(p));
    else asm("{.reg .b16 l, h; mov.b32 {l, h}, %1; cvt.f32.f16 %0, l;}" : "=f"(r) : "r"(p));
    return r;
}
#endif
// The per-child meta decode is done four children at a time in packed bytes
// (bits4 = unary triangle count or 1 for an inner child, pos4 = bit position
// in the hit mask with the octant permutation already applied to inner
// children), after Ylitie et al. 2017, so a hit child costs two byte
// extracts, a shift and an OR.
#if defined(RTB_NODE_CVT_HALF)
struct PlanePairs { uint32_t nxl, nxh, nyl, nyh, nzl, nzh, fxl, fxh, fyl, fyh, fzl, fzh; };
template <int J>
__device__ __forceinline__ uint32_t child_hit_bits(const NodeFrame &f, uint32_t bits4, uint32_t pos4, const PlanePairs &q, float tmax) {
    const float tnx = ffma(biased_byte<J>(q.nxl, q.nxh), f.ax, f.nx);
    const float tny = ffma(biased_byte<J>(q.nyl, q.nyh), f.ay, f.ny);
    const float tnz = ffma(biased_byte<J>(q.nzl, q.nzh), f.az, f.nz);
    const float tfx = ffma(biased_byte<J>(q.fxl, q.fxh), f.ax, f.fx);
    const float tfy = ffma(biased_byte<J>(q.fyl, q.fyh), f.ay, f.fy);
    const float tfz = ffma(biased_byte<J>(q.fzl, q.fzh), f.az, f.fz);
    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.f));
    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
    if (tn <= RTB_NODE_TF(tf)) return byte_of(bits4, J) << byte_of(pos4, J);
    return 0u;
}
#endif
template <int J>
RTB_HD uint32_t child_hit_bits(const NodeFrame &f, uint32_t bits4, uint32_t pos4, uint32_t nx4, uint32_t ny4, uint32_t nz4,
                               uint32_t fx4, uint32_t fy4, uint32_t fz4, float tmax) {
    const float tnx = ffma((float)byte_of(nx4, J), f.ax, f.nx);
    const float tny = ffma((float)byte_of(ny4, J), f.ay, f.ny);
    const float tnz = ffma((float)byte_of(nz4, J), f.az, f.nz);
    const float tfx = ffma((float)byte_of(fx4, J), f.ax, f.fx);
    const float tfy = ffma((float)byte_of(fy4, J), f.ay, f.fy);
    const float tfz = ffma((float)byte_of(fz4, J), f.az, f.fz);
    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.f));
    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
    if (tn <= RTB_NODE_TF(tf)) return byte_of(bits4, J) << byte_of(pos4, J);
    return 0u;
}
// meta4 -> (bits4, pos4) for four children at once
RTB_HD void meta_decode4(uint32_t meta4, uint32_t oct4, uint32_t &bits4, uint32_t &pos4) {
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;  // low five bits in 24..31
    const uint32_t inner_mask4 = (is_inner4 >> 4) * 0xffu;            // 0xff in the bytes of inner children
    pos4 = (meta4 ^ (oct4 & inner_mask4)) & 0x1f1f1f1fu;
    bits4 = (meta4 >> 5) & 0x07070707u;
}

RTB_HD uint32_t node_hitmask(const Q4 &n0, const Q4 &n1, const Q4 &n2, const Q4 &n3, const Q4 &n4,
                             const RaySetup &r, float tmax) {
    const NodeFrame f = node_frame(n0, r);
    const bool px = (r.octinv & 1u) != 0, py = (r.octinv & 2u) != 0, pz = (r.octinv & 4u) != 0;
    uint32_t mask = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t meta4 = half ? n1.w : n1.z;
        const uint32_t lox = half ? n2.y : n2.x, loy = half ? n2.w : n2.z, loz = half ? n3.y : n3.x;
        const uint32_t hix = half ? n3.w : n3.z, hiy = half ? n4.y : n4.x, hiz = half ? n4.w : n4.z;
        const uint32_t nx4 = px ? lox : hix, fx4 = px ? hix : lox;
        const uint32_t ny4 = py ? loy : hiy, fy4 = py ? hiy : loy;
        const uint32_t nz4 = pz ? loz : hiz, fz4 = pz ? hiz : loz;
        uint32_t bits4, pos4;
        meta_decode4(meta4, r.octinv * 0x01010101u, bits4, pos4);
#if defined(RTB_NODE_CVT_HALF)
        PlanePairs q;
        q.nxl = byte_pair_half2<0>(nx4); q.nxh = byte_pair_half2<1>(nx4); q.nyl = byte_pair_half2<0>(ny4); q.nyh = byte_pair_half2<1>(ny4);
        q.nzl = byte_pair_half2<0>(nz4); q.nzh = byte_pair_half2<1>(nz4); q.fxl = byte_pair_half2<0>(fx4); q.fxh = byte_pair_half2<1>(fx4);
        q.fyl = byte_pair_half2<0>(fy4); q.fyh = byte_pair_half2<1>(fy4); q.fzl = byte_pair_half2<0>(fz4); q.fzh = byte_pair_half2<1>(fz4);
        mask |= child_hit_bits<0>(f, bits4, pos4, q, tmax);
        mask |= child_hit_bits<1>(f, bits4, pos4, q, tmax);
        mask |= child_hit_bits<2>(f, bits4, pos4, q, tmax);
        mask |= child_hit_bits<3>(f, bits4, pos4, q, tmax);
#else
        mask |= child_hit_bits<0>(f, bits4, pos4, nx4, ny4, nz4, fx4, fy4, fz4, tmax);
        mask |= child_hit_bits<1>(f, bits4, pos4, nx4, ny4, nz4, fx4, fy4, fz4, tmax);
        mask |= child_hit_bits<2>(f, bits4, pos4, nx4, ny4, nz4, fx4, fy4, fz4, tmax);
        mask |= child_hit_bits<3>(f, bits4, pos4, nx4, ny4, nz4, fx4, fy4, fz4, tmax);
#endif
    }
    return mask;
}

constexpr int kStackSize = 48;
// The traversal stack: two dynamically indexed arrays in LOCAL memory.  Its top entries live in L1 (write-back), a
// push or pop is one STL / LDL pair per node group that leaves siblings behind; the scalar state of the traversal
// stays in registers because the arrays are outside the Traversal struct.  (rtb_cuda.cu has a variant that keeps the
// first entries in shared memory, HybridStack: measured, not the default — profiles/r1/README.md.)
struct LocalStack {
    uint32_t x[kStackSize], y[kStackSize];
    RTB_HD void put(int i, uint32_t a, uint32_t b) { x[i] = a; y[i] = b; }
    RTB_HD void get(int i, uint32_t &a, uint32_t &b) const { a = x[i]; b = y[i]; }
};

// One ray through the tree, as a resumable state machine so that the
// persistent kernels can swap finished rays for new ones while the rest of
// the warp keeps going.  ANY: stop at the first accepted triangle whose
// leaf-order index differs from `excluded` (the light's own triangle,
// bvh.cuh:239-248).  Otherwise find the closest hit with the reference's
// accept rule 0 < t <= tmax, tmax shrinking (bvh.cuh:222-236).
//
// INST (two-level scenes): the tree at node 0 is the one over the instances.  The leaf lists of that tree name
// instances instead of triangles; entering one takes the ray into the mesh's object space (origin through the
// world->object matrix, direction through its linear part WITHOUT renormalising, so t means the same in both spaces
// and tmax / the accept rule carry over) and continues at the root of the mesh's own tree on the same stack.  The
// instance is left when the stack is back at the height it was entered at; the world ray is then set up again from
// the (o, d) the caller still holds.  What is left of a leaf list waits on the stack as an entry whose top byte is 0
// (node groups always have hit bits there).
template <bool ANY, bool COUNT, bool INST = false>
struct Traversal {
    RaySetup r;
    float tmax;
    uint32_t gx, gy;  // node group: child base | hits<<24 | imask
    int32_t sp, excluded;
    HitRec hit;
    bool found;
    TraceCounters cnt;
    // INST only (dead otherwise)
    int32_t cur;            // instance being traversed, -1 = the tree over the instances
    int32_t sp_enter;       // stack height at which it was entered
    int32_t hit_inst;       // instance of `hit`
    int32_t excluded_inst;  // ANY: instance `excluded` belongs to, -1 = whichever
    // the stack lives OUTSIDE the struct (caller-provided arrays) so that the
    // scalar state above stays in registers instead of following the
    // dynamically indexed arrays into local memory

    RTB_HD void init(V3 o, V3 d, float tmax_, int32_t excluded_) {
        r = ray_setup(o, d);
        tmax = tmax_; excluded = excluded_;
        gx = 0; gy = 0x80000000u; sp = 0;
        hit.t = 0.f; hit.u = 0.f; hit.v = 0.f; hit.tri = -1;
        found = false;
        cnt.nodes = 0; cnt.tris = 0;
        cur = -1; sp_enter = 0; hit_inst = -1; excluded_inst = -1;
    }
    // The three parts of one traversal step.  The persistent kernels call them separately so that
    // the triangle tests of a whole warp can be pooled (rtb_cuda.cu, coop_triangles); step() below
    // is the plain per-ray composition.
    // (1) pop the nearest pending child, fetch its node, slab-test the 8 child boxes;
    //     (tx, ty) = triangle base | hit triangle bits of that node
    template <class ST>
    RTB_HD void node_part(const Bvh8View &B, ST &st, uint32_t &tx, uint32_t &ty) {
        tx = 0; ty = 0;
        if (gy & 0xff000000u) {
            const int bit = bfind(gy);
            gy &= ~(1u << bit);
            const uint32_t base = gx, imask = gy & 0xffu;
            if (gy & 0xff000000u) { st.put(sp, gx, gy); ++sp; }
            const uint32_t slot = ((uint32_t)bit - 24u) ^ r.octinv;
            const uint32_t rel = popc(imask & ~(0xffffffffu << slot));
            const Q4 *np = B.nodes + (size_t)(base + rel) * kNodeWords;
            const Q4 n0 = ldg(np), n1 = ldg(np + 1), n2 = ldg(np + 2), n3 = ldg(np + 3), n4 = ldg(np + 4);
            if (COUNT) cnt.nodes++;
            const uint32_t hm = node_hitmask(n0, n1, n2, n3, n4, r, tmax);
            gx = n1.x; gy = (hm & 0xff000000u) | (n0.w >> 24);
            tx = n1.y; ty = hm & 0x00ffffffu;
        }
    }
    // (2) the accept rule of Triangle::intersect / intersect_leaf (triangle.cuh:49, bvh.cuh:222-248)
    //     for a candidate whose barycentric test passed; true when an any-hit ray is finished.
    //     `t <= tmax` with tmax shrinking: on an exact tie (a ray through the shared edge of two
    //     triangles, bit-identical t) the triangle tested LAST wins, as in the reference.  Which one
    //     that is depends on the tree; with the default collapse it is the reference's choice on every
    //     tie of the fixtures (more than a hundred in tests/golden/s1_hits.npz: rays through the wall
    //     diagonals and corners).  An index-based, tree-independent rule was tried (higher caller index
    //     wins): it disagrees with the reference on 89 of those rays; the SAH-optimal collapse on 48.
    RTB_HD bool accept(const Bvh8View &B, int idx, float t, float u, float v) {
        (void)B;
        if (0.0f < t && t <= tmax) {
            if (ANY) {
                if (idx != excluded || (INST && excluded_inst >= 0 && cur != excluded_inst)) { found = true; return true; }
            } else {
                tmax = t; hit.t = t; hit.u = u; hit.v = v; hit.tri = idx; found = true;
                if (INST) hit_inst = cur;
            }
        }
        return false;
    }
    // (3) next pending node group; false when the ray is finished
    template <class ST>
    RTB_HD bool advance(const ST &st) {
        if ((gy & 0xff000000u) == 0) {
            if (sp == 0) return false;
            --sp; st.get(sp, gx, gy);
        }
        return true;
    }
    // ---- INST ----
    // (tx, ty) is a leaf list of the top tree: enter the instance of its highest entry.  The node group still pending
    // and the rest of the list go on the stack first, so they are resumed when the instance has been left.
    template <class ST>
    RTB_HD void enter_instance(const Bvh8View &B, ST &st, uint32_t tx, uint32_t &ty, V3 o, V3 d) {
        const int bit = bfind(ty);
        ty &= ~(1u << bit);
        if (gy & 0xff000000u) { st.put(sp, gx, gy); ++sp; }
        if (ty) { st.put(sp, tx, ty); ++sp; ty = 0u; }
        cur = B.top_inst[tx + (uint32_t)bit];
        sp_enter = sp;
        const F4 *m = B.inst + (size_t)cur * kInstWords;
        const F4 r0 = ldg(m), r1 = ldg(m + 1), r2 = ldg(m + 2);
        const int root = f2i(ldg(m + 6).x);
        r = ray_setup(xform_point(r0, r1, r2, o), xform_vector(r0, r1, r2, d));
        gx = (uint32_t)root; gy = 0x80000000u;
    }
    // next pending node group or leaf list; (o, d) = the world ray, set up again when an instance is left
    template <class ST>
    RTB_HD bool advance_inst(const ST &st, uint32_t &tx, uint32_t &ty, V3 o, V3 d) {
        if ((gy & 0xff000000u) == 0) {
            if (cur >= 0 && sp == sp_enter) { cur = -1; r = ray_setup(o, d); }
            if (sp == 0) return false;
            --sp;
            uint32_t x, y;
            st.get(sp, x, y);
            if (y & 0xff000000u) { gx = x; gy = y; }
            else { tx = x; ty = y; gx = 0u; gy = 0u; }  // the rest of a leaf list of the top tree
        }
        return true;
    }
    template <class ST>
    RTB_HD bool step_inst(const Bvh8View &B, ST &st, uint32_t &tx, uint32_t &ty, V3 o, V3 d) {
        if (ty == 0u) node_part(B, st, tx, ty);
        if (cur < 0) {
            if (ty) enter_instance(B, st, tx, ty, o, d);
        } else {
            while (ty) {
                const int bit = bfind(ty);
                ty &= ~(1u << bit);
                const int idx = (int)(tx + (uint32_t)bit);
                const Tri48 tr = load_tri(B.tris, idx);
                if (COUNT) cnt.tris++;
                float u, v;
                const float t = tri_candidate(tr, r.o, r.d, u, v);
                if (accept(B, idx, t, u, v)) return false;
            }
        }
        return advance_inst(st, tx, ty, o, d);
    }
    // one node (its 8 child boxes) plus the triangles it exposes; false when the ray is finished
    template <class ST>
    RTB_HD bool step(const Bvh8View &B, ST &st) {
        uint32_t tx, ty;  // triangle group: tri base | hit bits
        node_part(B, st, tx, ty);
        while (ty) {
            const int bit = bfind(ty);
            ty &= ~(1u << bit);
            const int idx = (int)(tx + (uint32_t)bit);
            const Tri48 tr = load_tri(B.tris, idx);
            if (COUNT) cnt.tris++;
            float u, v;
            const float t = tri_candidate(tr, r.o, r.d, u, v);
            if (accept(B, idx, t, u, v)) return false;
        }
        return advance(st);
    }
};

template <bool ANY, bool COUNT>
RTB_HD bool bvh8_trace(const Bvh8View &B, V3 o, V3 d, float tmax, int32_t excluded, HitRec &hit,
                       TraceCounters *cnt) {
    Traversal<ANY, COUNT> T;
    LocalStack st;
    T.init(o, d, tmax, excluded);
    while (T.step(B, st)) {}
    hit = T.hit;
    if (COUNT) { cnt->nodes = T.cnt.nodes; cnt->tris = T.cnt.tris; }
    return T.found;
}
// the same through a two-level scene; hit_inst = instance of the hit (-1: miss), excluded_inst = instance of
// `excluded` or -1 (whichever instance)
template <bool ANY, bool COUNT>
RTB_HD bool bvh8_trace_inst(const Bvh8View &B, V3 o, V3 d, float tmax, int32_t excluded, int32_t excluded_inst, HitRec &hit,
                            int32_t &hit_inst, TraceCounters *cnt) {
    Traversal<ANY, COUNT, true> T;
    LocalStack st;
    T.init(o, d, tmax, excluded);
    T.excluded_inst = excluded_inst;
    uint32_t tx = 0u, ty = 0u;
    while (T.step_inst(B, st, tx, ty, o, d)) {}
    hit = T.hit; hit_inst = T.hit_inst;
    if (COUNT) { cnt->nodes = T.cnt.nodes; cnt->tris = T.cnt.tris; }
    return T.found;
}
// whichever the scene is
template <bool ANY, bool COUNT>
RTB_HD bool scene_trace(const Bvh8View &B, V3 o, V3 d, float tmax, int32_t excluded, int32_t excluded_inst, HitRec &hit,
                        int32_t &hit_inst, TraceCounters *cnt) {
    if (B.inst) return bvh8_trace_inst<ANY, COUNT>(B, o, d, tmax, excluded, excluded_inst, hit, hit_inst, cnt);
    hit_inst = -1;
    return bvh8_trace<ANY, COUNT>(B, o, d, tmax, excluded, hit, cnt);
}

}  // namespace rtb
