// oracle.cpp — TEST INFRASTRUCTURE: scalar host-C++ restatement of the render
// path of lashhw/rtcuda.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this; the product
// (rtcuda_b200/librtb.so) never does and has no CPU fallback.
//
// What is restated, and from where (paths relative to the reference repo):
//   Triangle ctor / bounding_box / center / intersect   triangle.cuh:4-76
//   full-sweep SAH binary BVH build                      bvh.cuh:30-219
//   ordered closest-hit / any-hit traversal + leaf loops bvh.cuh:222-357
//   slab test without [0,tmax] clamp                     aabb_intersector.cuh:14-36
//   camera                                               camera.cuh:15-34
//   BSDFs                                                material.cuh:47-109
//   lights                                               light.cuh:29-64
//   ray offset, sphere sampling, power heuristic         utility.cuh:31-77
//   per-slot state machine and shading, per path         render.cuh:84-338
// Device-side arithmetic that decides WHICH triangle is hit (triangle test,
// slab test, camera ray) uses the FMA contraction pattern nvcc emits for the
// reference on sm_100a, written out with fmaf(); host-side
// arithmetic (Triangle ctor, BVH build, Camera ctor) is plain float as g++
// compiles it.  Build with -ffp-contract=off.
//
// Parity status: the reference ships no golden vectors.  This
// oracle is pinned against outputs of the reference's own CUDA code run on a
// B200 through oracle/ref_harness.cu; the fixtures and the script that made
// them are in tests/golden/.  The RNG is a seam: the reference's per-slot
// cuRAND XORWOW streams are not reproducible outside it, so the oracle and the
// product share a counter-based generator (restated below) and image parity
// with the reference itself is statistical (tests/golden/ref_*).
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <float.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <memory>
#include <numeric>
#include <stack>
#include <thread>
#include <vector>

#include "rtb.h"  // POD interface types only (rtb_scene_desc, rtb_ray, rtb_hit, rtb_camera, ...)

namespace {

struct Vec3 {
    float x, y, z;
};
inline Vec3 V(float x, float y, float z) { return Vec3{x, y, z}; }
inline Vec3 operator+(Vec3 a, Vec3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator*(Vec3 a, Vec3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
inline Vec3 operator*(Vec3 a, float t) { return V(a.x * t, a.y * t, a.z * t); }
inline Vec3 operator*(float t, Vec3 a) { return V(a.x * t, a.y * t, a.z * t); }
inline Vec3 operator-(Vec3 a) { return V(-a.x, -a.y, -a.z); }
inline Vec3 operator/(Vec3 a, float t) { float i = 1.f / t; return V(a.x * i, a.y * i, a.z * i); }  // vec3.cuh:56-59
// host flavour (no contraction): vec3.cuh:61-69 as g++ compiles it
inline float dot_host(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross_host(Vec3 a, Vec3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// device flavour: the same expressions as nvcc contracts them (SASS of `ch`)
inline float dot_dev(Vec3 a, Vec3 b) { return fmaf(a.z, b.z, fmaf(a.x, b.x, a.y * b.y)); }
inline Vec3 cross_dev(Vec3 a, Vec3 b) {
    return V(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
inline float length_dev(Vec3 a) { return sqrtf(dot_dev(a, a)); }
inline Vec3 unit_dev(Vec3 a) { float i = 1.f / length_dev(a); return V(a.x * i, a.y * i, a.z * i); }
inline float max3(Vec3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }

struct Ray {
    Vec3 origin, unit_d;
    float tmax;
};
struct Isect {
    float t, u, v;
};

struct Triangle {  // triangle.cuh:4-21
    Vec3 p0, e1, e2, n;
    Triangle() {}
    Triangle(Vec3 a, Vec3 b, Vec3 c) : p0(a), e1(a - b), e2(c - a), n(cross_host(e1, e2)) {}
    Vec3 p1() const { return p0 - e1; }
    Vec3 p2() const { return p0 + e2; }
    Vec3 center() const { return (p0 + p1() + p2()) * (1.f / 3.f); }
    Vec3 p(float u, float v) const { return V(fmaf(v, e2.x, fmaf(-u, e1.x, p0.x)), fmaf(v, e2.y, fmaf(-u, e1.y, p0.y)), fmaf(v, e2.z, fmaf(-u, e1.z, p0.z))); }
    // triangle.cuh:39-58
    bool intersect(const Ray &ray, Isect &is) const {
        Vec3 c = p0 - ray.origin;
        Vec3 r = cross_dev(ray.unit_d, c);
        float inv_det = 1.f / dot_dev(ray.unit_d, n);
        float u = inv_det * dot_dev(e2, r);
        float v = inv_det * dot_dev(e1, r);
        if (u >= 0.0f && v >= 0.0f && (u + v) <= 1.0f) {
            float t = inv_det * dot_dev(c, n);
            if (0 < t && t <= ray.tmax) { is.t = t; is.u = u; is.v = v; return true; }
        }
        return false;
    }
};

struct Box {  // bounding_box.cuh:4-37, bounds = xmin xmax ymin ymax zmin zmax
    float b[6];
    void reset() { b[0] = b[2] = b[4] = FLT_MAX; b[1] = b[3] = b[5] = -FLT_MAX; }
    void extend(const Box &o) {
        b[0] = fminf(b[0], o.b[0]); b[1] = fmaxf(b[1], o.b[1]); b[2] = fminf(b[2], o.b[2]);
        b[3] = fmaxf(b[3], o.b[3]); b[4] = fminf(b[4], o.b[4]); b[5] = fmaxf(b[5], o.b[5]);
    }
    float half_area() const { float e1 = b[1] - b[0], e2 = b[3] - b[2], e3 = b[5] - b[4]; return (e1 + e2) * e3 + e1 * e2; }
};
Box tri_box(const Triangle &t) {  // triangle.cuh:23-37
    Vec3 a = t.p0, b = t.p1(), c = t.p2();
    Box r;
    r.b[0] = fminf(a.x, fminf(b.x, c.x)); r.b[2] = fminf(a.y, fminf(b.y, c.y)); r.b[4] = fminf(a.z, fminf(b.z, c.z));
    r.b[1] = fmaxf(a.x, fmaxf(b.x, c.x)); r.b[3] = fmaxf(a.y, fmaxf(b.y, c.y)); r.b[5] = fmaxf(a.z, fmaxf(b.z, c.z));
    return r;
}

struct Node {  // bvh.cuh:5-14
    Box bbox;
    int num_primitives;  // 0 => inner
    int index;           // left child (right = left+1) or first primitive
};

constexpr int kMaxDepth = 30;  // BVH_MAX_DEPTH, constant.hpp:7

struct Scene {
    std::vector<Triangle> tris;   // caller order
    std::vector<int> mat_id, light_id;
    std::vector<rtb_material> materials;
    std::vector<rtb_light> lights;
    std::vector<int> order;       // BVH primitive order -> caller index (bvh.cuh:208)
    std::vector<Node> nodes;
    int max_depth = 0;
};

// bvh.cuh:30-219
void build_bvh(Scene &s) {
    const int n = (int)s.tris.size();
    s.nodes.assign(std::max(2 * n, 1), Node());
    s.order.resize(n);
    if (n == 0) { s.nodes[0].bbox.reset(); s.nodes[0].num_primitives = 0; s.nodes[0].index = 0; s.nodes.resize(1); return; }
    std::vector<Box> boxes(n);
    std::vector<Vec3> centers(n);
    std::vector<float> costs(n);
    std::vector<char> marks(n);
    std::vector<int> refs[3];
    int num_nodes = 1;
    s.nodes[0].bbox.reset();
    for (int i = 0; i < n; i++) {
        boxes[i] = tri_box(s.tris[i]);
        s.nodes[0].bbox.extend(boxes[i]);
        centers[i] = s.tris[i].center();
    }
    for (int a = 0; a < 3; ++a) {
        refs[a].resize(n);
        std::iota(refs[a].begin(), refs[a].end(), 0);
    }
    std::sort(refs[0].begin(), refs[0].end(), [&](int i, int j) { return centers[i].x < centers[j].x; });
    std::sort(refs[1].begin(), refs[1].end(), [&](int i, int j) { return centers[i].y < centers[j].y; });
    std::sort(refs[2].begin(), refs[2].end(), [&](int i, int j) { return centers[i].z < centers[j].z; });
    struct Job { int node, begin, end, depth; };
    std::stack<Job> todo;
    Job cur{0, 0, n, 0};
    auto next = [&]() { if (todo.empty()) return false; cur = todo.top(); todo.pop(); return true; };
    while (true) {
        Node &node = s.nodes[cur.node];
        const int count = cur.end - cur.begin;
        bool leaf = count <= 1 || cur.depth >= kMaxDepth;
        float best_cost = FLT_MAX; int best_axis = -1, best_split = -1;
        if (!leaf) {
            for (int axis = 0; axis < 3; axis++) {
                Box acc; acc.reset();
                for (int i = cur.end - 1; i > cur.begin; i--) {
                    acc.extend(boxes[refs[axis][i]]);
                    costs[i] = acc.half_area() * (cur.end - i);
                }
                acc.reset();
                for (int i = cur.begin; i < cur.end - 1; i++) {
                    acc.extend(boxes[refs[axis][i]]);
                    float cost = acc.half_area() * (i + 1 - cur.begin) + costs[i + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = i + 1; }
                }
            }
            if (best_cost >= node.bbox.half_area() * (count - 1)) leaf = true;
        }
        if (leaf) {
            node.num_primitives = count;
            node.index = cur.begin;
            if (next()) continue;
            break;
        }
        const int left = num_nodes, right = num_nodes + 1;
        s.nodes[left].bbox.reset();
        s.nodes[right].bbox.reset();
        for (int i = cur.begin; i < best_split; i++) { s.nodes[left].bbox.extend(boxes[refs[best_axis][i]]); marks[refs[best_axis][i]] = 1; }
        for (int i = best_split; i < cur.end; i++) { s.nodes[right].bbox.extend(boxes[refs[best_axis][i]]); marks[refs[best_axis][i]] = 0; }
        for (int k = 1; k <= 2; ++k) {
            std::vector<int> &r = refs[(best_axis + k) % 3];
            std::stable_partition(r.begin() + cur.begin, r.begin() + cur.end, [&](int i) { return marks[i] != 0; });
        }
        num_nodes += 2;
        node.num_primitives = 0;
        node.index = left;
        s.max_depth = std::max(s.max_depth, cur.depth + 1);
        const int ls = best_split - cur.begin, rs = cur.end - best_split;
        if (ls < rs) {  // smaller subtree first
            todo.push(Job{right, best_split, cur.end, cur.depth + 1});
            cur = Job{left, cur.begin, best_split, cur.depth + 1};
        } else {
            todo.push(Job{left, cur.begin, best_split, cur.depth + 1});
            cur = Job{right, best_split, cur.end, cur.depth + 1};
        }
    }
    s.nodes.resize(num_nodes);
    for (int i = 0; i < n; ++i) s.order[i] = refs[0][i];  // bvh.cuh:208
}

// aabb_intersector.cuh:4-36
struct Slab {
    int ox, oy, oz;
    Vec3 inv, so;
    explicit Slab(const Ray &r) {
        ox = r.unit_d.x < 0 ? 1 : 0; oy = r.unit_d.y < 0 ? 1 : 0; oz = r.unit_d.z < 0 ? 1 : 0;
        auto safe = [](float d) { return 1.f / ((fabsf(d) < FLT_EPSILON) ? copysignf(FLT_EPSILON, d) : d); };
        inv = V(safe(r.unit_d.x), safe(r.unit_d.y), safe(r.unit_d.z));
        so = V(-r.origin.x * inv.x, -r.origin.y * inv.y, -r.origin.z * inv.z);
    }
    bool hit(const Box &b, float &entry) const {
        float ex = fmaf(inv.x, b.b[0 + ox], so.x), ey = fmaf(inv.y, b.b[2 + oy], so.y), ez = fmaf(inv.z, b.b[4 + oz], so.z);
        entry = fmaxf(ex, fmaxf(ey, ez));
        float xx = fmaf(inv.x, b.b[1 - ox], so.x), xy = fmaf(inv.y, b.b[3 - oy], so.y), xz = fmaf(inv.z, b.b[5 - oz], so.z);
        return entry <= fminf(xx, fminf(xy, xz));
    }
};

struct Counters {
    uint64_t node_pairs = 0, tri_tests = 0;
};

// bvh.cuh:222-236 / 239-248
inline bool leaf_closest(const Scene &s, const Node &nd, Ray &ray, Isect &is, int &prim, Counters *c) {
    bool any = false;
    for (int i = nd.index; i < nd.index + nd.num_primitives; i++) {
        if (c) c->tri_tests++;
        if (s.tris[s.order[i]].intersect(ray, is)) { any = true; prim = s.order[i]; ray.tmax = is.t; }
    }
    return any;
}
inline bool leaf_any(const Scene &s, const Node &nd, const Ray &ray, int excluded) {
    for (int i = nd.index; i < nd.index + nd.num_primitives; i++) {
        Isect is;
        if (s.tris[s.order[i]].intersect(ray, is) && s.order[i] != excluded) return true;
    }
    return false;
}

// bvh.cuh:251-303 (ANY=false) and bvh.cuh:306-357 (ANY=true)
template <bool ANY>
bool traverse(const Scene &s, Ray &ray, Isect &is, int &prim, int excluded, Counters *c) {
    const Node *nodes = s.nodes.data();
    if (s.tris.empty()) return false;
    if (nodes[0].num_primitives > 0) return ANY ? leaf_any(s, nodes[0], ray, excluded) : leaf_closest(s, nodes[0], ray, is, prim, c);
    bool found = false;
    Slab slab(ray);
    int stack[kMaxDepth - 1 + 4], sp = 0;
    int left = nodes[0].index;
    while (true) {
        if (c) c->node_pairs++;
        const Node &L = nodes[left], &R = nodes[left + 1];
        float el, er;
        bool go_l = false, go_r = false;
        if (slab.hit(L.bbox, el)) {
            if (L.num_primitives > 0) { if (ANY) { if (leaf_any(s, L, ray, excluded)) return true; } else found |= leaf_closest(s, L, ray, is, prim, c); }
            else go_l = true;
        }
        if (slab.hit(R.bbox, er)) {
            if (R.num_primitives > 0) { if (ANY) { if (leaf_any(s, R, ray, excluded)) return true; } else found |= leaf_closest(s, R, ray, is, prim, c); }
            else go_r = true;
        }
        if (go_l && go_r) {
            if (el > er) { stack[sp++] = L.index; left = R.index; }
            else { stack[sp++] = R.index; left = L.index; }
        } else if (go_l) left = L.index;
        else if (go_r) left = R.index;
        else { if (sp == 0) break; left = stack[--sp]; }
    }
    return found;
}

// ---- counter-based RNG shared with the product (rtcuda_b200/csrc/rtb_core.h
// states the same function; it replaces curand XORWOW, render.cuh:68-73) ----
struct U4 { uint32_t x, y, z, w; };
inline U4 pcg4d(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    U4 v{a * 1664525u + 1013904223u, b * 1664525u + 1013904223u, c * 1664525u + 1013904223u, d * 1664525u + 1013904223u};
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    v.x ^= v.x >> 16; v.y ^= v.y >> 16; v.z ^= v.z >> 16; v.w ^= v.w >> 16;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    return v;
}
inline float u01(uint32_t b) { return (float)((b >> 8) + 1u) * 5.9604644775390625e-08f; }  // (0,1] like curand_uniform
struct R4 { float a, b, c, d; };
inline R4 rand4(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t block) {
    U4 r = pcg4d(pixel, sample, block, seed);
    return R4{u01(r.x), u01(r.y), u01(r.z), u01(r.w)};
}

// utility.cuh:31-47
inline float as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline int as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
Vec3 offset_ray_origin(Vec3 p, Vec3 n) {
    const float int_scale = 256.f, float_scale = 1.f / 65536.f, origin = 1.f / 32.f;
    int ix = (int)(int_scale * n.x), iy = (int)(int_scale * n.y), iz = (int)(int_scale * n.z);
    float px = as_float(as_int(p.x) + (p.x < 0 ? -ix : ix));
    float py = as_float(as_int(p.y) + (p.y < 0 ? -iy : iy));
    float pz = as_float(as_int(p.z) + (p.z < 0 ? -iz : iz));
    return V(fabsf(p.x) < origin ? fmaf(float_scale, n.x, p.x) : px, fabsf(p.y) < origin ? fmaf(float_scale, n.y, p.y) : py,
             fabsf(p.z) < origin ? fmaf(float_scale, n.z, p.z) : pz);
}

const float PI_F = 3.14159265358979323846f, TWO_PI_F = 6.28318530717958647692f, INV_PI_F = 0.31830988618379067153f;

Vec3 reflect(Vec3 v, Vec3 n) { return v - n * (2.f * dot_dev(v, n)); }  // vec3.cuh:71-73
Vec3 sample_sphere(float u1, float u2) {  // utility.cuh:70-77
    float z = 1.f - 2.f * u1;
    float r = sqrtf(fmaxf(0.f, 1.f - z * z));
    float phi = TWO_PI_F * u2, sn, cs;
    sincosf(phi, &sn, &cs);
    return V(r * cs, r * sn, z);
}

// ---- RTB_GLOSSY: not in the reference (rtb.h); an energy-normalised Phong lobe around the mirror direction ----
void onb(Vec3 n, Vec3 &t, Vec3 &b) {  // Duff et al. 2017
    float sg = copysignf(1.f, n.z);
    float a = -1.f / (sg + n.z);
    float c = n.x * n.y * a;
    t = V(fmaf(sg * (n.x * n.x), a, 1.f), sg * c, -sg * n.x);
    b = V(c, fmaf(n.y * n.y, a, sg), -n.y);
}
void glossy_eval(const rtb_material &m, Vec3 wo, Vec3 n, Vec3 wi, Vec3 &f, float &pdf) {
    float ca = dot_dev(reflect(wo, n), wi);
    if (ca <= 0.f) { f = V(0, 0, 0); pdf = 0.f; return; }
    float lobe = powf(ca, m.ior) * 0.15915494309189535f;
    pdf = (m.ior + 1.f) * lobe;
    f = V(m.albedo[0], m.albedo[1], m.albedo[2]) * ((m.ior + 2.f) * lobe);
}

// material.cuh:60-109
Vec3 sample_f(const rtb_material &m, Vec3 wo, float u1, float u2, Vec3 &n, Vec3 &wi, float &pdf) {
    Vec3 albedo = V(m.albedo[0], m.albedo[1], m.albedo[2]);
    if (m.type == RTB_GLOSSY) {
        if (dot_dev(wo, n) > 0.f) n = -n;
        Vec3 r = reflect(wo, n);
        float e = m.ior;
        float ca = powf(u1, 1.f / (e + 1.f));
        float sa = sqrtf(fmaxf(0.f, 1.f - ca * ca));
        float sp, cp;
        sincosf(TWO_PI_F * u2, &sp, &cp);
        Vec3 t, b;
        onb(r, t, b);
        wi = (t * (sa * cp) + b * (sa * sp)) + r * ca;
        float lobe = powf(ca, e) * 0.15915494309189535f;
        pdf = (e + 1.f) * lobe;
        return albedo * ((e + 2.f) * lobe);
    }
    if (m.type == RTB_MATTE || m.type == RTB_MIRROR) {
        if (dot_dev(wo, n) > 0.f) n = -n;
        if (m.type == RTB_MATTE) {
            wi = unit_dev(n + sample_sphere(u1, u2));
            pdf = dot_dev(wi, n) * INV_PI_F;
            return albedo * INV_PI_F;
        }
        wi = reflect(wo, n);
        pdf = 1.f;
        return albedo / dot_dev(wi, n);
    }
    float cos_theta = dot_dev(wo, n);
    bool front = cos_theta < 0.f;
    if (front) cos_theta = -cos_theta;
    float inv_cos = 1.f / cos_theta;
    float eta = front ? 1.f / m.ior : m.ior;
    float sin_theta = sqrtf(1.f - cos_theta * cos_theta);
    if (!front) n = -n;
    if (eta * sin_theta > 1.f) { wi = reflect(wo, n); pdf = 1.f; return V(inv_cos, inv_cos, inv_cos); }
    float r0 = (1.f - m.ior) / (1.f + m.ior);
    r0 = r0 * r0;
    float refl = fmaf(1.f - r0, powf(1.f - cos_theta, 5.f), r0);
    if (u1 < refl) { wi = reflect(wo, n); pdf = refl; float f = refl * inv_cos; return V(f, f, f); }
    Vec3 par = V(fmaf(cos_theta, n.x, wo.x), fmaf(cos_theta, n.y, wo.y), fmaf(cos_theta, n.z, wo.z)) * eta;  // vec3.cuh:82-86
    Vec3 perp = n * -sqrtf(1.f - dot_dev(par, par));
    wi = par + perp;
    n = -n;
    pdf = 1.f - refl;
    float f = pdf * eta * eta / dot_dev(wi, n);
    return V(f, f, f);
}

struct RenderStats { uint64_t paths = 0, extend = 0, shadow = 0; };

// render.cuh:84-328 for one camera path
void trace_path(const Scene &s, const rtb_camera &cam, const rtb_render_params &p, uint32_t pixel, uint32_t sample,
                float *fb, RenderStats &st) {
    const int i = (int)(pixel % (uint32_t)p.width), j = (int)(pixel / (uint32_t)p.width);
    float jx = 0.5f, jy = 0.5f;
    if (!(p.flags & RTB_RENDER_PIXEL_CENTRE)) { R4 r = rand4(p.seed, pixel, sample, 0); jx = r.a; jy = r.b; }
    const float x = ((float)i + jx) / (float)p.width, y = ((float)j + jy) / (float)p.height;
    const Vec3 ul = V(cam.upper_left[0], cam.upper_left[1], cam.upper_left[2]), hz = V(cam.horizontal[0], cam.horizontal[1], cam.horizontal[2]),
               vt = V(cam.vertical[0], cam.vertical[1], cam.vertical[2]), from = V(cam.lookfrom[0], cam.lookfrom[1], cam.lookfrom[2]);
    // camera.cuh:31-34
    Vec3 dir = V(fmaf(y, vt.x, fmaf(x, hz.x, ul.x)), fmaf(y, vt.y, fmaf(x, hz.y, ul.y)), fmaf(y, vt.z, fmaf(x, hz.z, ul.z))) - from;
    Ray ray{from, unit_dev(dir), FLT_MAX};
    Vec3 beta = V(1, 1, 1);
    int b = 0;
    st.paths++;
    Isect is; int prim = -1;
    st.extend++;
    bool hit = traverse<false>(s, ray, is, prim, -1, nullptr);
    const int nl = (int)s.lights.size();
    // beyond the reference (off by default): rtb.h RTB_RENDER_TRUE_MIS / RTB_RENDER_RR_TERMINATE / env_L
    const bool true_mis = (p.flags & RTB_RENDER_TRUE_MIS) != 0;
    const bool has_env = p.env_L[0] != 0.f || p.env_L[1] != 0.f || p.env_L[2] != 0.f;
    auto add_env = [&](const Vec3 &bt) {
        Vec3 L = bt * V(p.env_L[0], p.env_L[1], p.env_L[2]);
        if (fabsf(L.x) <= FLT_MAX && fabsf(L.y) <= FLT_MAX && fabsf(L.z) <= FLT_MAX) { fb[0] += L.x; fb[1] += L.y; fb[2] += L.z; }
    };
    if (!hit && has_env) add_env(beta);
    float prev_pdf = 0.f;  // solid-angle pdf of the BSDF sample that produced `ray`, 0 = delta / camera
    while (true) {
        if (hit && s.light_id[prim] >= 0 && (b == 0 || true_mis)) {  // render.cuh:98-107
            const rtb_light &l = s.lights[s.light_id[prim]];
            if (b == 0) {
                fb[0] += l.L[0]; fb[1] += l.L[1]; fb[2] += l.L[2];
            } else {  // the path ray is the BSDF sample of the MIS pair
                float w = 1.f;
                const Triangle &lt = s.tris[prim];
                if (prev_pdf > 0.f) {
                    float area = 0.5f * length_dev(lt.n);
                    float cosl = fabsf(dot_dev(unit_dev(lt.n), ray.unit_d));
                    float pdf_l = ((is.t * is.t) / (area * cosl)) / (float)nl;
                    float a2 = prev_pdf * prev_pdf;
                    w = a2 / (a2 + pdf_l * pdf_l);
                }
                Vec3 L = (V(l.L[0], l.L[1], l.L[2]) * w) * beta;
                if (fabsf(L.x) <= FLT_MAX && fabsf(L.y) <= FLT_MAX && fabsf(L.z) <= FLT_MAX) { fb[0] += L.x; fb[1] += L.y; fb[2] += L.z; }
            }
        }
        if (b >= p.max_bounces) break;  // render.cuh:109
        if (!hit) break;                // the reference idles the slot instead (Quirk B)
        if (b > p.rr_start) {           // render.cuh:112-124
            float bm = max3(beta);
            if (bm < p.rr_threshold) {
                float pt = fmaxf(0.05f, 1.f - bm);
                float u = rand4(p.seed, pixel, sample, 2u * (uint32_t)b + 1u).a;
                if (u < pt) {
                    if (p.flags & RTB_RENDER_RR_TERMINATE) break;
                    b++; continue;  // Quirk A: pause, roll again on the same hit
                }
                beta = beta / (1.f - pt);
            }
        }
        const R4 xi = rand4(p.seed, pixel, sample, 2u * (uint32_t)b + 2u);
        b++;
        // render.cuh:139-168
        const Triangle &tr = s.tris[prim];
        const rtb_material &m = s.materials[s.mat_id[prim]];
        const Vec3 wo = ray.unit_d;
        const Vec3 P = tr.p(is.u, is.v);
        const Vec3 ng = -unit_dev(tr.n);
        const Vec3 beta_old = beta;
        Vec3 n1 = ng, wi1; float pdf1;
        Vec3 f1 = sample_f(m, wo, xi.a, xi.b, n1, wi1, pdf1);
        beta = beta * ((f1 * dot_dev(wi1, n1)) * (1.f / pdf1));
        Ray next{offset_ray_origin(P, n1), wi1, FLT_MAX};
        const bool next_ok = !(m.type == RTB_GLOSSY && !(dot_dev(wi1, n1) > 0.f && pdf1 > 0.f));  // lobe sample below the surface
        // render.cuh:170-211
        if (nl > 0 && !(p.flags & RTB_RENDER_NO_SHADOW)) {
            int li = (int)(xi.c * (float)nl);
            if (li > nl - 1) li = nl - 1;
            const rtb_light &l = s.lights[li];
            const float u2 = rand4(p.seed, pixel, sample, 0x80000000u + (uint32_t)b).a;
            Vec3 wiL, Li; float tL, pdfL; int excl = -1;
            if (l.type == RTB_POINT_LIGHT) {  // light.cuh:31-37
                Vec3 w = V(l.pos[0], l.pos[1], l.pos[2]) - P;
                tL = length_dev(w);
                Li = V(l.L[0], l.L[1], l.L[2]) * (1.f / (tL * tL));
                wiL = w * (1.f / tL);
                pdfL = 1.f;
            } else {  // light.cuh:38-46, triangle.cuh:78-86
                const Triangle &lt = s.tris[(size_t)l.triangle];
                float area = 0.5f * length_dev(lt.n);
                float a = sqrtf(xi.d);
                Vec3 q = lt.p(1.f - a, u2 * a);
                Vec3 w = q - P;
                tL = length_dev(w);
                wiL = w * (1.f / tL);
                Li = V(l.L[0], l.L[1], l.L[2]);
                pdfL = (1.f / area) * (dot_dev(w, w) / fabsf(dot_dev(unit_dev(lt.n), wiL)));
                excl = (int)l.triangle;
            }
            Vec3 nL = dot_dev(ng, wiL) > 0.f ? ng : -ng;
            if ((m.type == RTB_MATTE || m.type == RTB_GLOSSY) && dot_dev(wo, nL) * dot_dev(wiL, nL) < 0.f) {  // get_f + same_hemisphere
                float cosl = dot_dev(wiL, nL);
                Vec3 f = (V(m.albedo[0], m.albedo[1], m.albedo[2]) * INV_PI_F) * cosl;
                float spdf = cosl * INV_PI_F;
                if (m.type == RTB_GLOSSY) { glossy_eval(m, wo, nL, wiL, f, spdf); f = f * cosl; }
                Vec3 L = ((beta_old * (float)nl) * f) * Li;
                if (l.type != RTB_POINT_LIGHT) {
                    if (true_mis) {
                        float pl = pdfL / (float)nl;
                        float a2 = pl * pl;
                        L = L * (a2 / (a2 + spdf * spdf));
                    } else if (m.type == RTB_MATTE) {  // power_heuristic(float, int): Quirk C (RTB_GLOSSY: weight 1)
                        int g = (int)spdf;
                        float f2 = pdfL * pdfL;
                        L = L * (f2 / (f2 + (float)(g * g)));
                    }
                }
                L = L * (1.f / pdfL);
                Ray sh{offset_ray_origin(P, nL), wiL, tL};
                Isect dummy; int dp;
                st.shadow++;
                if (!traverse<true>(s, sh, dummy, dp, excl, nullptr)) {
                    if (fabsf(L.x) <= FLT_MAX && fabsf(L.y) <= FLT_MAX && fabsf(L.z) <= FLT_MAX) { fb[0] += L.x; fb[1] += L.y; fb[2] += L.z; }
                }
            }
            // render.cuh:213-245: the BSDF-sampled MIS ray targets the shading
            // triangle itself (Quirk D) and never contributes: not traced.
        }
        if (b >= p.max_bounces) break;  // the reference traces `next` and discards the result
        if (!next_ok) break;
        ray = next;
        prev_pdf = (m.type == RTB_MATTE || m.type == RTB_GLOSSY) ? pdf1 : 0.f;
        st.extend++;
        hit = traverse<false>(s, ray, is, prim, -1, nullptr);
        if (!hit && has_env) add_env(beta);
    }
}

template <class F>
void parallel_for(int64_t n, int nthreads, F f) {
    if (nthreads <= 1) { for (int64_t i = 0; i < n; ++i) f(i, 0); return; }
    std::atomic<int64_t> next(0);
    std::vector<std::thread> th;
    const int64_t chunk = std::max<int64_t>(1, n / (nthreads * 16));
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t] {
            while (true) {
                int64_t b = next.fetch_add(chunk);
                if (b >= n) break;
                for (int64_t i = b; i < std::min(n, b + chunk); ++i) f(i, t);
            }
        });
    for (auto &t : th) t.join();
}

}  // namespace

extern "C" {
#define ORC_API __attribute__((visibility("default")))

ORC_API void *orc_scene_create(const rtb_scene_desc *d) {
    Scene *s = new Scene();
    const int64_t n = d->num_triangles;
    s->tris.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const float *v = d->vertices + 9 * i;
        s->tris.emplace_back(V(v[0], v[1], v[2]), V(v[3], v[4], v[5]), V(v[6], v[7], v[8]));
    }
    s->mat_id.assign(d->material_ids, d->material_ids + n);
    if (d->light_ids) s->light_id.assign(d->light_ids, d->light_ids + n); else s->light_id.assign((size_t)n, -1);
    s->materials.assign(d->materials, d->materials + d->num_materials);
    if (d->num_lights) s->lights.assign(d->lights, d->lights + d->num_lights);
    build_bvh(*s);
    return s;
}
ORC_API void orc_scene_destroy(void *h) { delete (Scene *)h; }
ORC_API void orc_bvh_stats(void *h, int *num_nodes, int *max_depth) {
    Scene *s = (Scene *)h;
    *num_nodes = (int)s->nodes.size();
    *max_depth = s->max_depth;
}
ORC_API void orc_trace_closest(void *h, const rtb_ray *rays, int64_t n, rtb_hit *hits, int nthreads) {
    const Scene &s = *(Scene *)h;
    parallel_for(n, nthreads, [&](int64_t i, int) {
        Ray r{V(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), V(rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]), rays[i].tmax};
        Isect is{0, 0, 0}; int prim = -1;
        bool hit = traverse<false>(s, r, is, prim, -1, nullptr);
        hits[i].t = hit ? is.t : 0.f; hits[i].u = hit ? is.u : 0.f; hits[i].v = hit ? is.v : 0.f; hits[i].prim = hit ? prim : -1;
    });
}
ORC_API void orc_trace_any(void *h, const rtb_ray *rays, const int32_t *excluded, int64_t n, uint8_t *occ, int nthreads) {
    const Scene &s = *(Scene *)h;
    parallel_for(n, nthreads, [&](int64_t i, int) {
        Ray r{V(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), V(rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]), rays[i].tmax};
        Isect is; int prim;
        occ[i] = traverse<true>(s, r, is, prim, excluded ? excluded[i] : -1, nullptr) ? 1 : 0;
    });
}
// mean sibling pairs visited / triangles tested per closest-hit ray (reference layout: 64 B / 72 B each)
ORC_API void orc_trace_counts(void *h, const rtb_ray *rays, int64_t n, double *pairs, double *tris) {
    const Scene &s = *(Scene *)h;
    Counters c;
    for (int64_t i = 0; i < n; ++i) {
        Ray r{V(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), V(rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]), rays[i].tmax};
        Isect is; int prim;
        traverse<false>(s, r, is, prim, -1, &c);
    }
    *pairs = (double)c.node_pairs / (double)n;
    *tris = (double)c.tri_tests / (double)n;
}
// accum_out (optional): radiance sums, 3 floats per pixel; rgb_out (optional): sqrt(sum/total_spp), render.cuh:330-338.
// pixel_begin/pixel_end bound the work (bench samples); stats = {paths, extend rays, shadow rays}
ORC_API void orc_render(void *h, const rtb_camera *cam, const rtb_render_params *p, int64_t pixel_begin, int64_t pixel_end,
                        float *rgb_out, float *accum_out, uint64_t *stats, int nthreads) {
    const Scene &s = *(Scene *)h;
    const int64_t npix = (int64_t)p->width * p->height;
    if (pixel_end > npix || pixel_end < 0) pixel_end = npix;
    std::vector<float> acc((size_t)npix * 3, 0.f);
    std::vector<RenderStats> st((size_t)std::max(1, nthreads));
    parallel_for(pixel_end - pixel_begin, nthreads, [&](int64_t k, int t) {
        const int64_t pix = pixel_begin + k;
        for (int sidx = 0; sidx < p->spp; ++sidx)
            trace_path(s, *cam, *p, (uint32_t)pix, (uint32_t)(p->first_sample + sidx), &acc[(size_t)pix * 3], st[(size_t)t]);
    });
    if (stats) {
        stats[0] = stats[1] = stats[2] = 0;
        for (auto &x : st) { stats[0] += x.paths; stats[1] += x.extend; stats[2] += x.shadow; }
    }
    if (accum_out) memcpy(accum_out, acc.data(), sizeof(float) * acc.size());
    if (rgb_out) {
        const float inv = 1.f / (float)(p->total_spp > 0 ? p->total_spp : p->spp);
        for (size_t i = 0; i < acc.size(); ++i) rgb_out[i] = sqrtf(acc[i] * inv);
    }
}
// ---- known-answer helpers for unit tests ----
ORC_API int orc_tri_intersect(const float v[9], const rtb_ray *r, float tuv[3]) {
    Triangle t(V(v[0], v[1], v[2]), V(v[3], v[4], v[5]), V(v[6], v[7], v[8]));
    Ray ray{V(r->origin[0], r->origin[1], r->origin[2]), V(r->dir[0], r->dir[1], r->dir[2]), r->tmax};
    Isect is{0, 0, 0};
    bool hit = t.intersect(ray, is);
    tuv[0] = is.t; tuv[1] = is.u; tuv[2] = is.v;
    return hit ? 1 : 0;
}
ORC_API void orc_offset_ray_origin(const float p[3], const float n[3], float out[3]) {
    Vec3 r = offset_ray_origin(V(p[0], p[1], p[2]), V(n[0], n[1], n[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
ORC_API void orc_sample_f(const rtb_material *m, const float wo[3], const float n_in[3], float u1, float u2, float f[3],
                          float n_out[3], float wi[3], float *pdf) {
    Vec3 n = V(n_in[0], n_in[1], n_in[2]), w;
    Vec3 r = sample_f(*m, V(wo[0], wo[1], wo[2]), u1, u2, n, w, *pdf);
    f[0] = r.x; f[1] = r.y; f[2] = r.z; n_out[0] = n.x; n_out[1] = n.y; n_out[2] = n.z; wi[0] = w.x; wi[1] = w.y; wi[2] = w.z;
}
// Light::sample_Li for an area light (light.cuh:38-46 with Triangle::sample_p / area, triangle.cuh:78-86): the same
// statements as in trace_path above, on a triangle given by its vertices
ORC_API void orc_sample_li_area(const float v[9], const float p[3], float u1, float u2, float wi[3], float *t, float *pdf) {
    Triangle lt(V(v[0], v[1], v[2]), V(v[3], v[4], v[5]), V(v[6], v[7], v[8]));
    const Vec3 P = V(p[0], p[1], p[2]);
    float area = 0.5f * length_dev(lt.n);
    float a = sqrtf(u1);
    Vec3 q = lt.p(1.f - a, u2 * a);
    Vec3 w = q - P;
    *t = length_dev(w);
    Vec3 wiL = w * (1.f / *t);
    *pdf = (1.f / area) * (dot_dev(w, w) / fabsf(dot_dev(unit_dev(lt.n), wiL)));
    wi[0] = wiL.x; wi[1] = wiL.y; wi[2] = wiL.z;
}
ORC_API void orc_rand4(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t block, float out[4]) {
    R4 r = rand4(seed, pixel, sample, block);
    out[0] = r.a; out[1] = r.b; out[2] = r.c; out[3] = r.d;
}
// camera.cuh:15-29, host arithmetic
ORC_API void orc_camera_look_at(const float from[3], const float at[3], const float up[3], float vfov, float aspect, rtb_camera *c) {
    float vfov_rad = vfov * (PI_F / 180.f);
    float vh = 2.f * tanf(vfov_rad * 0.5f), vw = vh * aspect;
    Vec3 lf = V(from[0], from[1], from[2]);
    Vec3 w = lf - V(at[0], at[1], at[2]);
    w = w * (1.f / sqrtf(dot_host(w, w)));
    Vec3 upv = V(up[0], up[1], up[2]);
    Vec3 v = upv - dot_host(upv, w) * w;
    v = v * (1.f / sqrtf(dot_host(v, v)));
    Vec3 u = cross_host(v, w);
    Vec3 hz = vw * u, vt = -vh * v;
    Vec3 ul = lf - w - 0.5f * hz - 0.5f * vt;
    float *o[4] = {c->lookfrom, c->upper_left, c->horizontal, c->vertical};
    Vec3 src[4] = {lf, ul, hz, vt};
    for (int k = 0; k < 4; ++k) { o[k][0] = src[k].x; o[k][1] = src[k].y; o[k][2] = src[k].z; }
}

}  // extern "C"
