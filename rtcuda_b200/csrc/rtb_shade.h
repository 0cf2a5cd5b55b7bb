// rtb_shade.h — materials, lights, camera and the per-path state machine.
//
// Restates, per path instead of per pool slot, what the reference's stage
// kernels compute: `init` (render.cuh:84-137: emission at bounce 0, depth
// cut, Russian roulette), `mat` (render.cuh:139-248: BSDF sample, light
// pick, next-event estimation with the MIS weight) and `gen`
// (render.cuh:250-275).  Image-affecting quirks of the reference are kept on
// purpose:
//   A  a Russian-roulette "kill" only pauses the path for one depth level;
//   C  power_heuristic(float, int) truncates the BSDF pdf, so the light
//      sample weight is exactly 1 for MATTE;
//   D  the BSDF-sampled MIS ray targets the shading triangle itself and can
//      never contribute, so it is not traced at all here;
//   -  emission is added for camera rays only; NEE exists for MATTE only;
//      area lights emit from both faces.
#pragma once
#include "rtb_bvh8.h"
#include "rtb.h"

namespace rtb {

struct TriMeta {
    int32_t material;  // index into materials (low 24 bits) | material type << 24
    int32_t light;     // index into lights or -1
};

struct LightDev {
    int32_t type;
    float px, py, pz;
    int32_t tri;  // leaf-order triangle index of an area light, -1 for point lights
    float Lx, Ly, Lz;
};

struct SceneView {
    Bvh8View bvh;
    const TriMeta *tri_meta;  // leaf order
    const rtb_material *materials;
    const LightDev *lights;
    int32_t num_lights;
    int32_t num_materials;
};

// material word (index | type << 24) of a hit: the instance's material if it has one, else the triangle's own
RTB_HD int hit_material(const SceneView &S, int tri, int inst) {
    if (inst >= 0) {
        const int m = inst_info(S.bvh.inst, inst).material;
        if (m >= 0) return m;
    }
    return S.tri_meta[tri].material;
}

struct RenderConsts {
    rtb_camera cam;
    int32_t width, height;
    int32_t spp, first_sample;
    int32_t max_bounces, rr_start;
    float rr_threshold;
    uint32_t seed;
    int32_t flags;
    int32_t path_offset, path_stride;  // this wavefront renders paths offset, offset + stride, ...
    float env[3];                      // constant environment radiance (rtb_render_params.env_L)
};

// ------------------------------------------------------------------ camera
// Camera::get_ray, camera.cuh:31-34
RTB_HD void camera_ray(const rtb_camera &c, float x, float y, V3 &o, V3 &d) {
    V3 ul = v3(c.upper_left[0], c.upper_left[1], c.upper_left[2]);
    V3 h = v3(c.horizontal[0], c.horizontal[1], c.horizontal[2]);
    V3 v = v3(c.vertical[0], c.vertical[1], c.vertical[2]);
    o = v3(c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]);
    V3 dir = vsub(vmad(vmad(ul, x, h), y, v), o);
    d = vnormalize(dir);
}

// gen, render.cuh:250-275: pixel = id / spp, jitter from dimension block 0
RTB_HD void generate_camera_ray(const RenderConsts &rc, uint32_t pixel, uint32_t sample, V3 &o, V3 &d) {
    int i = (int)(pixel % (uint32_t)rc.width), j = (int)(pixel / (uint32_t)rc.width);
    float jx = 0.5f, jy = 0.5f;
    if (!(rc.flags & RTB_RENDER_PIXEL_CENTRE)) {
        Rand4 r = rand4(rc.seed, pixel, sample, 0u);
        jx = r.a; jy = r.b;
    }
    float x = fdiv(fadd((float)i, jx), (float)rc.width);
    float y = fdiv(fadd((float)j, jy), (float)rc.height);
    camera_ray(rc.cam, x, y, o, d);
}

// ------------------------------------------------------------------ BSDFs
RTB_HD V3 reflect(V3 v, V3 n) {  // vec3.cuh:71-73
    return vsub(v, vscale(n, fmul(2.f, vdot(v, n))));
}

RTB_HD V3 uniform_sample_sphere(float u1, float u2) {  // utility.cuh:70-77
    float z = fsub(1.f, fmul(2.f, u1));
    float r = fsqrt(fmaxf(0.f, fsub(1.f, fmul(z, z))));
    float phi = fmul(kTwoPi, u2);
    float s, c;
    sincosf(phi, &s, &c);
    return v3(fmul(r, c), fmul(r, s), z);
}

// orthonormal basis around a unit vector (Duff et al. 2017)
RTB_HD void onb(V3 n, V3 &t, V3 &b) {
    const float sg = copysignf(1.f, n.z);
    const float a = fdiv(-1.f, fadd(sg, n.z));
    const float c = fmul(fmul(n.x, n.y), a);
    t = v3(ffma(fmul(sg, fmul(n.x, n.x)), a, 1.f), fmul(sg, c), fmul(-sg, n.x));
    b = v3(c, ffma(fmul(n.y, n.y), a, sg), -n.y);
}
// BRDF value and solid-angle pdf of RTB_GLOSSY for a given pair of directions (n on the side of wi)
RTB_HD void glossy_eval(const rtb_material &m, V3 wo, V3 n, V3 wi, V3 &f, float &pdf) {
    const float ca = vdot(reflect(wo, n), wi);
    if (ca <= 0.f) { f = v3(0.f); pdf = 0.f; return; }
    const float lobe = fmul(powf(ca, m.ior), 0.15915494309189535f);
    pdf = fmul(fadd(m.ior, 1.f), lobe);
    f = vscale(v3(m.albedo[0], m.albedo[1], m.albedo[2]), fmul(fadd(m.ior, 2.f), lobe));
}

struct BsdfSample {
    V3 f, n, wi;
    float pdf;
};

// Material::sample_f, material.cuh:60-109.  `n` comes in as the geometric
// normal and leaves flipped into the hemisphere of wi.  wo is the INCOMING
// ray direction (points into the surface, render.cuh:146).
RTB_HD BsdfSample sample_f(const rtb_material &m, V3 wo, V3 n, float u1, float u2) {
    BsdfSample s;
    V3 albedo = v3(m.albedo[0], m.albedo[1], m.albedo[2]);
    if (m.type == RTB_MATTE || m.type == RTB_MIRROR) {
        if (vdot(wo, n) > 0.f) n = vneg(n);
        if (m.type == RTB_MATTE) {
            s.wi = vnormalize(vadd(n, uniform_sample_sphere(u1, u2)));
            s.pdf = fmul(vdot(s.wi, n), kInvPi);
            s.f = vscale(albedo, kInvPi);
        } else {
            s.wi = reflect(wo, n);
            s.pdf = 1.f;
            s.f = vscale(albedo, frcp(vdot(s.wi, n)));
        }
        s.n = n;
        return s;
    }
    if (m.type == RTB_GLOSSY) {  // not in the reference: Phong lobe around the mirror direction
        if (vdot(wo, n) > 0.f) n = vneg(n);
        const V3 r = reflect(wo, n);
        const float e = m.ior;
        const float ca = powf(u1, frcp(fadd(e, 1.f)));
        const float sa = fsqrt(fmaxf(0.f, fsub(1.f, fmul(ca, ca))));
        float sp, cp;
        sincosf(fmul(kTwoPi, u2), &sp, &cp);
        V3 t, b;
        onb(r, t, b);
        s.wi = vadd(vadd(vscale(t, fmul(sa, cp)), vscale(b, fmul(sa, sp))), vscale(r, ca));
        const float lobe = fmul(powf(ca, e), 0.15915494309189535f);  // cos^e(alpha) / 2 pi
        s.pdf = fmul(fadd(e, 1.f), lobe);
        s.f = vscale(albedo, fmul(fadd(e, 2.f), lobe));
        s.n = n;
        return s;
    }
    // GLASS
    float cos_theta = vdot(wo, n);
    bool front = cos_theta < 0.f;
    if (front) cos_theta = -cos_theta;
    float inv_cos = frcp(cos_theta);
    float eta = front ? frcp(m.ior) : m.ior;
    float sin_theta = fsqrt(fsub(1.f, fmul(cos_theta, cos_theta)));
    if (!front) n = vneg(n);
    if (fmul(eta, sin_theta) > 1.f) {  // total internal reflection
        s.wi = reflect(wo, n);
        s.pdf = 1.f;
        s.f = v3(inv_cos);
        s.n = n;
        return s;
    }
    float r0 = fdiv(fsub(1.f, m.ior), fadd(1.f, m.ior));
    r0 = fmul(r0, r0);
    float reflectance = ffma(fsub(1.f, r0), powf(fsub(1.f, cos_theta), 5.f), r0);  // Schlick
    if (u1 < reflectance) {
        s.wi = reflect(wo, n);
        s.pdf = reflectance;
        s.f = v3(fmul(reflectance, inv_cos));
        s.n = n;
    } else {
        // refract, vec3.cuh:82-86 (cos_theta passed in)
        V3 v_par = vscale(vmad(wo, cos_theta, n), eta);
        V3 v_perp = vscale(n, -fsqrt(fsub(1.f, vlen2(v_par))));
        s.wi = vadd(v_par, v_perp);
        s.n = vneg(n);
        s.pdf = fsub(1.f, reflectance);
        s.f = v3(fdiv(fmul(fmul(s.pdf, eta), eta), vdot(s.wi, s.n)));
    }
    return s;
}

// ------------------------------------------------------------------ lights
struct LightSample {
    V3 wi, Li;
    float t, pdf;
};
// Light::sample_Li, light.cuh:29-48 with Triangle::sample_p / area, triangle.cuh:78-86
// the area-light branch on a triangle record the caller holds (also evaluated on its own by rtb_kat_eval)
RTB_HD LightSample sample_Li_area(const LightDev &l, const Tri48 &tr, V3 p, float u1, float u2);
RTB_HD LightSample sample_Li(const LightDev &l, const Bvh8View &bvh, V3 p, float u1, float u2) {
    if (l.type == RTB_POINT_LIGHT) {
        LightSample s;
        V3 w = vsub(v3(l.px, l.py, l.pz), p);
        s.t = vlen(w);
        s.Li = vscale(v3(l.Lx, l.Ly, l.Lz), frcp(fmul(s.t, s.t)));
        s.wi = vscale(w, frcp(s.t));
        s.pdf = 1.f;
        return s;
    }
    return sample_Li_area(l, load_tri(bvh.tris, l.tri), p, u1, u2);
}
RTB_HD LightSample sample_Li_area(const LightDev &l, const Tri48 &tr, V3 p, float u1, float u2) {
    LightSample s;
    V3 n = tri_n(tr);
    float area = fmul(0.5f, vlen(n));
    float pdf = frcp(area);
    float a = fsqrt(u1);
    float bu = fsub(1.f, a), bv = fmul(u2, a);
    V3 q = vmad(vmad(tri_p0(tr), -bu, tri_e1(tr)), bv, tri_e2(tr));  // Triangle::p(u,v), triangle.cuh:15
    V3 w = vsub(q, p);
    s.t = vlen(w);
    s.wi = vscale(w, frcp(s.t));
    s.Li = v3(l.Lx, l.Ly, l.Lz);
    s.pdf = fmul(pdf, fdiv(vlen2(w), fabsf(vdot(vnormalize(n), s.wi))));
    return s;
}

// power_heuristic(float f_pdf, int g_pdf), utility.cuh:52-55 — the int
// parameter is the reference's Quirk C and is reproduced literally.
RTB_HD float power_heuristic_ref(float f_pdf, float g_pdf_float) {
    int g = (int)g_pdf_float;
    float f2 = fmul(f_pdf, f_pdf);
    return fdiv(f2, fadd(f2, (float)(g * g)));
}

RTB_HD bool finite3(V3 a) {
    return fabsf(a.x) <= FLT_MAX && fabsf(a.y) <= FLT_MAX && fabsf(a.z) <= FLT_MAX;
}

// ------------------------------------------------------------------ path step
struct PathStepIn {
    V3 wo;  // direction of the ray that produced the hit
    HitRec hit;  // hit.tri >= 0
    V3 beta;
    uint32_t pixel, sample;
    int32_t bounces;
    int32_t material;  // TriMeta::material of the hit triangle (index | type << 24)
    float prev_pdf;    // RTB_RENDER_TRUE_MIS: solid-angle pdf of the BSDF sample that produced this ray (0 = delta)
    float t;           // RTB_RENDER_TRUE_MIS: hit distance
    int32_t inst;      // two-level scenes: instance of the hit triangle (-1 otherwise); read by the EXT kernels only
};
struct PathStepOut {
    bool emit; V3 emission;
    bool extend; V3 o, d, beta; int32_t bounces; float pdf;  // pdf: of the BSDF sample, 0 for MIRROR / GLASS
    bool shadow; V3 so, sd, sL; float stmax; int32_t sexcl;
};

// One visit of a path that HIT something: everything the reference does to it
// between two closest-hit traversals.  MT >= 0 tells the compiler the material
// type of the hit (the per-type shade kernels), -1 = read it from the material.
// EXT = false compiles the beyond-the-reference branches (RTB_RENDER_TRUE_MIS / RR_TERMINATE) out of the parity kernels.
template <int MT, bool EXT = true>
RTB_HD void path_step(const SceneView &S, const RenderConsts &rc, const PathStepIn &in, PathStepOut &out) {
    out.emit = false; out.extend = false; out.shadow = false;
    // issue every load that only depends on the hit before the roulette logic: the shade kernel is
    // bound by memory latency (ncu r1: long-scoreboard stalls dominate), not by instruction count
    // (two-level scenes run the EXT kernels: the hit triangle is taken to world space, everything below is unchanged)
    const Tri48 tr = EXT ? load_tri_world(S.bvh, in.hit.tri, in.inst) : load_tri(S.bvh.tris, in.hit.tri);
    rtb_material m = S.materials[in.material & 0xffffff];  // type is packed in the top byte
    if (MT >= 0) m.type = MT;
    int b = in.bounces;
    V3 beta = in.beta;
    // init, render.cuh:98-107: only camera rays see emitters
    const bool true_mis = EXT && (rc.flags & RTB_RENDER_TRUE_MIS) != 0;
    if (b == 0 || true_mis) {
        const int light = S.tri_meta[in.hit.tri].light;
        if (light >= 0) {
            const LightDev &l = S.lights[light];
            out.emit = true;
            out.emission = v3(l.Lx, l.Ly, l.Lz);
            if (b > 0) {
                // RTB_RENDER_TRUE_MIS: this path ray was the BSDF sample of the previous bounce; weigh it against
                // the light sample that could have produced the same direction (Light::pdf_Li, light.cuh:50-64,
                // times the 1/num_lights of the uniform light pick); a delta BSDF has no competitor
                float w = 1.f;
                if (in.prev_pdf > 0.f) {
                    const float area = fmul(0.5f, vlen(tri_n(tr)));
                    const float cosl = fabsf(vdot(vnormalize(tri_n(tr)), in.wo));
                    const float pdf_l = fdiv(fdiv(fmul(in.t, in.t), fmul(area, cosl)), (float)S.num_lights);
                    const float a2 = fmul(in.prev_pdf, in.prev_pdf);
                    w = fdiv(a2, fadd(a2, fmul(pdf_l, pdf_l)));
                }
                out.emission = vmul(vscale(out.emission, w), in.beta);
                if (!finite3(out.emission)) out.emit = false;
            }
        }
    }
    // init, render.cuh:109-126: depth cut + Russian roulette (Quirk A: a kill
    // costs one depth level and the roulette is rolled again on the same hit)
    while (true) {
        if (b >= rc.max_bounces) return;
        if (b > rc.rr_start) {
            float bm = vmax(beta);
            if (bm < rc.rr_threshold) {
                float p = fmaxf(0.05f, fsub(1.f, bm));
                float u = rand4(rc.seed, in.pixel, in.sample, 2u * (uint32_t)b + 1u).a;
                if (u < p) {
                    if (EXT && (rc.flags & RTB_RENDER_RR_TERMINATE)) return;
                    b++; continue;
                }
                beta = vscale(beta, frcp(fsub(1.f, p)));
            }
        }
        break;
    }
    const Rand4 xi = rand4(rc.seed, in.pixel, in.sample, 2u * (uint32_t)b + 2u);
    b++;
    // mat, render.cuh:139-168
    const V3 P = vmad(vmad(tri_p0(tr), -in.hit.u, tri_e1(tr)), in.hit.v, tri_e2(tr));
    const V3 ng = vneg(vnormalize(tri_n(tr)));
    const V3 beta_old = beta;
    {
        BsdfSample bs = sample_f(m, in.wo, ng, xi.a, xi.b);
        beta = vmul(beta, vscale(vscale(bs.f, vdot(bs.wi, bs.n)), frcp(bs.pdf)));
        out.o = offset_ray_origin(P, bs.n);
        out.d = bs.wi;
        out.beta = beta;
        out.bounces = b;
        out.pdf = (m.type == RTB_MATTE || m.type == RTB_GLOSSY) ? bs.pdf : 0.f;
        // the reference traces this ray even when the depth cut will discard
        // its result (render.cuh:109); skipping it does not change the image
        out.extend = b < rc.max_bounces;
        if (m.type == RTB_GLOSSY && !(vdot(bs.wi, bs.n) > 0.f && bs.pdf > 0.f)) out.extend = false;  // lobe sample below the surface
    }
    // mat, render.cuh:170-211: one uniformly picked light, MATTE only (get_f)
    if (S.num_lights == 0 || (rc.flags & RTB_RENDER_NO_SHADOW)) return;
    int li = (int)fmul(xi.c, (float)S.num_lights);
    if (li > S.num_lights - 1) li = S.num_lights - 1;
    const LightDev l = S.lights[li];
    const Rand4 xl = rand4(rc.seed, in.pixel, in.sample, 0x80000000u + (uint32_t)b);
    LightSample ls = sample_Li(l, S.bvh, P, xi.d, xl.a);
    V3 nl = vdot(ng, ls.wi) > 0.f ? ng : vneg(ng);
    if ((m.type == RTB_MATTE || m.type == RTB_GLOSSY) && fmul(vdot(in.wo, nl), vdot(ls.wi, nl)) < 0.f) {  // same_hemisphere, utility.cuh:57-59
        float cosl = vdot(ls.wi, nl);
        V3 f = vscale(vscale(v3(m.albedo[0], m.albedo[1], m.albedo[2]), kInvPi), cosl);
        float scattering_pdf = fmul(cosl, kInvPi);
        if (m.type == RTB_GLOSSY) {
            glossy_eval(m, in.wo, nl, ls.wi, f, scattering_pdf);
            f = vscale(f, cosl);
        }
        V3 mult = vscale(beta_old, (float)S.num_lights);
        V3 L = vmul(vmul(mult, f), ls.Li);
        if (l.type != RTB_POINT_LIGHT) {
            if (true_mis) {  // both pdfs in solid angle, the light's including the 1/num_lights of the pick
                const float pl = fdiv(ls.pdf, (float)S.num_lights);
                const float a2 = fmul(pl, pl);
                L = vscale(L, fdiv(a2, fadd(a2, fmul(scattering_pdf, scattering_pdf))));
            } else if (m.type == RTB_MATTE) {
                L = vscale(L, power_heuristic_ref(ls.pdf, scattering_pdf));
            }  // RTB_GLOSSY without RTB_RENDER_TRUE_MIS: light sampling only, weight 1 (what Quirk C makes of MATTE)
        }
        L = vscale(L, frcp(ls.pdf));
        out.shadow = true;
        out.so = offset_ray_origin(P, nl);
        out.sd = ls.wi;
        out.stmax = ls.t;
        out.sexcl = l.tri;
        out.sL = L;
    }
    // mat, render.cuh:213-245 (BSDF-sampled MIS ray): Quirk D — never
    // contributes, consumes no dimensions of the counter-based RNG.
}


}  // namespace rtb
