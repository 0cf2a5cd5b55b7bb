// rtb_wavefront.h — ray queues and the bodies of the wavefront stage kernels
// (generate / extend / shade / shadow / control).
//
// Replaces the reference's pools RayPool / PathRayPayload / ShadowRayPayload
// (render.cuh:5-23), its flag arrays + cub::DeviceSelect compaction
// (render.cuh:348-364) and the stage kernels gen / ch / mat / ah / init
// (render.cuh:84-328).  Design for HBM3e:
//   - there are no path "slots": a path's state travels WITH its ray, stored
//     compacted at its queue position as three 16-byte words, so every stage
//     reads and writes whole 128-bit words at consecutive addresses
//     (the reference gathers 7 scalars per ray through an id indirection);
//   - the hit queues are filled by the extend kernel through warp-aggregated
//     atomics (ballot + prefix popcount); the ray queues are written by shade
//     at the path's own hit-queue position (holes where a path casts no ray):
//     no compaction pass, no flag arrays, no per-iteration device->host copy
//     of queue sizes;
//   - a finished path frees queue capacity at once and `generate` tops the
//     extend queue up with new camera paths every iteration (no lock-step
//     generations, SURVEY §3.3 Quirk B);
//   - hits are written into one queue per material type, so each shade launch
//     runs a single BSDF (material sort without a sort);
//   - per-path RNG is a counter hash (no 48-byte XORWOW state traffic).
// One single-thread control kernel per iteration does the queue bookkeeping.
// Bodies are RTB_HD: kernels in rtb_cuda.cu are thin wrappers, and tests/emu
// runs the same bodies sequentially on the CPU.
#pragma once
#include "rtb_shade.h"

namespace rtb {

constexpr int kNumMaterialTypes = 4;  // MATTE, MIRROR, GLASS, GLOSSY: one hit queue each

struct Counters {
    int32_t n_extend, n_shadow;        // entries of this iteration's ray queues (holes included), set by control
    int32_t n_mat[kNumMaterialTypes];  // hit queue sizes: pushed by extend, consumed by shade, reset by control
    int32_t extend_head, shadow_head;  // fetch cursors of the persistent kernels
    int32_t done, _pad;
    unsigned long long next_path, total_paths;
    unsigned long long stat_extend, stat_shadow, stat_paths, stat_iters, stat_hits;
    unsigned long long work[4];  // extend nodes, extend tris, shadow nodes, shadow tris (counting variants)
};

// Record layouts (each field array has `pool` entries, 16 bytes per entry):
//   extend queue   ea = origin.xyz | pixel      eb = dir.xyz | sample<<8|bounces   ec = beta.xyz | pdf of the BSDF sample
//   hit queue[t]   ma = dir.xyz    | pixel      mb = beta.xyz | sample<<8|bounces  mc = material word,u,v | leaf-order triangle
//   shadow queue   sh_o = origin.xyz | tmax     sh_d = dir.xyz | excluded triangle sh_L = radiance | pixel
// Hit queues are dense (extend appends to them with warp-aggregated atomics).  The two ray queues
// are NOT compacted: the path in slot i of the concatenated hit queues writes its next ray and its
// shadow ray to slot i of the ray queues, or a hole marker (pixel = kHolePixel / tmax = 0) when it
// casts none; `generate` appends new camera paths behind them.  ncu r1 on the compacting version:
// 52 % of the shade kernel's stall samples sat on the return of the two queue-append atomics.
struct WaveState {
    F4 *ea, *eb, *ec;
    F4 *ma, *mb, *mc;  // one block of `pool` entries per material type PRESENT in the scene; type t starts at qbase[t]
    F4 *sh_o, *sh_d, *sh_L;
    Counters *c;
    int32_t *host_done;  // mapped pinned host word (or null): lets the host poll without a stream sync
    F4 *accum;           // one 16-byte word per pixel: radiance sums in x, y, z — one vector reduction per splat
    unsigned long long *accum_fx;  // RTB_RENDER_DETERMINISTIC (null otherwise): three 64-bit fixed-point sums per pixel; integer
                                   // adds commute, so the image no longer depends on the order of the splats or on how the
                                   // samples were spread over wavefronts and GPUs
    int32_t pool;
    int32_t qbase[kNumMaterialTypes];  // first entry of type t's hit queue (slot of t among the present types x pool)
    // beyond the reference (off in parity mode): per-hit {pdf of the BSDF sample that produced the ray, hit
    // distance} for RTB_RENDER_TRUE_MIS ([kNumMaterialTypes][pool], null otherwise); constant environment radiance
    float *mis;
    float env[3];
    int32_t has_env;
    // two-level scenes: instance of every hit ([kNumMaterialTypes][pool], null for flat scenes)
    int32_t *hit_inst;
};

// ------------------------------------------------------------ framebuffer splat
// atomic_add(Vec3), vec3.cuh:149-153, is three scalar float atomics on a 12-byte pixel.  Here a pixel is one aligned
// 16-byte word and a splat is ONE vector reduction (red.global.add.v4.f32, sm_90+: no return value, nothing waits).
// Deterministic mode: 2^-28 fixed point in 64-bit integers; a splat is clamped to +-2^24 so that 2^11 clamped splats fit.
constexpr float kFixedScale = 268435456.f;              // 2^28
constexpr double kFixedInvScale = 1.0 / 268435456.0;
RTB_HD long long to_fixed(float v) {
    v = fminf(fmaxf(v, -16777216.f), 16777216.f);
#if defined(__CUDA_ARCH__)
    return __float2ll_rn(fmul(v, kFixedScale));
#else
    return llrintf(fmul(v, kFixedScale));
#endif
}
RTB_HD float from_fixed(long long v) { return (float)((double)v * kFixedInvScale); }
RTB_HD void accum_add(const WaveState &W, uint32_t pixel, V3 L) {
#if defined(__CUDA_ARCH__)
    if (W.accum_fx) {
        unsigned long long *p = W.accum_fx + 3 * (size_t)pixel;
        atomicAdd(p, (unsigned long long)to_fixed(L.x)); atomicAdd(p + 1, (unsigned long long)to_fixed(L.y));
        atomicAdd(p + 2, (unsigned long long)to_fixed(L.z));
    } else {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(W.accum + pixel), "f"(L.x), "f"(L.y), "f"(L.z), "f"(0.f) : "memory");
    }
#else
    if (W.accum_fx) {
        unsigned long long *p = W.accum_fx + 3 * (size_t)pixel;
        p[0] += (unsigned long long)to_fixed(L.x); p[1] += (unsigned long long)to_fixed(L.y); p[2] += (unsigned long long)to_fixed(L.z);
    } else {
        F4 &a = W.accum[pixel];
        a.x = fadd(a.x, L.x); a.y = fadd(a.y, L.y); a.z = fadd(a.z, L.z);
    }
#endif
}

constexpr int kMaxBounces = 255;       // bounces share a word with the sample index
constexpr int kMaxSampleIndex = 1 << 24;
constexpr uint32_t kHolePixel = 0xffffffffu;

// ------------------------------------------------------------ queue pushes
#if defined(__CUDA_ARCH__)
// warp-aggregated append: one atomic per warp, positions by ballot prefix
RTB_HD int queue_push(int32_t *counter) {
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}
// ray statistics of a shade launch: tallied per thread over its grid-stride loop, then one
// reduction per warp without a return value (RED: nothing waits for it)
struct ShadeTally { uint32_t extend, shadow; };
RTB_HD void tally_flush(Counters *c, const ShadeTally &t) {
    const unsigned e = __reduce_add_sync(0xffffffffu, t.extend), s = __reduce_add_sync(0xffffffffu, t.shadow);
    if ((threadIdx.x & 31) == 0) {
        if (e) atomicAdd(&c->stat_extend, (unsigned long long)e);
        if (s) atomicAdd(&c->stat_shadow, (unsigned long long)s);
    }
}
RTB_HD void work_add(unsigned long long *p, unsigned v) { atomicAdd(p, (unsigned long long)v); }
#else
RTB_HD int queue_push(int32_t *counter) { return (*counter)++; }
struct ShadeTally { uint32_t extend, shadow; };
RTB_HD void tally_flush(Counters *c, const ShadeTally &t) { c->stat_extend += t.extend; c->stat_shadow += t.shadow; }
RTB_HD void work_add(unsigned long long *p, unsigned v) { *p += v; }
#endif

RTB_HD F4 f4(V3 a, float w) { F4 r; r.x = a.x; r.y = a.y; r.z = a.z; r.w = w; return r; }
RTB_HD V3 xyz(F4 a) { return v3(a.x, a.y, a.z); }

// ------------------------------------------------------------ generate
// gen, render.cuh:250-275.  Tops the extend queue up to `pool` entries with
// new camera paths; pixel = path / spp (the samples of one pixel are
// consecutive ids, so one warp's primary rays are coherent).
RTB_HD int hit_total(const Counters &c) { return c.n_mat[0] + c.n_mat[1] + c.n_mat[2] + c.n_mat[3]; }
RTB_HD int generate_count(const WaveState &W) {
    const Counters &c = *W.c;
    const unsigned long long remaining = c.total_paths - c.next_path;
    const unsigned long long room = (unsigned long long)(W.pool - hit_total(c));
    return (int)(room < remaining ? room : remaining);
}
RTB_HD void generate_body(const WaveState &W, const RenderConsts &rc, int tid) {
    const Counters &c = *W.c;
    const unsigned long long path = (unsigned long long)rc.path_offset + (c.next_path + (unsigned long long)tid) * (unsigned long long)rc.path_stride;
    const uint32_t pixel = (uint32_t)(path / (unsigned long long)rc.spp);
    const uint32_t sample = (uint32_t)rc.first_sample + (uint32_t)(path % (unsigned long long)rc.spp);
    V3 o, d;
    generate_camera_ray(rc, pixel, sample, o, d);
    const int q = hit_total(c) + tid;  // behind the slots of the paths that were shaded this iteration
    W.ea[q] = f4(o, u2f(pixel));
    W.eb[q] = f4(d, u2f(sample << 8));
    W.ec[q] = f4(v3(1.f), 0.f);
}

// ------------------------------------------------------------ extend
// ch, render.cuh:297-328 (PATH_RAY part), one queue entry.  A miss ends the
// path (the reference parks the slot until max_bounces, Quirk B).
// append to the hit queue of one material type (the queues are [kNumMaterialTypes][pool]); the branches keep each
// warp-aggregated atomic among the lanes of one type
RTB_HD int hit_queue_push(const WaveState &W, int type) {
    if (type == RTB_MATTE) return W.qbase[0] + queue_push(&W.c->n_mat[0]);
    if (type == RTB_MIRROR) return W.qbase[1] + queue_push(&W.c->n_mat[1]);
    if (type == RTB_GLASS) return W.qbase[2] + queue_push(&W.c->n_mat[2]);
    return W.qbase[3] + queue_push(&W.c->n_mat[3]);
}
RTB_HD void extend_miss(const WaveState &W, uint32_t pixel, V3 beta) {  // environment light, rtb_render_params.env_L
    const V3 L = vmul(beta, v3(W.env[0], W.env[1], W.env[2]));
    if (finite3(L)) accum_add(W, pixel, L);
}
RTB_HD void extend_finish(const WaveState &W, const SceneView &S, int qi, const HitRec &h, int hit_inst = -1) {
    if (h.tri < 0) {
        if (W.has_env) extend_miss(W, f2u(ldg(W.ea + qi).w), xyz(ldg(W.ec + qi)));
        return;
    }
    const int mat = hit_material(S, h.tri, hit_inst);
    const int type = mat >> 24;
    const int j = hit_queue_push(W, type);
    // the shade kernel needs u, v and the triangle, not t: the slot carries the material word
    // (index | type << 24) instead, which saves shade one dependent load per hit
    F4 hr; hr.x = i2f(mat); hr.y = h.u; hr.z = h.v; hr.w = i2f(h.tri);
    const F4 a = ldg(W.ea + qi), b = ldg(W.eb + qi), beta = ldg(W.ec + qi);
    W.ma[j] = f4(xyz(b), a.w);
    W.mb[j] = f4(xyz(beta), b.w);
    W.mc[j] = hr;
    if (W.mis) { W.mis[2 * (size_t)j] = beta.w; W.mis[2 * (size_t)j + 1] = h.t; }
    if (W.hit_inst) W.hit_inst[j] = hit_inst;
}
template <bool COUNT>
RTB_HD void extend_body(const WaveState &W, const SceneView &S, int qi) {
    const F4 a = ldg(W.ea + qi);
    if (f2u(a.w) == kHolePixel) return;
    const F4 b = ldg(W.eb + qi);
    HitRec h;
    TraceCounters tc; tc.nodes = 0; tc.tris = 0;
    int32_t hi = -1;
    scene_trace<false, COUNT>(S.bvh, xyz(a), xyz(b), FLT_MAX, -1, -1, h, hi, &tc);
    if (COUNT) { work_add(&W.c->work[0], tc.nodes); work_add(&W.c->work[1], tc.tris); }
    extend_finish(W, S, qi, h, hi);
}

// ------------------------------------------------------------ shade
// init + mat (render.cuh:84-248) for entry `tid` of hit queue `type`
// one hit record (a, b, hr = its three words, already loaded) of hit queue TYPE, entry tid
template <int TYPE, bool EXT = true>
RTB_HD void shade_item(const WaveState &W, const SceneView &S, const RenderConsts &rc, bool shadows, int tid, const F4 &a,
                       const F4 &b, const F4 &hr, ShadeTally &tally) {
    const int q = W.qbase[TYPE] + tid;
    PathStepIn in;
    in.wo = xyz(a);
    in.hit.t = 0.f; in.hit.u = hr.y; in.hit.v = hr.z; in.hit.tri = f2i(hr.w);
    in.material = f2i(hr.x);
    in.beta = xyz(b);
    const uint32_t packed = f2u(b.w);
    in.bounces = (int)(packed & 0xffu);
    in.sample = packed >> 8;
    in.pixel = f2u(a.w);
    in.prev_pdf = 0.f; in.t = 0.f;
    if (EXT && W.mis) { in.prev_pdf = W.mis[2 * (size_t)q]; in.t = W.mis[2 * (size_t)q + 1]; }
    in.inst = (EXT && W.hit_inst) ? W.hit_inst[q] : -1;
    PathStepOut out;
    path_step<TYPE, EXT>(S, rc, in, out);
    if (out.emit) accum_add(W, in.pixel, out.emission);
    // A ray with a non-finite component can hit nothing (every comparison of the triangle test fails),
    // but it would pass every slab test and walk the whole tree: retire it here with the result it
    // would have had — a miss ends the path, an unoccluded shadow ray splats.  (MATTE sampling yields
    // normalize(0) = NaN with probability 2^-24 per bounce: a few rays per frame, each worth seconds
    // on a 10 M-triangle scene.)
    if (out.extend && !(finite3(out.o) && finite3(out.d))) out.extend = false;
    if (out.shadow && !(finite3(out.so) && finite3(out.sd) && out.stmax > 0.f)) {
        out.shadow = false;
        if (finite3(out.sL)) accum_add(W, in.pixel, out.sL);
    }
    // slot of this path in the ray queues = its position in the concatenated hit queues
    const Counters &c = *W.c;
    const int j = tid + (TYPE > 0 ? c.n_mat[0] : 0) + (TYPE > 1 ? c.n_mat[1] : 0) + (TYPE > 2 ? c.n_mat[2] : 0);
    if (out.extend) {
        W.ea[j] = f4(out.o, a.w);
        W.eb[j] = f4(out.d, u2f((in.sample << 8) | (uint32_t)out.bounces));
        W.ec[j] = f4(out.beta, out.pdf);
    } else {
        W.ea[j] = f4(v3(0.f), u2f(kHolePixel));
    }
    if (shadows) {
        if (out.shadow) {
            W.sh_o[j] = f4(out.so, out.stmax);
            W.sh_d[j] = f4(out.sd, i2f(out.sexcl));
            W.sh_L[j] = f4(out.sL, a.w);
        } else {
            W.sh_o[j] = f4(v3(0.f), 0.f);  // tmax = 0: hole
        }
    }
    tally.extend += out.extend ? 1u : 0u;
    tally.shadow += out.shadow ? 1u : 0u;
}
template <int TYPE, bool EXT = true>
RTB_HD void shade_body(const WaveState &W, const SceneView &S, const RenderConsts &rc, bool shadows, int tid, ShadeTally &tally) {
    const int q = W.qbase[TYPE] + tid;
    const F4 a = ldg(W.ma + q), b = ldg(W.mb + q), hr = ldg(W.mc + q);
    shade_item<TYPE, EXT>(W, S, rc, shadows, tid, a, b, hr, tally);
}

// ------------------------------------------------------------ shadow
// ah, render.cuh:278-294, one queue entry
RTB_HD void shadow_finish(const WaveState &W, int si, bool occluded) {
    if (occluded) return;
    const F4 l = ldg(W.sh_L + si);
    const V3 L = xyz(l);
    if (finite3(L)) accum_add(W, f2u(l.w), L);
}
template <bool COUNT>
RTB_HD void shadow_body(const WaveState &W, const SceneView &S, int si) {
    const F4 o = ldg(W.sh_o + si);
    if (!(o.w > 0.f)) return;  // hole (a ray with tmax <= 0 can hit nothing and would splat unoccluded)
    const F4 d = ldg(W.sh_d + si);
    HitRec h;
    TraceCounters tc; tc.nodes = 0; tc.tris = 0;
    int32_t hi = -1;  // (the excluded triangle is an emitter: its mesh is instanced once, the triangle index identifies it)
    const bool occluded = scene_trace<true, COUNT>(S.bvh, xyz(o), xyz(d), o.w, f2i(d.w), -1, h, hi, &tc);
    if (COUNT) { work_add(&W.c->work[2], tc.nodes); work_add(&W.c->work[3], tc.tris); }
    shadow_finish(W, si, occluded);
}

// ------------------------------------------------------------ control
// one thread, between (shade, generate) and the traversal kernels: size the ray queues of this
// iteration, account the new camera paths, reset what shade consumed, arm the fetch cursors, raise `done`
RTB_HD void control_body(const WaveState &W, bool shadows) {
    Counters &c = *W.c;
    const int nh = hit_total(c);
    const int started = generate_count(W);
    c.next_path += (unsigned long long)started;
    c.stat_paths += (unsigned long long)started;
    c.stat_extend += (unsigned long long)started;
    c.stat_hits += (unsigned long long)nh;
    c.n_extend = nh + started;
    c.n_shadow = shadows ? nh : 0;
    c.n_mat[0] = c.n_mat[1] = c.n_mat[2] = c.n_mat[3] = 0;
    c.extend_head = 0;
    c.shadow_head = 0;
    if (c.n_extend == 0) {
        c.done = 1;
        if (W.host_done) *W.host_done = 1;
    } else {
        c.stat_iters += 1;
    }
}

}  // namespace rtb
