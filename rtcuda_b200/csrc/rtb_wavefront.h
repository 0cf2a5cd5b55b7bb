// rtb_wavefront.h — path pool, ray queues and the bodies of the wavefront
// stage kernels (generate / extend / shade / shadow / control).
//
// Replaces the reference's pools RayPool / PathRayPayload / ShadowRayPayload
// (render.cuh:5-23), its flag arrays + cub::DeviceSelect compaction
// (render.cuh:348-364) and the stage kernels gen / ch / mat / ah / init
// (render.cuh:84-328).  Differences that matter for speed:
//   - queues are filled by the producing kernel through warp-aggregated
//     atomics (ballot + prefix popcount), no separate compaction pass and no
//     per-iteration device->host copy of queue sizes;
//   - a path slot is refilled with a new camera path the moment its path
//     ends (no lock-step generations, SURVEY §3.3 Quirk B);
//   - hits are queued per material type, so each shade launch runs one BSDF;
//   - shadow rays live compacted at their queue position (coalesced reads);
//   - per-path RNG is a counter hash (no 48-byte XORWOW state traffic).
// Bodies are RTB_HD: kernels in rtb_kernels.cu are thin grid-stride wrappers,
// and tests/emu runs the same bodies sequentially on the CPU.
#pragma once
#include "rtb_shade.h"

namespace rtb {

struct Counters {
    int32_t n_extend, n_shadow, n_free;
    int32_t n_mat[3];
    int32_t extend_head, shadow_head;  // fetch cursors of the persistent kernels
    int32_t done, _pad;
    unsigned long long next_path, total_paths;
    unsigned long long stat_extend, stat_shadow, stat_paths, stat_iters;
    unsigned long long work[4];  // extend nodes, extend tris, shadow nodes, shadow tris (counting variants)
};

struct WaveState {
    // path state, indexed by pool slot
    F4 *ray_o;  // origin.xyz, -
    F4 *ray_d;  // dir.xyz, -
    F4 *hit;    // t, u, v, leaf-order triangle (int bits)
    F4 *beta;   // beta.xyz, bounces (int bits)
    uint32_t *pixel, *sample;
    // shadow rays, indexed by shadow-queue position
    F4 *sh_o;   // origin.xyz, tmax
    F4 *sh_d;   // dir.xyz, excluded leaf-order triangle (int bits)
    F4 *sh_L;   // radiance to add on a miss, pixel (int bits)
    int32_t *extend_q;  // slots with a ray to extend            [pool]
    int32_t *mat_q;     // slots with a hit, per material type   [3][pool]
    int32_t *free_q;    // slots whose path ended                [pool]
    Counters *c;
    float *accum;  // 3 floats per pixel, radiance sums
    int32_t pool;
};

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
RTB_HD void accum_add(float *accum, uint32_t pixel, V3 L) {
    float *p = accum + 3 * (size_t)pixel;
    atomicAdd(p, L.x); atomicAdd(p + 1, L.y); atomicAdd(p + 2, L.z);
}
#else
RTB_HD int queue_push(int32_t *counter) { return (*counter)++; }
RTB_HD void accum_add(float *accum, uint32_t pixel, V3 L) {
    float *p = accum + 3 * (size_t)pixel;
    p[0] += L.x; p[1] += L.y; p[2] += L.z;
}
#endif

RTB_HD F4 f4(V3 a, float w) { F4 r; r.x = a.x; r.y = a.y; r.z = a.z; r.w = w; return r; }
RTB_HD V3 xyz(F4 a) { return v3(a.x, a.y, a.z); }

// ------------------------------------------------------------ generate
// gen, render.cuh:250-275.  Thread i takes free slot i and camera path
// next_path + i; pixel = path / spp (the samples of one pixel are
// consecutive ids, so one warp's primary rays are coherent).
RTB_HD void generate_body(const WaveState &W, const RenderConsts &rc, int tid) {
    const Counters &c = *W.c;
    if (tid >= c.n_free) return;
    const unsigned long long path = c.next_path + (unsigned long long)tid;
    if (path >= c.total_paths) return;
    const int slot = W.free_q[tid];
    const uint32_t pixel = (uint32_t)(path / (unsigned long long)rc.spp);
    const uint32_t sample = (uint32_t)rc.first_sample + (uint32_t)(path % (unsigned long long)rc.spp);
    V3 o, d;
    generate_camera_ray(rc, pixel, sample, o, d);
    W.ray_o[slot] = f4(o, 0.f);
    W.ray_d[slot] = f4(d, 0.f);
    W.beta[slot] = f4(v3(1.f), i2f(0));
    W.pixel[slot] = pixel;
    W.sample[slot] = sample;
    W.extend_q[c.n_extend + tid] = slot;  // n_extend is advanced by control_a
}

// ------------------------------------------------------------ extend
RTB_HD void extend_finish(const WaveState &W, const SceneView &S, int slot, const HitRec &h) {
    if (h.tri < 0) {
        W.free_q[queue_push(&W.c->n_free)] = slot;
        return;
    }
    F4 hr; hr.x = h.t; hr.y = h.u; hr.z = h.v; hr.w = i2f(h.tri);
    W.hit[slot] = hr;
    const int type = S.tri_meta[h.tri].material >> 24;
    if (type == RTB_MATTE) W.mat_q[queue_push(&W.c->n_mat[0])] = slot;
    else if (type == RTB_MIRROR) W.mat_q[W.pool + queue_push(&W.c->n_mat[1])] = slot;
    else W.mat_q[2 * W.pool + queue_push(&W.c->n_mat[2])] = slot;
}

// ch, render.cuh:297-328 (PATH_RAY part), one queue entry
#if defined(__CUDA_ARCH__)
RTB_HD void work_add(unsigned long long *p, unsigned v) { atomicAdd(p, (unsigned long long)v); }
#else
RTB_HD void work_add(unsigned long long *p, unsigned v) { *p += v; }
#endif
template <bool COUNT>
RTB_HD void extend_body(const WaveState &W, const SceneView &S, int qi) {
    const int slot = W.extend_q[qi];
    const V3 o = xyz(W.ray_o[slot]), d = xyz(W.ray_d[slot]);
    HitRec h;
    TraceCounters tc; tc.nodes = 0; tc.tris = 0;
    bvh8_trace<false, COUNT>(S.bvh, o, d, FLT_MAX, -1, h, &tc);
    if (COUNT) { work_add(&W.c->work[0], tc.nodes); work_add(&W.c->work[1], tc.tris); }
    extend_finish(W, S, slot, h);
}

// ------------------------------------------------------------ shade
// init + mat (render.cuh:84-248) for one slot of material queue `type`
RTB_HD void shade_body(const WaveState &W, const SceneView &S, const RenderConsts &rc, int type, int tid) {
    if (tid >= W.c->n_mat[type]) return;
    const int slot = W.mat_q[type * W.pool + tid];
    PathStepIn in;
    const F4 hr = W.hit[slot], bt = W.beta[slot];
    in.wo = xyz(W.ray_d[slot]);
    in.hit.t = hr.x; in.hit.u = hr.y; in.hit.v = hr.z; in.hit.tri = f2i(hr.w);
    in.beta = xyz(bt);
    in.bounces = f2i(bt.w);
    in.pixel = W.pixel[slot];
    in.sample = W.sample[slot];
    PathStepOut out;
    path_step(S, rc, in, out);
    if (out.emit) accum_add(W.accum, in.pixel, out.emission);
    if (out.extend) {
        W.ray_o[slot] = f4(out.o, 0.f);
        W.ray_d[slot] = f4(out.d, 0.f);
        W.beta[slot] = f4(out.beta, i2f(out.bounces));
        W.extend_q[queue_push(&W.c->n_extend)] = slot;
    } else {
        W.free_q[queue_push(&W.c->n_free)] = slot;
    }
    if (out.shadow) {
        const int si = queue_push(&W.c->n_shadow);
        W.sh_o[si] = f4(out.so, out.stmax);
        W.sh_d[si] = f4(out.sd, i2f(out.sexcl));
        W.sh_L[si] = f4(out.sL, i2f((int)in.pixel));
    }
}

// ------------------------------------------------------------ shadow
RTB_HD void shadow_finish(const WaveState &W, int si, bool occluded) {
    if (occluded) return;
    const F4 l = W.sh_L[si];
    const V3 L = xyz(l);
    if (finite3(L)) accum_add(W.accum, (uint32_t)f2i(l.w), L);
}
// ah, render.cuh:278-294, one queue entry
template <bool COUNT>
RTB_HD void shadow_body(const WaveState &W, const SceneView &S, int si) {
    const F4 o = W.sh_o[si], d = W.sh_d[si];
    HitRec h;
    TraceCounters tc; tc.nodes = 0; tc.tris = 0;
    const bool occluded = bvh8_trace<true, COUNT>(S.bvh, xyz(o), xyz(d), o.w, f2i(d.w), h, &tc);
    if (COUNT) { work_add(&W.c->work[2], tc.nodes); work_add(&W.c->work[3], tc.tris); }
    shadow_finish(W, si, occluded);
}

// ------------------------------------------------------------ control
// after shade + generate: account the new camera paths, reset what shade and
// generate consumed, arm the fetch cursors, raise `done` when nothing is left
RTB_HD void control_a_body(const WaveState &W) {
    Counters &c = *W.c;
    unsigned long long remaining = c.total_paths - c.next_path;
    unsigned long long started = (unsigned long long)c.n_free < remaining ? (unsigned long long)c.n_free : remaining;
    c.next_path += started;
    c.stat_paths += started;
    c.n_extend += (int32_t)started;
    c.n_free = 0;
    c.n_mat[0] = c.n_mat[1] = c.n_mat[2] = 0;
    c.extend_head = 0;
    c.shadow_head = 0;
    c.stat_extend += (unsigned long long)c.n_extend;
    c.stat_shadow += (unsigned long long)c.n_shadow;
    c.stat_iters += 1;
    if (c.n_extend == 0 && c.n_shadow == 0) c.done = 1;
}
// after extend + shadow
RTB_HD void control_b_body(const WaveState &W) {
    Counters &c = *W.c;
    c.n_extend = 0;
    c.n_shadow = 0;
}

}  // namespace rtb
