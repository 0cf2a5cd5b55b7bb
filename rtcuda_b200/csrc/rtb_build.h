// rtb_build.h — bodies of the GPU BVH builder kernels.
//
// Replaces Bvh::Bvh (bvh.cuh:30-219: single-threaded host full-sweep SAH over
// three std::sort'ed index arrays, ~260 ms for 69k triangles, minutes for
// 10M).  Pipeline, all on the device:
//   1. prim_setup     per-triangle record {p0,e1,e2,n}, bounds, scene bounds
//   2. morton         63-bit Morton code of the centroid, radix sort
//   3. PLOC           parallel locally-ordered clustering (Meister & Bittner
//                     2018): nearest neighbour within a window of the Morton
//                     order, merge mutual pairs, compact; repeat to one root
//   4. collapse plan  (optional, RTB_COLLAPSE_SAH_OPTIMAL; default is the greedy
//                     "largest child first" of step 5)
//                     bottom-up dynamic programme over the binary tree (after
//                     Ylitie, Karras & Laine 2017): for every subtree, the
//                     cheapest way (SAH) to lay it out in 1..7 child slots of
//                     a wide node — one leaf child (<= 3 triangles), one inner
//                     child (its own wide node), or split between its two
//                     children; run round by round in PLOC merge order, which
//                     is a topological order
//   5. collapse       top-down, level-synchronous: expand each wide node's
//                     children as the plan says, slot assignment for
//                     octant-ordered traversal, 8-bit quantisation, triangles
//                     rewritten in leaf order
// Bodies are RTB_HD; see rtb_wavefront.h for why.
#pragma once
#include "rtb_shade.h"

namespace rtb {

struct B2Node {  // binary node, 32 bytes
    float lox, loy, loz; int32_t left;   // leaf: left = caller's triangle index
    float hix, hiy, hiz; int32_t right;  // leaf: right = -1
};

#if defined(__CUDA_ARCH__)
RTB_HD int atomic_add_i(int32_t *p, int v) { return atomicAdd(p, v); }
RTB_HD void atomic_min_i(int32_t *p, int v) { atomicMin(p, v); }
RTB_HD void atomic_max_i(int32_t *p, int v) { atomicMax(p, v); }
RTB_HD void atomic_add_f(float *p, float v) { atomicAdd(p, v); }
#else
RTB_HD int atomic_add_i(int32_t *p, int v) { int o = *p; *p += v; return o; }
RTB_HD void atomic_min_i(int32_t *p, int v) { if (v < *p) *p = v; }
RTB_HD void atomic_max_i(int32_t *p, int v) { if (v > *p) *p = v; }
RTB_HD void atomic_add_f(float *p, float v) { *p += v; }
#endif

// order-preserving float <-> int map for atomic min/max on floats
RTB_HD int32_t float_to_ordered(float f) { int32_t b = f2i(f); return b >= 0 ? b : b ^ 0x7fffffff; }
RTB_HD float ordered_to_float(int32_t b) { return i2f(b >= 0 ? b : b ^ 0x7fffffff); }

RTB_HD float half_area(float ex, float ey, float ez) { return ffma(fadd(ex, ey), ez, fmul(ex, ey)); }
RTB_HD float b2_half_area(const B2Node &n) {
    return half_area(fsub(n.hix, n.lox), fsub(n.hiy, n.loy), fsub(n.hiz, n.loz));
}
RTB_HD float union_half_area(const B2Node &a, const B2Node &b) {
    return half_area(fsub(fmaxf(a.hix, b.hix), fminf(a.lox, b.lox)), fsub(fmaxf(a.hiy, b.hiy), fminf(a.loy, b.loy)),
                     fsub(fmaxf(a.hiz, b.hiz), fminf(a.loz, b.loz)));
}

// ------------------------------------------------------------ 1. prim setup
struct PrimSetupArgs {
    const float *vertices;        // 9 per triangle, or null when tri_in is pre-filled
    Tri48 *tri_in;                // [n] caller order
    F4 *prim_lo, *prim_hi;        // [n] bounds (w unused)
    int32_t *scene_bounds;        // 6 ordered ints: min xyz, max xyz
    int32_t *bad;                 // raised when a vertex is not finite or beyond kMaxCoord
    int32_t n;
};
// Coordinates the builder accepts: finite and |x| <= 2^100.  (An infinite vertex gives a cluster whose merged area is
// never below any other, so PLOC could not pair it; beyond 2^100 box extents and areas overflow.  The reference has no
// such check: its host SAH sweep just produces a useless tree.)
constexpr float kMaxCoord = 1.2676506e30f;
// bounds of one triangle record, the vertex check and the scene bounds; t is triangle i's record (already stored)
RTB_HD void prim_setup_bounds(const PrimSetupArgs &a, int i, const Tri48 &t);
RTB_HD void prim_setup_box(const PrimSetupArgs &a, int i, const Tri48 &t, V3 &lo, V3 &hi);
RTB_HD void prim_setup_body(const PrimSetupArgs &a, int i) {
    if (i >= a.n) return;
    Tri48 t;
    if (a.vertices) {
        const float *v = a.vertices + 9 * (size_t)i;
        t = tri_from_vertices(v3(v[0], v[1], v[2]), v3(v[3], v[4], v[5]), v3(v[6], v[7], v[8]));
        a.tri_in[i] = t;
    } else {
        t = a.tri_in[i];
    }
    prim_setup_bounds(a, i, t);
}
RTB_HD void prim_setup_bounds(const PrimSetupArgs &a, int i, const Tri48 &t) {
    V3 lo, hi;
    prim_setup_box(a, i, t, lo, hi);
#if defined(__CUDA_ARCH__)
    // one reduction per warp (redux.sync), then six atomics by one lane: 10 M triangles used to send 60 M atomics to six addresses
    const unsigned m = __activemask();
    const int lx = __reduce_min_sync(m, float_to_ordered(lo.x)), ly = __reduce_min_sync(m, float_to_ordered(lo.y)),
              lz = __reduce_min_sync(m, float_to_ordered(lo.z));
    const int hx = __reduce_max_sync(m, float_to_ordered(hi.x)), hy = __reduce_max_sync(m, float_to_ordered(hi.y)),
              hz = __reduce_max_sync(m, float_to_ordered(hi.z));
    if ((threadIdx.x & 31) == __ffs(m) - 1) {
        atomic_min_i(a.scene_bounds + 0, lx); atomic_min_i(a.scene_bounds + 1, ly); atomic_min_i(a.scene_bounds + 2, lz);
        atomic_max_i(a.scene_bounds + 3, hx); atomic_max_i(a.scene_bounds + 4, hy); atomic_max_i(a.scene_bounds + 5, hz);
    }
#else
    atomic_min_i(a.scene_bounds + 0, float_to_ordered(lo.x));
    atomic_min_i(a.scene_bounds + 1, float_to_ordered(lo.y));
    atomic_min_i(a.scene_bounds + 2, float_to_ordered(lo.z));
    atomic_max_i(a.scene_bounds + 3, float_to_ordered(hi.x));
    atomic_max_i(a.scene_bounds + 4, float_to_ordered(hi.y));
    atomic_max_i(a.scene_bounds + 5, float_to_ordered(hi.z));
#endif
}
RTB_HD void prim_setup_box(const PrimSetupArgs &a, int i, const Tri48 &t, V3 &lo, V3 &hi) {
    // Triangle::bounding_box, triangle.cuh:23-37 (p1 = p0 - e1, p2 = p0 + e2)
    V3 p0 = tri_p0(t), p1 = vsub(p0, tri_e1(t)), p2 = vadd(p0, tri_e2(t));
    lo = v3(fminf(p0.x, fminf(p1.x, p2.x)), fminf(p0.y, fminf(p1.y, p2.y)), fminf(p0.z, fminf(p1.z, p2.z)));
    hi = v3(fmaxf(p0.x, fmaxf(p1.x, p2.x)), fmaxf(p0.y, fmaxf(p1.y, p2.y)), fmaxf(p0.z, fmaxf(p1.z, p2.z)));
    // (fminf / fmaxf drop a NaN operand, so the vertices are looked at themselves)
    const float worst = fmaxf(fmaxf(fmaxf(fabsf(p0.x), fabsf(p0.y)), fmaxf(fabsf(p0.z), fabsf(p1.x))),
                              fmaxf(fmaxf(fabsf(p1.y), fabsf(p1.z)), fmaxf(fabsf(p2.x), fmaxf(fabsf(p2.y), fabsf(p2.z)))));
    const bool nan = p0.x != p0.x || p0.y != p0.y || p0.z != p0.z || p1.x != p1.x || p1.y != p1.y || p1.z != p1.z || p2.x != p2.x ||
                     p2.y != p2.y || p2.z != p2.z;
    if (nan || !(worst <= kMaxCoord)) *a.bad = 1;
    F4 l; l.x = lo.x; l.y = lo.y; l.z = lo.z; l.w = 0.f;
    F4 h; h.x = hi.x; h.y = hi.y; h.z = hi.z; h.w = 0.f;
    a.prim_lo[i] = l; a.prim_hi[i] = h;
}

// ------------------------------------------------------------ 2. morton
RTB_HD uint64_t spread21(uint64_t x) {  // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
struct MortonArgs {
    const F4 *prim_lo, *prim_hi;
    const int32_t *scene_bounds;
    uint64_t *keys; int32_t *vals;
    int32_t n;
};
RTB_HD void morton_body(const MortonArgs &a, int i) {
    if (i >= a.n) return;
    float sx = ordered_to_float(a.scene_bounds[0]), sy = ordered_to_float(a.scene_bounds[1]), sz = ordered_to_float(a.scene_bounds[2]);
    float ex = fsub(ordered_to_float(a.scene_bounds[3]), sx), ey = fsub(ordered_to_float(a.scene_bounds[4]), sy),
          ez = fsub(ordered_to_float(a.scene_bounds[5]), sz);
    float e = fmaxf(ex, fmaxf(ey, ez));
    float inv = e > 0.f ? fdiv(2097151.f, e) : 0.f;  // same scale on all axes keeps cells cubic
    const F4 lo = a.prim_lo[i], hi = a.prim_hi[i];
    float cx = fmul(0.5f, fadd(lo.x, hi.x)), cy = fmul(0.5f, fadd(lo.y, hi.y)), cz = fmul(0.5f, fadd(lo.z, hi.z));
    float qx = fminf(fmaxf(fmul(fsub(cx, sx), inv), 0.f), 2097151.f);
    float qy = fminf(fmaxf(fmul(fsub(cy, sy), inv), 0.f), 2097151.f);
    float qz = fminf(fmaxf(fmul(fsub(cz, sz), inv), 0.f), 2097151.f);
    a.keys[i] = spread21((uint64_t)qx) | (spread21((uint64_t)qy) << 1) | (spread21((uint64_t)qz) << 2);
    a.vals[i] = i;
}

// ------------------------------------------------------------ 3. PLOC
struct PlocArgs {
    B2Node *nodes;        // [2n-1]; leaves 0..n-1 in Morton order
    int32_t *count;       // [2n-1] triangles below each node
    const int32_t *cin;   // clusters (node indices), Morton order kept
    int32_t *cout;        // after merge: node index or -1
    int32_t *nn;          // nearest neighbour position
    int32_t *node_counter;  // next free inner node, starts at n
    int32_t ncl;          // current cluster count — or, with ncl_dev set, an upper bound of it known to the host
    int32_t radius;
    // device-driven rounds (round 2): the count lives in device memory, the host enqueues several rounds without
    // reading it back (every PLOC round used to cost a device->host copy of the count and a stream synchronisation)
    const int32_t *ncl_dev;  // null: ncl is the count
    int32_t *log;            // [round] = cluster count at the start of that round (null: not kept)
    int32_t round;
};
RTB_HD int ploc_ncl(const PlocArgs &a) { return a.ncl_dev ? *a.ncl_dev : a.ncl; }
RTB_HD void ploc_leaf_body(const F4 *prim_lo, const F4 *prim_hi, const int32_t *sorted, B2Node *nodes,
                           int32_t *count, int32_t *clusters, int n, int i) {
    if (i >= n) return;
    const int p = sorted[i];
    const F4 lo = prim_lo[p], hi = prim_hi[p];
    B2Node b;
    b.lox = lo.x; b.loy = lo.y; b.loz = lo.z; b.left = p;
    b.hix = hi.x; b.hiy = hi.y; b.hiz = hi.z; b.right = -1;
    nodes[i] = b;
    count[i] = 1;
    clusters[i] = i;
}
// nearest neighbour by merged surface area within +-radius positions; ties go
// to the lowest position, which guarantees at least one mutual pair per round
RTB_HD void ploc_nn_body(const PlocArgs &a, int i) {
    const int ncl = ploc_ncl(a);
    if (i == 0 && a.log) a.log[a.round] = ncl;
    if (i >= ncl) return;
    const B2Node me = a.nodes[a.cin[i]];
    float best = FLT_MAX;
    int bj = -1;
    const int j0 = i - a.radius < 0 ? 0 : i - a.radius;
    const int j1 = i + a.radius > ncl - 1 ? ncl - 1 : i + a.radius;
    for (int j = j0; j <= j1; ++j) {
        if (j == i) continue;
        const float ar = union_half_area(me, a.nodes[a.cin[j]]);
        if (bj < 0 || ar < best) { best = ar; bj = j; }  // total: every cluster names a neighbour, whatever its area
    }
    a.nn[i] = bj;
}
RTB_HD void ploc_merge_body(const PlocArgs &a, int n_leaves, int i) {
    if (i >= ploc_ncl(a)) {
        if (i < a.ncl) a.cout[i] = -1;  // between the count and the host's bound: dropped by the compaction over the bound
        return;
    }
    const int j = a.nn[i];
    const int ci = a.cin[i];
    if (j >= 0 && a.nn[j] == i) {
        if (i < j) {
            const int cj = a.cin[j];
            const int id = atomic_add_i(a.node_counter, 1);
            const B2Node x = a.nodes[ci], y = a.nodes[cj];
            B2Node m;
            m.lox = fminf(x.lox, y.lox); m.loy = fminf(x.loy, y.loy); m.loz = fminf(x.loz, y.loz);
            m.hix = fmaxf(x.hix, y.hix); m.hiy = fmaxf(x.hiy, y.hiy); m.hiz = fmaxf(x.hiz, y.hiz);
            m.left = ci; m.right = cj;
            a.nodes[id] = m;
            a.count[id] = a.count[ci] + a.count[cj];
            a.cout[i] = id;
        } else {
            a.cout[i] = -1;
        }
    } else {
        a.cout[i] = ci;
    }
    (void)n_leaves;
}

// ------------------------------------------------------------ 4. collapse plan
// cost[n][i-1], i = 1..7: SAH cost of subtree n laid out in at most i child
// slots.  plan[n][i-1]: kPlanLeaf / kPlanInner (one slot), k = 1..6 (split:
// the left child gets k slots, the right one i-k), 0 (same as with i-1
// slots).  plan[n][7]: how a wide node rooted at n splits its 8 slots.
constexpr uint8_t kPlanLeaf = 100, kPlanInner = 101;
struct PlanArgs {
    const B2Node *nodes;
    const int32_t *count;
    float *cost;      // [num binary nodes][7]
    uint8_t *plan;    // [num binary nodes][8]
    int32_t first, n; // this launch covers binary nodes first .. first+n-1 (one PLOC round: children are older)
    int32_t max_leaf;
};
RTB_HD void plan_child_costs(const PlanArgs &a, int c, float *d) {
    const B2Node x = a.nodes[c];
    if (x.right < 0) {  // a single triangle: a leaf child whatever the slot budget
        const float v = fmul(b2_half_area(x), kSahTriCost);
        for (int i = 0; i < 7; ++i) d[i] = v;
    } else {
        for (int i = 0; i < 7; ++i) d[i] = a.cost[(size_t)c * 7 + i];
    }
}
RTB_HD void plan_body(const PlanArgs &a, int tid) {
    if (tid >= a.n) return;
    const int id = a.first + tid;
    const B2Node self = a.nodes[id];
    if (self.right < 0) return;
    float dl[7], dr[7];
    plan_child_costs(a, self.left, dl);
    plan_child_costs(a, self.right, dr);
    // best split of i slots between the two children, i = 2..8
    float split[9]; uint8_t split_k[9];
    for (int i = 2; i <= 8; ++i) {
        float best = FLT_MAX; int bk = 1;
        for (int k = 1; k < i; ++k) {
            if (k > 7 || i - k > 7) continue;
            const float v = fadd(dl[k - 1], dr[i - k - 1]);
            if (v < best) { best = v; bk = k; }
        }
        split[i] = best; split_k[i] = (uint8_t)bk;
    }
    const float area = b2_half_area(self);
    const int cnt = a.count[id];
    const float as_inner = ffma(area, kSahNodeCost, split[8]);
    const float as_leaf = cnt <= a.max_leaf ? fmul(area, fmul(kSahTriCost, (float)cnt)) : FLT_MAX;
    float *cost = a.cost + (size_t)id * 7;
    uint8_t *plan = a.plan + (size_t)id * 8;
    float cur = as_leaf <= as_inner ? as_leaf : as_inner;
    cost[0] = cur; plan[0] = as_leaf <= as_inner ? kPlanLeaf : kPlanInner;
    for (int i = 2; i <= 7; ++i) {
        if (split[i] < cur) { cur = split[i]; plan[i - 1] = split_k[i]; }
        else plan[i - 1] = 0;
        cost[i - 1] = cur;
    }
    plan[7] = split_k[8];
}
// children of the wide node rooted at binary node `root`, as planned; returns their number (<= 8)
RTB_HD int plan_expand(const B2Node *nodes, const uint8_t *plan, int root, int *ch) {
    int nc = 0;
    int st_n[16], st_i[16]; int sp = 0;
    const B2Node r = nodes[root];
    const int k8 = plan[(size_t)root * 8 + 7];
    st_n[sp] = r.right; st_i[sp] = 8 - k8; ++sp;
    st_n[sp] = r.left; st_i[sp] = k8; ++sp;
    while (sp) {
        --sp;
        const int n = st_n[sp]; int i = st_i[sp];
        const B2Node x = nodes[n];
        if (x.right < 0) { ch[nc++] = n; continue; }
        if (i > 7) i = 7;
        const uint8_t *pl = plan + (size_t)n * 8;
        while (i > 1 && pl[i - 1] == 0) --i;
        const int d = pl[i - 1];
        if (i == 1 || d == kPlanLeaf || d == kPlanInner) { ch[nc++] = n; continue; }
        st_n[sp] = x.right; st_i[sp] = i - d; ++sp;
        st_n[sp] = x.left; st_i[sp] = d; ++sp;
    }
    return nc;
}

// ------------------------------------------------------------ 5. collapse
struct WorkItem {
    int32_t b2;    // binary node to expand
    int32_t wide;  // index of the 8-wide node to write
};
struct CollapseArgs {
    const B2Node *nodes;
    const int32_t *count;
    const uint8_t *plan;       // collapse plan, or null: open the largest child first (A/B)
    const Tri48 *tri_in;       // caller order
    const TriMeta *meta_in;    // caller order
    Q4 *nodes8;
    Tri48 *tris_out;           // leaf order
    TriMeta *meta_out;
    int32_t *prim_out;         // leaf order -> caller index
    int32_t *leaf_of_prim;     // caller index -> leaf order
    int32_t *node_counter;     // next free wide node (root = 0 pre-allocated)
    int32_t *tri_counter;
    const WorkItem *work_in; int32_t n_in;  // (n_in_dev set: n_in is unused)
    const int32_t *n_in_dev;                // work items of this level, in device memory: levels are enqueued without reading it back
    WorkItem *work_out; int32_t *n_out;
    float *sah;                // accumulated SAH cost (unnormalised)
    int32_t max_leaf;          // 1..3
    int32_t *gather;           // [n] leaf order: binary subtree whose triangles start at this slot, or -1 (gather_body)
};
// triangles of the leaf children, in leaf order: slot d holds the binary subtree whose (<= 3) triangles go to d, d+1, ..
// (left to right) or -1; one thread per slot, after the collapse
RTB_HD void gather_body(const CollapseArgs &a, int n, int d) {
    if (d >= n) return;
    const int root = a.gather[d];
    if (root < 0) return;
    int st[4]; int sp = 0; st[sp++] = root;
    int dst = d;
    while (sp) {
        const int ni = st[--sp];
        const B2Node x = a.nodes[ni];
        if (x.right < 0) {
            if (a.tri_in) {  // (null: the tree over the instances of a two-level scene, whose leaves are boxes)
                a.tris_out[dst] = a.tri_in[x.left];
                a.meta_out[dst] = a.meta_in[x.left];
            }
            a.prim_out[dst] = x.left;
            a.leaf_of_prim[x.left] = dst;
            dst++;
        } else {
            st[sp++] = x.right; st[sp++] = x.left;
        }
    }
}

RTB_HD void collapse_body(const CollapseArgs &a, int tid) {
    if (tid >= (a.n_in_dev ? *a.n_in_dev : a.n_in)) return;
    const WorkItem item = a.work_in[tid];
    const B2Node self = a.nodes[item.b2];
    int ch[8];
    int nc;
    // boxes and triangle counts of the children chosen so far: every level of this loop costs one round trip to
    // memory (the two nodes a child is opened into), and nothing is read twice
    B2Node cb[8];
    int ccnt[8];
    if (a.count[item.b2] <= a.max_leaf) { ch[0] = item.b2; nc = 1; }  // tiny scene: root is one leaf
    else if (a.plan) { nc = plan_expand(a.nodes, a.plan, item.b2, ch); }
    else { ch[0] = self.left; ch[1] = self.right; nc = 2; }
    for (int k = 0; k < nc; ++k) { cb[k] = a.nodes[ch[k]]; ccnt[k] = a.count[ch[k]]; }
    // without a plan: open the largest child until 8 children or only leaves remain
    while (!a.plan && nc < 8) {
        int best = -1; float best_a = -1.f;
        for (int k = 0; k < nc; ++k) {
            if (ccnt[k] > a.max_leaf) {
                const float ar = b2_half_area(cb[k]);
                if (ar > best_a) { best_a = ar; best = k; }
            }
        }
        if (best < 0) break;
        const int l = cb[best].left, r = cb[best].right;
        ch[best] = l; ch[nc] = r;
        cb[best] = a.nodes[l]; cb[nc] = a.nodes[r];
        ccnt[best] = a.count[l]; ccnt[nc] = a.count[r];
        ++nc;
    }
    // slot assignment (greedy): slot s "looks" towards (s&1?+:-, s&2?+:-, s&4?+:-);
    // a child goes to the slot best aligned with its offset from the node
    // centre, so that `slot ^ octinv` orders children front to back
    const float pcx = fmul(0.5f, fadd(self.lox, self.hix)), pcy = fmul(0.5f, fadd(self.loy, self.hiy)),
                pcz = fmul(0.5f, fadd(self.loz, self.hiz));
    // (every index into cost / todo / free below is a compile-time constant once the loops are unrolled: the matrix stays
    // in registers; dynamically indexed it lived in local memory, and the 8 x 64 dependent compares of the greedy
    // assignment were a third of a level's critical path)
    float cost[8][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const B2Node &c = cb[k < nc ? k : 0];
        const float dx = fsub(fmul(0.5f, fadd(c.lox, c.hix)), pcx);
        const float dy = fsub(fmul(0.5f, fadd(c.loy, c.hiy)), pcy);
        const float dz = fsub(fmul(0.5f, fadd(c.loz, c.hiz)), pcz);
#pragma unroll
        for (int s = 0; s < 8; ++s)
            cost[k][s] = fadd(fadd((s & 1) ? dx : -dx, (s & 2) ? dy : -dy), (s & 4) ? dz : -dz);
    }
    int slot_child[8];
    for (int s = 0; s < 8; ++s) slot_child[s] = -1;
    uint32_t todo = (1u << nc) - 1u, free_slots = 0xffu;  // children not placed yet, slots not taken yet
    for (int round = 0; round < nc; ++round) {
        float bc = -FLT_MAX; int bk = -1, bs = -1;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (!(todo >> k & 1u)) continue;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                if (!(free_slots >> s & 1u)) continue;
                if (cost[k][s] > bc) { bc = cost[k][s]; bk = k; bs = s; }
            }
        }
        slot_child[bs] = bk;
        todo &= ~(1u << bk); free_slots &= ~(1u << bs);
    }
    // allocate children and triangles
    int n_inner = 0, n_tris = 0;
    for (int k = 0; k < nc; ++k) {
        const int cnt = ccnt[k];
        if (cnt > a.max_leaf) n_inner++; else n_tris += cnt;
    }
    const int child_base = n_inner ? atomic_add_i(a.node_counter, n_inner) : 0;
    const int tri_base = n_tris ? atomic_add_i(a.tri_counter, n_tris) : 0;
    const int work_base = n_inner ? atomic_add_i(a.n_out, n_inner) : 0;
    // quantisation frame
    const uint32_t ex = quant_exponent(fsub(self.hix, self.lox)), ey = quant_exponent(fsub(self.hiy, self.loy)),
                   ez = quant_exponent(fsub(self.hiz, self.loz));
    uint32_t meta[8], qlx[8], qly[8], qlz[8], qhx[8], qhy[8], qhz[8];
    uint32_t imask = 0;
    int inner_seen = 0, tri_off = 0;
    float sah = fmul(b2_half_area(self), kSahNodeCost);
    for (int s = 0; s < 8; ++s) {
        const int k = slot_child[s];
        if (k < 0) {
            meta[s] = 0; qlx[s] = qly[s] = qlz[s] = 255u; qhx[s] = qhy[s] = qhz[s] = 0u;
            continue;
        }
        const B2Node &c = cb[k];
        qlx[s] = quant_lo(c.lox, self.lox, ex); qhx[s] = quant_hi(c.hix, self.lox, ex);
        qly[s] = quant_lo(c.loy, self.loy, ey); qhy[s] = quant_hi(c.hiy, self.loy, ey);
        qlz[s] = quant_lo(c.loz, self.loz, ez); qhz[s] = quant_hi(c.hiz, self.loz, ez);
        const int cnt = ccnt[k];
        if (cnt > a.max_leaf) {
            meta[s] = 0x20u | (24u + (uint32_t)s);
            imask |= 1u << s;
            WorkItem w; w.b2 = ch[k]; w.wide = child_base + inner_seen;
            a.work_out[work_base + inner_seen] = w;
            inner_seen++;
        } else {
            meta[s] = (((1u << cnt) - 1u) << 5) | (uint32_t)tri_off;
            sah = ffma(b2_half_area(c), fmul(kSahTriCost, (float)cnt), sah);
            // the (<= 3) triangles of this subtree go to tri_base + tri_off ...: gathered by gather_body, one thread per
            // leaf child, after the last level (walking the subtree here put 8 x 3 dependent loads on every level's
            // critical path).  Every triangle slot is written once: the subtree at a leaf child's first slot, -1 behind it
            a.gather[tri_base + tri_off] = ch[k];
            for (int t = 1; t < cnt; ++t) a.gather[tri_base + tri_off + t] = -1;
            tri_off += cnt;
        }
    }
#if defined(__CUDA_ARCH__)
    {   // one float atomic per warp (a level of a 10 M-triangle build has a million nodes, and floating-point
        // reductions to one address are not aggregated by the compiler); the cost is a statistic, its summation order is free
        const unsigned m = __activemask();  // (whatever lanes arrive here together: the loops above diverge)
        float v = 0.f;
        for (unsigned r = m; r != 0u; r &= r - 1u) v += __shfl_sync(m, sah, __ffs(r) - 1);
        if ((threadIdx.x & 31) == __ffs(m) - 1) atomic_add_f(a.sah, v);
    }
#else
    atomic_add_f(a.sah, sah);
#endif
    Q4 w0, w1, w2, w3, w4;
    w0.x = f2u(self.lox); w0.y = f2u(self.loy); w0.z = f2u(self.loz);
    w0.w = ex | (ey << 8) | (ez << 16) | (imask << 24);
    w1.x = (uint32_t)child_base; w1.y = (uint32_t)tri_base;
    w1.z = meta[0] | (meta[1] << 8) | (meta[2] << 16) | (meta[3] << 24);
    w1.w = meta[4] | (meta[5] << 8) | (meta[6] << 16) | (meta[7] << 24);
#define RTB_PACK4(q, o) ((q)[o] | ((q)[o + 1] << 8) | ((q)[o + 2] << 16) | ((q)[o + 3] << 24))
    w2.x = RTB_PACK4(qlx, 0); w2.y = RTB_PACK4(qlx, 4); w2.z = RTB_PACK4(qly, 0); w2.w = RTB_PACK4(qly, 4);
    w3.x = RTB_PACK4(qlz, 0); w3.y = RTB_PACK4(qlz, 4); w3.z = RTB_PACK4(qhx, 0); w3.w = RTB_PACK4(qhx, 4);
    w4.x = RTB_PACK4(qhy, 0); w4.y = RTB_PACK4(qhy, 4); w4.z = RTB_PACK4(qhz, 0); w4.w = RTB_PACK4(qhz, 4);
#undef RTB_PACK4
    Q4 *dst = a.nodes8 + (size_t)item.wide * kNodeWords;
    dst[0] = w0; dst[1] = w1; dst[2] = w2; dst[3] = w3; dst[4] = w4;
}

// ------------------------------------------------------------ two-level scenes
// bounds of pre-computed boxes (the instances' world boxes) instead of prim_setup
RTB_HD void box_bounds_body(const F4 *lo, const F4 *hi, int32_t *scene_bounds, int n, int i) {
    if (i >= n) return;
    const F4 l = lo[i], h = hi[i];
    atomic_min_i(scene_bounds + 0, float_to_ordered(l.x));
    atomic_min_i(scene_bounds + 1, float_to_ordered(l.y));
    atomic_min_i(scene_bounds + 2, float_to_ordered(l.z));
    atomic_max_i(scene_bounds + 3, float_to_ordered(h.x));
    atomic_max_i(scene_bounds + 4, float_to_ordered(h.y));
    atomic_max_i(scene_bounds + 5, float_to_ordered(h.z));
}
// world box of every instance from its mesh's transformed VERTICES (the box of the transformed mesh box is up to 40 %
// wider for a rotated mesh, and every box a ray crosses costs a visit of that mesh's root): one thread per
// (instance, run of kInstBoundsRun triangles)
constexpr int kInstBoundsRun = 64;
struct InstBoundsArgs {
    const Tri48 *tris;          // leaf order; mesh m = [first, first + count)
    const float *xforms;        // 12 floats per instance, object -> world
    const int32_t *first, *count;  // per instance: its mesh's range
    int32_t *bounds;            // 6 ordered ints per instance: min xyz, max xyz
    int32_t num_inst, runs;     // runs = ceil(largest mesh / kInstBoundsRun)
};
RTB_HD void inst_bounds_body(const InstBoundsArgs &a, int tid) {
    const int i = tid / a.runs, c = tid % a.runs;
    if (i >= a.num_inst) return;
    const int cnt = a.count[i], t0 = c * kInstBoundsRun;
    if (t0 >= cnt) return;
    const int t1 = t0 + kInstBoundsRun < cnt ? t0 + kInstBoundsRun : cnt;
    const float *m = a.xforms + 12 * (size_t)i;
    F4 r0, r1, r2;
    r0.x = m[0]; r0.y = m[1]; r0.z = m[2]; r0.w = m[3]; r1.x = m[4]; r1.y = m[5]; r1.z = m[6]; r1.w = m[7];
    r2.x = m[8]; r2.y = m[9]; r2.z = m[10]; r2.w = m[11];
    V3 lo = v3(FLT_MAX), hi = v3(-FLT_MAX);
    for (int t = t0; t < t1; ++t) {
        const Tri48 tr = a.tris[a.first[i] + t];
        const V3 p0 = tri_p0(tr);
        const V3 q[3] = {xform_point(r0, r1, r2, p0), xform_point(r0, r1, r2, vsub(p0, tri_e1(tr))), xform_point(r0, r1, r2, vadd(p0, tri_e2(tr)))};
        for (int k = 0; k < 3; ++k) {
            lo = v3(fminf(lo.x, q[k].x), fminf(lo.y, q[k].y), fminf(lo.z, q[k].z));
            hi = v3(fmaxf(hi.x, q[k].x), fmaxf(hi.y, q[k].y), fmaxf(hi.z, q[k].z));
        }
    }
    int32_t *b = a.bounds + 6 * (size_t)i;
    atomic_min_i(b + 0, float_to_ordered(lo.x)); atomic_min_i(b + 1, float_to_ordered(lo.y)); atomic_min_i(b + 2, float_to_ordered(lo.z));
    atomic_max_i(b + 3, float_to_ordered(hi.x)); atomic_max_i(b + 4, float_to_ordered(hi.y)); atomic_max_i(b + 5, float_to_ordered(hi.z));
}
// a mesh's tree is built with indices relative to itself; in the scene's arrays its nodes start at node_off and its
// triangles at tri_off
RTB_HD void rebase_node_body(const Q4 *src, Q4 *dst, uint32_t node_off, uint32_t tri_off, int n, int i) {
    if (i >= n) return;
    const Q4 *s = src + (size_t)i * kNodeWords;
    Q4 *d = dst + (size_t)i * kNodeWords;
    Q4 w1 = s[1];
    w1.x += node_off; w1.y += tri_off;
    d[0] = s[0]; d[1] = w1; d[2] = s[2]; d[3] = s[3]; d[4] = s[4];
}

// area lights refer to their triangle by leaf-order index after the build
RTB_HD void light_fix_body(LightDev *lights, const int64_t *light_tri, const int32_t *leaf_of_prim, int n, int i) {
    if (i >= n) return;
    if (lights[i].type == RTB_AREA_LIGHT) lights[i].tri = leaf_of_prim[light_tri[i]];
    else lights[i].tri = -1;
}

}  // namespace rtb
