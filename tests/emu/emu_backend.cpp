// emu_backend.cpp — TEST INFRASTRUCTURE, not product code.
//
// Compiles the kernel bodies of rtcuda_b200/csrc (rtb_*.h) for the host and
// runs them in plain sequential loops behind the same C ABI (include/rtb.h),
// so that `pytest -m "not gpu"` can check the builder, the traversal and the
// wavefront state machine against the oracle on a machine without a GPU.
// It is built into tests/emu/librtb_emu.so, loaded only by tests/, never by
// the rtcuda_b200 package: the shipped library (librtb.so) has no CPU path
// and fails with RTB_ERR_NO_DEVICE when no sm_100 GPU is present.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <numeric>
#include <vector>

#include "rtb_engine.h"

namespace rtb {

struct HostBackend {
    int ordinal_;
    explicit HostBackend(int ordinal) : ordinal_(ordinal) {}
    ~HostBackend() { wave_cache_release(*this, wave_cache_); }
    WaveCache wave_cache_;
    WaveCache &wave_cache() { return wave_cache_; }
    int device() const { return -1; }  // (no GPU: device masks are not checked against it)
    void make_current() {}
    void sync() {}
    // the schedule options of the CUDA backend are accepted and remembered (they select kernels the emulation does not
    // have; "pool" is honoured)
    std::map<std::string, long long> opts_;
    bool set_option(const std::string &name, long long v) {
        static const char *const names[] = {"refill", "chunk", "prefetch", "tri_step", "pooled", "fused", "smem_stack", "pipelines", "pool",
                                            "ploc_tail", "trace_blocks", "nn_tiled"};
        for (const char *n : names) if (name == n) { if (name == "pool" && v != 0 && v < 1024) return false; opts_[name] = v; return true; }
        return false;
    }
    bool get_option(const std::string &name, long long &v) const {
        auto it = opts_.find(name);
        if (it == opts_.end()) return false;
        v = it->second;
        return true;
    }
    long long default_pool(size_t) const { auto it = opts_.find("pool"); return it == opts_.end() || it->second == 0 ? 1 << 16 : it->second; }

    template <class T> T *alloc(size_t n) {
        void *p = nullptr;
        if (posix_memalign(&p, 64, sizeof(T) * (n ? n : 1)) != 0) throw Error(RTB_ERR_OOM, "emu: out of memory");
        return (T *)p;
    }
    void free(void *p) { ::free(p); }
    template <class T> void upload(T *dst, const T *src, size_t n) { memcpy(dst, src, sizeof(T) * n); }
    template <class T> void download(T *dst, const T *src, size_t n) { memcpy(dst, src, sizeof(T) * n); }
    template <class T> void copy(T *dst, const T *src, size_t n) { memcpy(dst, src, sizeof(T) * n); }
    template <class T> void zero(T *p, size_t n) { memset(p, 0, sizeof(T) * n); }
    template <class T> void copy_from(HostBackend &, T *dst, const T *src, size_t n) { memcpy(dst, src, sizeof(T) * n); }
    // the "GPUs" of an emulated rtb_multi are HostBackend objects in one address space: a reduction is a loop
    struct Group { int n; };
    static Group *group_create(const std::vector<HostBackend *> &members) { return new Group{(int)members.size()}; }
    static void group_destroy(Group *g) { delete g; }
    template <class T>
    static void group_reduce(Group *, const std::vector<HostBackend *> &members, const std::vector<T *> &bufs, size_t n, int root) {
        for (size_t r = 0; r < members.size(); ++r) {  // rank order: the order NCCL's ring would not promise, but integers do not care
            if ((int)r == root) continue;
            for (size_t i = 0; i < n; ++i) bufs[root][i] += bufs[r][i];
        }
    }
    struct Comm { int rank, world; };
    static void comm_unique_id(uint8_t *out) { memset(out, 0, RTB_COMM_ID_BYTES); }
    Comm *comm_create(const uint8_t *, int rank, int world) {
        if (world != 1) throw Error(RTB_ERR_NO_DEVICE, "emu: a communicator over processes needs NCCL and GPUs");
        return new Comm{rank, world};
    }
    static void comm_destroy(Comm *c) { delete c; }
    template <class T> void comm_allreduce(Comm *, T *, size_t) {}

    template <class F> void launch(int n, F f) { for (int i = 0; i < n; ++i) f(i); }
    template <class F> void launch_trace(int n, F f) { launch(n, f); }
    void shade(const ShadeK &k) {
        const int n = k.W.c->n_mat[k.type];
        ShadeTally tally; tally.extend = 0; tally.shadow = 0;
        for (int i = 0; i < n; ++i) {
            if (k.type == 0) shade_body<0>(k.W, k.S, k.rc, k.shadows, i, tally);
            else if (k.type == 1) shade_body<1>(k.W, k.S, k.rc, k.shadows, i, tally);
            else if (k.type == 2) shade_body<2>(k.W, k.S, k.rc, k.shadows, i, tally);
            else shade_body<3>(k.W, k.S, k.rc, k.shadows, i, tally);
        }
        tally_flush(k.W.c, tally);
    }
    void generate(const GenerateK &k) {
        const int n = generate_count(k.W);
        for (int i = 0; i < n; ++i) generate_body(k.W, k.rc, i);
    }
    void control(const WaveState &W, bool shadows) { control_body(W, shadows); }
    void extend(const WaveState &W, const SceneView &S, int mode) {
        const int n = W.c->n_extend;
        for (int i = 0; i < n; ++i) { if (mode == 2) extend_body<true>(W, S, i); else extend_body<false>(W, S, i); }
    }
    void shadow(const WaveState &W, const SceneView &S, int mode) {
        const int n = W.c->n_shadow;
        for (int i = 0; i < n; ++i) { if (mode == 2) shadow_body<true>(W, S, i); else shadow_body<false>(W, S, i); }
    }
    bool trace_fused(const WaveState &, const SceneView &, int) { return false; }
    int32_t done_word_ = 0;
    int pipelines(const SceneView &) const { return 1; }
    void begin_render(int) {}
    void use_stream(int) {}
    void fork(int) {}
    void join(int) {}
    int32_t *done_flag_device(int) { return &done_word_; }
    void reset_done(int) { done_word_ = 0; }
    bool done(int) const { return done_word_ != 0; }
    void sort_pairs(uint64_t *keys, int32_t *vals, int n) {
        std::vector<int> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return keys[a] < keys[b]; });
        std::vector<uint64_t> k(n);
        std::vector<int32_t> v(n);
        for (int i = 0; i < n; ++i) { k[i] = keys[idx[i]]; v[i] = vals[idx[i]]; }
        memcpy(keys, k.data(), sizeof(uint64_t) * n);
        memcpy(vals, v.data(), sizeof(int32_t) * n);
    }
    void prim_setup(const PrimSetupArgs &a) { PrimSetupK k; k.a = a; launch(a.n, k); }
    void ploc_nn(const PlocArgs &a) { PlocNnK k; k.a = a; launch(a.ncl, k); }
    bool ploc_tail(const PlocArgs &, int, int32_t *) { return false; }  // (a launch-latency measure of the CUDA backend)
    int32_t *ploc_tail_counts() { return nullptr; }
    void compact_nonneg(const int32_t *in, int32_t *out, int n, int32_t *count) {
        int m = 0;
        for (int i = 0; i < n; ++i) if (in[i] >= 0) out[m++] = in[i];
        *count = m;
    }
    void collapse_level(const CollapseK &k, int, int32_t *zero, int32_t *levels) {
        const int n = *k.a.n_in_dev;
        *zero = 0;
        if (n > 0) *levels += 1;
        for (int i = 0; i < n; ++i) collapse_body(k.a, i);
    }
    using Time = std::chrono::steady_clock::time_point;
    Time now() { return std::chrono::steady_clock::now(); }
    float elapsed_ms(Time a, Time b) { return std::chrono::duration<float, std::milli>(b - a).count(); }
    float elapsed_keep(Time a, Time b) { return elapsed_ms(a, b); }
    void release(Time) {}
    void wait(Time) {}
};

}  // namespace rtb

#define RTB_BACKEND rtb::HostBackend
#include "rtb_api_impl.h"
