// ref_harness.cu — TEST INFRASTRUCTURE: drives the UNMODIFIED reference
// (lashhw/rtcuda) on a B200.  The reference's headers are #included from
// where they lie (-I/root/reference at build time, see oracle/Makefile);
// nothing of the reference is copied into this repository.  The binary
// (oracle/_ref/ref_harness, git-ignored) travels to the GPU box.
//
// main.cu hard-codes its scene, resolution and sample count (main.cu:60,159-170),
// so this harness builds the reference's own structs (Triangle, Material,
// Light, Primitive, Bvh, Scene, Camera) from a scene file written by
// rtb_scene_desc_save() and then calls the reference's own code:
//   trace   : Bvh::traverse closest-hit (bvh.cuh:251) on a ray file -> hits
//   any     : Bvh::traverse any-hit (bvh.cuh:306) on rays + excluded triangle
//   render  : render() (render.cuh:366) unmodified, timed as a whole call
//   loop    : the reference's stage kernels (init/mat/gen/ah/ch + compact,
//             render.cuh:84-364) driven by a restatement of its host loop
//             (render.cuh:428-449) so that rays can be counted, the loop timed
//             on its own, the seed chosen (init_rand_states takes it as an
//             argument, render.cuh:68) and several passes accumulated past the
//             int limit of num_pixels*num_samples (render.cuh:371)
// Output: one line starting with "JSON " per command, binary results to files.
#include <iostream>
#include <fstream>
#include <cmath>
#include <vector>
#include <numeric>
#include <memory>
#include <cfloat>
#include <array>
#include <cassert>
#include <chrono>
#include <algorithm>
#include <unordered_map>
#include <stack>
#include <string>
#include <cstring>
#include <cstdio>
#include <climits>

#include <curand_kernel.h>
#include <cub/cub.cuh>

// same order as main.cu:18-37 (happly.h is host I/O and not needed)
#include "constant.hpp"
#include "profiler.hpp"
#include "vec3.cuh"
#include "matrix4x4.hpp"
#include "transform.hpp"
#include "utility.cuh"
#include "ray.cuh"
#include "bounding_box.cuh"
#include "aabb_intersector.cuh"
#include "intersection.hpp"
#include "material.cuh"
#include "triangle.cuh"
#include "device_stack.cuh"
#include "light.cuh"
#include "primitive.cuh"
#include "bvh.cuh"
#include "scene.cuh"
#include "camera.cuh"
#include "render.cuh"

struct HitOut { float t, u, v; int prim; };
struct RayIn { float o[3], d[3], tmax; };  // == Ray, ray.cuh:16-18 (28 bytes)
struct MatIn { float albedo[3]; float ior; int type; };  // == Material (20 bytes)
struct LightIn { int type; float pos[3]; long long triangle; float L[3]; int pad; };

static_assert(sizeof(Ray) == 28, "Ray layout");
static_assert(sizeof(Material) == 20, "Material layout");
static_assert(sizeof(Triangle) == 48, "Triangle layout");
static_assert(sizeof(Primitive) == 24, "Primitive layout");
static_assert(sizeof(Light) == 40, "Light layout");
static_assert(sizeof(Camera) == 48, "Camera layout");

__global__ void h_trace_closest(Bvh bvh, const Triangle *d_tris, const Ray *rays, int n, HitOut *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DeviceStack stack;
    Ray ray = rays[i];
    Intersection isect;
    Primitive *prim;
    bool hit = bvh.traverse(stack, ray, isect, prim);
    HitOut o;
    if (hit) { o.t = isect.t; o.u = isect.u; o.v = isect.v; o.prim = (int)(prim->d_triangle - d_tris); }
    else { o.t = 0.f; o.u = 0.f; o.v = 0.f; o.prim = -1; }
    out[i] = o;
}

__global__ void h_trace_any(Bvh bvh, const Triangle *d_tris, const Ray *rays, const int *excluded, int n, unsigned char *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DeviceStack stack;
    Ray ray = rays[i];
    const Triangle *ex = excluded[i] >= 0 ? d_tris + excluded[i] : nullptr;
    out[i] = bvh.traverse(ex, stack, ray) ? 1 : 0;
}

struct RefScene {
    Triangle *d_triangles = nullptr;
    Material *d_materials = nullptr;
    Light *d_lights = nullptr;
    int num_triangles = 0, num_lights = 0;
    Scene scene;
    double bvh_build_ms = 0;
};

static bool load_scene(const char *path, RefScene &rs) {
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); return false; }
    unsigned magic; long long n; int nm, nl;
    if (fread(&magic, 4, 1, f) != 1 || magic != 0x53425452u || fread(&n, 8, 1, f) != 1 || fread(&nm, 4, 1, f) != 1 || fread(&nl, 4, 1, f) != 1) return false;
    std::vector<float> verts(9 * (size_t)n);
    std::vector<int> mat((size_t)n), light((size_t)n);
    std::vector<MatIn> mats((size_t)nm);
    std::vector<LightIn> lights((size_t)nl);
    bool ok = fread(verts.data(), 4, verts.size(), f) == verts.size() && fread(mat.data(), 4, (size_t)n, f) == (size_t)n &&
              fread(light.data(), 4, (size_t)n, f) == (size_t)n && fread(mats.data(), 20, (size_t)nm, f) == (size_t)nm &&
              fread(lights.data(), 40, (size_t)nl, f) == (size_t)nl;
    fclose(f);
    if (!ok) return false;
    // materials, as main.cu:41-56 does
    std::vector<Material> materials;
    for (auto &m : mats) {
        if (m.type == MATTE) materials.push_back(Material::make_matte(Vec3(m.albedo[0], m.albedo[1], m.albedo[2])));
        else if (m.type == MIRROR) materials.push_back(Material::make_mirror(Vec3(m.albedo[0], m.albedo[1], m.albedo[2])));
        else materials.push_back(Material::make_glass(m.ior));
    }
    CHECK_CUDA(cudaMalloc(&rs.d_materials, materials.size() * sizeof(Material)));
    CHECK_CUDA(cudaMemcpy(rs.d_materials, materials.data(), materials.size() * sizeof(Material), cudaMemcpyHostToDevice));
    // triangles through the reference's host constructor, main.cu:75-85
    std::vector<Triangle> triangles;
    triangles.reserve((size_t)n);
    for (long long i = 0; i < n; ++i) {
        const float *v = &verts[9 * (size_t)i];
        triangles.emplace_back(Vec3(v[0], v[1], v[2]), Vec3(v[3], v[4], v[5]), Vec3(v[6], v[7], v[8]));
    }
    rs.num_triangles = (int)n;
    CHECK_CUDA(cudaMalloc(&rs.d_triangles, (size_t)n * sizeof(Triangle)));
    CHECK_CUDA(cudaMemcpy(rs.d_triangles, triangles.data(), (size_t)n * sizeof(Triangle), cudaMemcpyHostToDevice));
    // lights, main.cu:125-138 (index order given by the scene file)
    std::vector<Light> ls;
    for (auto &l : lights) {
        if (l.type == AREA_LIGHT) ls.push_back(Light::make_area_light(&rs.d_triangles[l.triangle], Vec3(l.L[0], l.L[1], l.L[2])));
        else ls.push_back(Light::make_point_light(Vec3(l.pos[0], l.pos[1], l.pos[2]), Vec3(l.L[0], l.L[1], l.L[2])));
    }
    rs.num_lights = (int)ls.size();
    if (rs.num_lights) {
        CHECK_CUDA(cudaMalloc(&rs.d_lights, ls.size() * sizeof(Light)));
        CHECK_CUDA(cudaMemcpy(rs.d_lights, ls.data(), ls.size() * sizeof(Light), cudaMemcpyHostToDevice));
    }
    // primitives, main.cu:141-148
    std::vector<Primitive> primitives;
    primitives.reserve((size_t)n);
    for (long long i = 0; i < n; ++i) {
        if (light[(size_t)i] >= 0) primitives.emplace_back(&rs.d_triangles[i], &rs.d_materials[mat[(size_t)i]], &rs.d_lights[light[(size_t)i]]);
        else primitives.emplace_back(&rs.d_triangles[i], &rs.d_materials[mat[(size_t)i]]);
    }
    auto t0 = std::chrono::steady_clock::now();
    Bvh bvh(triangles, primitives);  // main.cu:151
    rs.bvh_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    rs.scene = {bvh, rs.num_lights, rs.d_lights};  // main.cu:156
    return true;
}

static Camera make_camera(int w, int h) {  // main.cu:162-166
    return Camera(Vec3(0.5f, 0.5f, 1.5f), Vec3(0.5f, 0.5f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), 37.8f, (float)w / (float)h);
}

template <class T>
static std::vector<T> read_file(const char *path) {
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    size_t bytes = (size_t)in.tellg();
    in.seekg(0);
    std::vector<T> v(bytes / sizeof(T));
    in.read((char *)v.data(), (std::streamsize)(v.size() * sizeof(T)));
    return v;
}
static void write_file(const char *path, const void *p, size_t bytes) {
    FILE *f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
    fwrite(p, 1, bytes, f);
    fclose(f);
}

// ---- the reference's device state, set up exactly like render.cuh:373-410 ----
struct LoopState {
    Vec3 *fb;
    int *mat_p, *gen_p, *ah_p, *ch_p;
    bool *mat_v, *gen_v, *ah_v, *ch_v;
    int *mat_c, *gen_c, *ah_c, *ch_c;
    int *n_mat, *n_gen, *n_ah, *n_ch;
    int num_pixels;
};
static LoopState setup_loop(int width, int height, int num_samples, int max_bounces, Camera camera, Scene scene) {
    LoopState s;
    s.num_pixels = width * height;
    int camera_ray_end_id = s.num_pixels * num_samples;
    cuda_malloc_symbol(d_rand_states, NUM_WORKING_PATHS * sizeof(curandState));
    s.fb = cuda_malloc_symbol(d_framebuffer, s.num_pixels * sizeof(Vec3));
    s.mat_p = cuda_malloc_symbol(d_mat_pending, NUM_WORKING_PATHS * sizeof(int));
    s.gen_p = cuda_malloc_symbol(d_gen_pending, NUM_WORKING_PATHS * sizeof(int));
    s.ah_p = cuda_malloc_symbol(d_ah_pending, NUM_WORKING_PATHS * sizeof(int));
    s.ch_p = cuda_malloc_symbol(d_ch_pending, 3 * NUM_WORKING_PATHS * sizeof(int));
    s.mat_v = cuda_malloc_symbol(d_mat_pending_valid, NUM_WORKING_PATHS * sizeof(bool));
    s.gen_v = cuda_malloc_symbol(d_gen_pending_valid, NUM_WORKING_PATHS * sizeof(bool));
    s.ah_v = cuda_malloc_symbol(d_ah_pending_valid, NUM_WORKING_PATHS * sizeof(bool));
    s.ch_v = cuda_malloc_symbol(d_ch_pending_valid, 3 * NUM_WORKING_PATHS * sizeof(bool));
    s.mat_c = cuda_malloc_symbol(d_mat_pending_compact, NUM_WORKING_PATHS * sizeof(int));
    s.gen_c = cuda_malloc_symbol(d_gen_pending_compact, NUM_WORKING_PATHS * sizeof(int));
    s.ah_c = cuda_malloc_symbol(d_ah_pending_compact, NUM_WORKING_PATHS * sizeof(int));
    s.ch_c = cuda_malloc_symbol(d_ch_pending_compact, 3 * NUM_WORKING_PATHS * sizeof(int));
    cuda_malloc_symbol(d_ray_pool, sizeof(RayPool));
    cuda_malloc_symbol(d_path_ray_payload, sizeof(PathRayPayload));
    cuda_malloc_symbol(d_ah_shadow_ray_payload, sizeof(ShadowRayPayload));
    cuda_malloc_symbol(d_ch_shadow_ray_payload, sizeof(ShadowRayPayload));
    CHECK_CUDA(cudaGetSymbolAddress((void **)&s.n_mat, d_num_mat_pending));
    CHECK_CUDA(cudaGetSymbolAddress((void **)&s.n_gen, d_num_gen_pending));
    CHECK_CUDA(cudaGetSymbolAddress((void **)&s.n_ah, d_num_ah_pending));
    CHECK_CUDA(cudaGetSymbolAddress((void **)&s.n_ch, d_num_ch_pending));
    CHECK_CUDA(cudaMemcpyToSymbol(d_width, &width, sizeof(int)));
    CHECK_CUDA(cudaMemcpyToSymbol(d_height, &height, sizeof(int)));
    CHECK_CUDA(cudaMemcpyToSymbol(d_num_samples, &num_samples, sizeof(int)));
    CHECK_CUDA(cudaMemcpyToSymbol(d_max_bounces, &max_bounces, sizeof(int)));
    CHECK_CUDA(cudaMemcpyToSymbol(d_camera_ray_end_id, &camera_ray_end_id, sizeof(int)));
    CHECK_CUDA(cudaMemcpyToSymbol(d_scene, &scene, sizeof(Scene)));
    CHECK_CUDA(cudaMemcpyToSymbol(d_camera, &camera, sizeof(Camera)));
    return s;
}

struct LoopCounts { unsigned long long ch = 0, ah = 0, mat = 0, gen = 0, iters = 0; float ms = 0; };

// one pass of the reference's host loop, render.cuh:416-449, with its kernels
static LoopCounts run_pass(const LoopState &s, int num_samples, int seed) {
    constexpr int B = 64;  // BLOCK_SIZE, render.cuh:413
    LoopCounts c;
    int camera_ray_start_id = 0, camera_ray_end_id = s.num_pixels * num_samples;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    init_rand_states<<<(NUM_WORKING_PATHS + B - 1) / B, B>>>(seed);
    init_path_ray_payload<<<(NUM_WORKING_PATHS + B - 1) / B, B>>>();
    CHECK_CUDA(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    int n_mat, n_gen, n_ah, n_ch;
    while (true) {
        init<<<(NUM_WORKING_PATHS + B - 1) / B, B>>>();
        compact(NUM_WORKING_PATHS, s.mat_p, s.mat_v, s.mat_c, s.n_mat);
        compact(NUM_WORKING_PATHS, s.gen_p, s.gen_v, s.gen_c, s.n_gen);
        CHECK_CUDA(cudaMemcpy(&n_mat, s.n_mat, sizeof(int), cudaMemcpyDeviceToHost));
        CHECK_CUDA(cudaMemcpy(&n_gen, s.n_gen, sizeof(int), cudaMemcpyDeviceToHost));
        if (n_mat == 0 && camera_ray_start_id >= camera_ray_end_id) break;
        if (n_mat > 0) mat<<<(n_mat + B - 1) / B, B>>>();
        if (n_gen > 0) gen<<<(n_gen + B - 1) / B, B>>>(camera_ray_start_id);
        camera_ray_start_id += n_gen;
        compact(NUM_WORKING_PATHS, s.ah_p, s.ah_v, s.ah_c, s.n_ah);
        compact(3 * NUM_WORKING_PATHS, s.ch_p, s.ch_v, s.ch_c, s.n_ch);
        CHECK_CUDA(cudaMemcpy(&n_ah, s.n_ah, sizeof(int), cudaMemcpyDeviceToHost));
        CHECK_CUDA(cudaMemcpy(&n_ch, s.n_ch, sizeof(int), cudaMemcpyDeviceToHost));
        if (n_ah > 0) ah<<<(n_ah + B - 1) / B, B>>>();
        if (n_ch > 0) ch<<<(n_ch + B - 1) / B, B>>>();
        c.ch += n_ch; c.ah += n_ah; c.mat += n_mat; c.gen += n_gen; c.iters++;
    }
    cudaEventRecord(e1);
    CHECK_CUDA(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&c.ms, e0, e1);
    CHECK_CUDA(cudaGetLastError());
    return c;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr,
                "usage: ref_harness <scene.rtbs> <cmd> ...\n"
                "  trace  <rays.bin> <hits.bin>\n"
                "  any    <rays.bin> <excluded.bin> <occluded.bin>\n"
                "  traceany <rays.bin> <hits.bin> <shadow_rays.bin> <excluded.bin> <occluded.bin>\n"
                "  render <W> <H> <spp> <bounces> <steps> <warmup> [out.f32]      (unmodified render())\n"
                "  loop   <W> <H> <spp_per_pass> <bounces> <passes> <seed0> [sum.f32]\n");
        return 2;
    }
    int dev_count = 0;
    if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) { fprintf(stderr, "no CUDA device\n"); return 3; }
    RefScene rs;
    if (!load_scene(argv[1], rs)) { fprintf(stderr, "bad scene file\n"); return 2; }
    std::string cmd = argv[2];
    printf("JSON {\"cmd\":\"scene\",\"triangles\":%d,\"nodes\":%d,\"max_depth\":%d,\"bvh_build_ms\":%.3f}\n", rs.num_triangles,
           rs.scene.bvh.num_nodes, rs.scene.bvh.max_depth, rs.bvh_build_ms);
    if (cmd == "trace" && argc >= 5) {
        std::vector<RayIn> rays = read_file<RayIn>(argv[3]);
        int n = (int)rays.size();
        Ray *d_rays; HitOut *d_hits;
        CHECK_CUDA(cudaMalloc(&d_rays, (size_t)n * sizeof(Ray)));
        CHECK_CUDA(cudaMalloc(&d_hits, (size_t)n * sizeof(HitOut)));
        CHECK_CUDA(cudaMemcpy(d_rays, rays.data(), (size_t)n * sizeof(Ray), cudaMemcpyHostToDevice));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            h_trace_closest<<<(n + 63) / 64, 64>>>(rs.scene.bvh, rs.d_triangles, d_rays, n, d_hits);
            cudaEventRecord(e1);
            CHECK_CUDA(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            best = std::min(best, ms);
        }
        CHECK_CUDA(cudaGetLastError());
        std::vector<HitOut> hits((size_t)n);
        CHECK_CUDA(cudaMemcpy(hits.data(), d_hits, (size_t)n * sizeof(HitOut), cudaMemcpyDeviceToHost));
        write_file(argv[4], hits.data(), hits.size() * sizeof(HitOut));
        printf("JSON {\"cmd\":\"trace\",\"rays\":%d,\"ms\":%.4f,\"mrays_s\":%.2f}\n", n, best, n / best * 1e-3);
    } else if (cmd == "any" && argc >= 6) {
        std::vector<RayIn> rays = read_file<RayIn>(argv[3]);
        std::vector<int> ex = read_file<int>(argv[4]);
        int n = (int)rays.size();
        Ray *d_rays; int *d_ex; unsigned char *d_out;
        CHECK_CUDA(cudaMalloc(&d_rays, (size_t)n * sizeof(Ray)));
        CHECK_CUDA(cudaMalloc(&d_ex, (size_t)n * sizeof(int)));
        CHECK_CUDA(cudaMalloc(&d_out, (size_t)n));
        CHECK_CUDA(cudaMemcpy(d_rays, rays.data(), (size_t)n * sizeof(Ray), cudaMemcpyHostToDevice));
        CHECK_CUDA(cudaMemcpy(d_ex, ex.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
        h_trace_any<<<(n + 63) / 64, 64>>>(rs.scene.bvh, rs.d_triangles, d_rays, d_ex, n, d_out);
        CHECK_CUDA(cudaDeviceSynchronize());
        std::vector<unsigned char> out((size_t)n);
        CHECK_CUDA(cudaMemcpy(out.data(), d_out, (size_t)n, cudaMemcpyDeviceToHost));
        write_file(argv[5], out.data(), out.size());
        printf("JSON {\"cmd\":\"any\",\"rays\":%d}\n", n);
    } else if (cmd == "traceany" && argc >= 8) {  // both queries on one scene load (the host SAH build of a 10 M-triangle scene takes 35 s)
        std::vector<RayIn> rays = read_file<RayIn>(argv[3]);
        std::vector<RayIn> srays = read_file<RayIn>(argv[5]);
        std::vector<int> ex = read_file<int>(argv[6]);
        int n = (int)rays.size(), ns = (int)srays.size();
        Ray *d_rays, *d_srays; HitOut *d_hits; int *d_ex; unsigned char *d_out;
        CHECK_CUDA(cudaMalloc(&d_rays, (size_t)n * sizeof(Ray)));
        CHECK_CUDA(cudaMalloc(&d_hits, (size_t)n * sizeof(HitOut)));
        CHECK_CUDA(cudaMalloc(&d_srays, (size_t)ns * sizeof(Ray)));
        CHECK_CUDA(cudaMalloc(&d_ex, (size_t)ns * sizeof(int)));
        CHECK_CUDA(cudaMalloc(&d_out, (size_t)ns));
        CHECK_CUDA(cudaMemcpy(d_rays, rays.data(), (size_t)n * sizeof(Ray), cudaMemcpyHostToDevice));
        CHECK_CUDA(cudaMemcpy(d_srays, srays.data(), (size_t)ns * sizeof(Ray), cudaMemcpyHostToDevice));
        CHECK_CUDA(cudaMemcpy(d_ex, ex.data(), (size_t)ns * sizeof(int), cudaMemcpyHostToDevice));
        h_trace_closest<<<(n + 63) / 64, 64>>>(rs.scene.bvh, rs.d_triangles, d_rays, n, d_hits);
        h_trace_any<<<(ns + 63) / 64, 64>>>(rs.scene.bvh, rs.d_triangles, d_srays, d_ex, ns, d_out);
        CHECK_CUDA(cudaDeviceSynchronize());
        CHECK_CUDA(cudaGetLastError());
        std::vector<HitOut> hits((size_t)n);
        std::vector<unsigned char> out((size_t)ns);
        CHECK_CUDA(cudaMemcpy(hits.data(), d_hits, (size_t)n * sizeof(HitOut), cudaMemcpyDeviceToHost));
        CHECK_CUDA(cudaMemcpy(out.data(), d_out, (size_t)ns, cudaMemcpyDeviceToHost));
        write_file(argv[4], hits.data(), hits.size() * sizeof(HitOut));
        write_file(argv[7], out.data(), out.size());
        printf("JSON {\"cmd\":\"traceany\",\"rays\":%d,\"shadow_rays\":%d}\n", n, ns);
    } else if (cmd == "render" && argc >= 9) {
        int W = atoi(argv[3]), H = atoi(argv[4]), spp = atoi(argv[5]), bounces = atoi(argv[6]), steps = atoi(argv[7]), warm = atoi(argv[8]);
        Camera cam = make_camera(W, H);
        std::vector<Vec3> fb;
        std::vector<double> ms_list;
        for (int it = 0; it < warm + steps; ++it) {
            CHECK_CUDA(cudaDeviceSynchronize());
            auto t0 = std::chrono::steady_clock::now();
            render(W, H, spp, bounces, cam, rs.scene, fb);  // render.cuh:366, unmodified
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (it >= warm) ms_list.push_back(ms);
        }
        double sum = 0; for (double m : ms_list) sum += m;
        if (argc >= 10) write_file(argv[9], fb.data(), fb.size() * sizeof(Vec3));
        printf("JSON {\"cmd\":\"render\",\"width\":%d,\"height\":%d,\"spp\":%d,\"bounces\":%d,\"steps\":%d,\"warmup\":%d,\"ms_per_call\":%.3f}\n", W, H, spp,
               bounces, steps, warm, steps ? sum / steps : 0.0);
    } else if (cmd == "loop" && argc >= 9) {
        int W = atoi(argv[3]), H = atoi(argv[4]), spp = atoi(argv[5]), bounces = atoi(argv[6]), passes = atoi(argv[7]), seed0 = atoi(argv[8]);
        if ((long long)W * H * spp > INT_MAX) { fprintf(stderr, "num_pixels*num_samples overflows int (render.cuh:371)\n"); return 2; }
        Camera cam = make_camera(W, H);
        LoopState s = setup_loop(W, H, spp, bounces, cam, rs.scene);
        init_framebuffer<<<(s.num_pixels + 63) / 64, 64>>>(s.num_pixels);
        LoopCounts total;
        float ms_min = 1e30f;
        std::string pass_ms = "[", pass_rays = "[";
        for (int p = 0; p < passes; ++p) {
            LoopCounts c = run_pass(s, spp, seed0 + p);
            pass_ms += (p ? "," : "") + std::to_string(c.ms);
            pass_rays += (p ? "," : "") + std::to_string(c.ch + c.ah);
            total.ch += c.ch; total.ah += c.ah; total.mat += c.mat; total.gen += c.gen; total.iters += c.iters; total.ms += c.ms;
            ms_min = std::min(ms_min, c.ms);
        }
        if (argc >= 10) {  // raw radiance sums (not divided, not gamma-encoded)
            std::vector<Vec3> fb((size_t)s.num_pixels);
            CHECK_CUDA(cudaMemcpy(fb.data(), s.fb, fb.size() * sizeof(Vec3), cudaMemcpyDeviceToHost));
            write_file(argv[9], fb.data(), fb.size() * sizeof(Vec3));
        }
        double rays = (double)(total.ch + total.ah);
        printf("JSON {\"cmd\":\"loop\",\"width\":%d,\"height\":%d,\"spp_per_pass\":%d,\"bounces\":%d,\"passes\":%d,\"seed0\":%d,"
               "\"ch_rays\":%llu,\"ah_rays\":%llu,\"mat\":%llu,\"gen\":%llu,\"iterations\":%llu,\"ms_loop_total\":%.3f,\"ms_loop_min\":%.3f,"
               "\"mrays_s\":%.2f,\"pass_ms\":%s],\"pass_rays\":%s]}\n",
               W, H, spp, bounces, passes, seed0, total.ch, total.ah, total.mat, total.gen, total.iters, total.ms, ms_min,
               rays / total.ms * 1e-3, pass_ms.c_str(), pass_rays.c_str());
    } else {
        fprintf(stderr, "bad command\n");
        return 2;
    }
    return 0;
}
