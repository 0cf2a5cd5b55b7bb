/*
 * rtcuda_compat.cuh — source-level drop-in for the reference's scene /
 * primitive / triangle / material / camera / light headers plus render().
 *
 * A program written like the reference's main.cu (it builds Triangle,
 * Material, Light and Primitive arrays by hand, constructs `Bvh`, fills
 * `Scene`, creates a `Camera` and calls `render(...)`) compiles against this
 * one header and librtb.so instead of the reference's 19 headers; see
 * INTEGRATION.md.  Struct names, field names, field order and sizes are the
 * reference's; everything behind them goes through the C ABI
 * of rtb.h.  Only the HOST-side surface is provided: the device-side members
 * (Triangle::intersect, Material::sample_f, ...) live inside the library.
 *
 *   replaces: vec3.cuh, triangle.cuh:4-21, material.cuh:4-45, light.cuh:4-28,76-84,
 *             primitive.cuh, bvh.cuh:4-30 (constructor signature), scene.cuh,
 *             camera.cuh:4-29, render.cuh:366-367 (signature), utility.cuh:4-13,79-81
 */
#ifndef RTCUDA_COMPAT_CUH
#define RTCUDA_COMPAT_CUH

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "rtb.h"

#define CHECK_CUDA(val) rtcuda_compat::check_cuda((val), #val, __FILE__, __LINE__)

namespace rtcuda_compat {
inline void check_cuda(cudaError_t result, const char *func, const char *file, int line) {
    if (result) {
        fprintf(stderr, "CUDA error at %s:%d code=%d(%s) \"%s\" \n", file, line, (int)result, cudaGetErrorName(result), func);
        exit(EXIT_FAILURE);
    }
}
inline void check_rtb(int rc, const char *what) {
    if (rc != RTB_OK) {
        fprintf(stderr, "rtcuda_b200 error in %s: %d (%s)\n", what, rc, rtb_last_error());
        exit(EXIT_FAILURE);
    }
}
// RTB_DEVICES="0,1,..." -> device ordinals; the first must be the current device.  Unset or one entry: that device alone.
inline std::vector<int32_t> devices_from_env(int current) {
    std::vector<int32_t> d;
    if (const char *e = getenv("RTB_DEVICES")) {
        for (const char *p = e; *p;) {
            char *end = nullptr;
            long v = strtol(p, &end, 10);
            if (end == p) break;
            d.push_back((int32_t)v);
            p = *end == ',' ? end + 1 : end;
        }
    }
    if (d.empty()) d.push_back(current);
    if (d[0] != current) {
        fprintf(stderr, "RTB_DEVICES must start with the current device %d (the caller's device arrays live there)\n", current);
        exit(EXIT_FAILURE);
    }
    return d;
}
}  // namespace rtcuda_compat

struct Vec3 {
    Vec3() {}
    constexpr Vec3(float x, float y, float z) : x(x), y(y), z(z) {}
    constexpr Vec3(float xyz) : x(xyz), y(xyz), z(xyz) {}
    Vec3 operator-() const { return Vec3(-x, -y, -z); }
    float length() const { return sqrtf(x * x + y * y + z * z); }
    Vec3 unit_vector() const { float i = 1.f / length(); return Vec3(x * i, y * i, z * i); }
    static Vec3 make_zeros() { return Vec3(0.f, 0.f, 0.f); }
    static Vec3 make_ones() { return Vec3(1.f, 1.f, 1.f); }
    float x, y, z;
};
inline Vec3 operator+(const Vec3 &a, const Vec3 &b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(const Vec3 &a, const Vec3 &b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator*(const Vec3 &a, float t) { return Vec3(a.x * t, a.y * t, a.z * t); }
inline Vec3 operator*(float t, const Vec3 &a) { return Vec3(a.x * t, a.y * t, a.z * t); }
inline float dot(const Vec3 &a, const Vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(const Vec3 &a, const Vec3 &b) { return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

struct Triangle {
    Triangle() {}
    Triangle(const Vec3 &p0, const Vec3 &p1, const Vec3 &p2) : p0(p0), e1(p0 - p1), e2(p2 - p0), n(cross(e1, e2)) {}
    Vec3 p1() const { return p0 - e1; }
    Vec3 p2() const { return p0 + e2; }
    Vec3 p0, e1, e2, n;
};

enum MaterialType { MATTE, MIRROR, GLASS };
struct Material {
    Material() {}
    static Material make_matte(const Vec3 &albedo) { Material m; m.albedo = albedo; m.index_of_refraction = 0.f; m.type = MATTE; return m; }
    static Material make_mirror(const Vec3 &albedo) { Material m; m.albedo = albedo; m.index_of_refraction = 0.f; m.type = MIRROR; return m; }
    static Material make_glass(float ior) { Material m; m.albedo = Vec3(0.f); m.index_of_refraction = ior; m.type = GLASS; return m; }
    Vec3 albedo;
    float index_of_refraction;
    MaterialType type;
};

enum LightType { POINT_LIGHT, AREA_LIGHT };
struct Light {
    Light() {}
    static Light make_point_light(const Vec3 &pos, const Vec3 &I) { Light l; l.type = POINT_LIGHT; l.pos = pos; l.d_triangle = nullptr; l.I = I; return l; }
    static Light make_area_light(Triangle *d_triangle, const Vec3 &L) { Light l; l.type = AREA_LIGHT; l.pos = Vec3(0.f); l.d_triangle = d_triangle; l.L = L; return l; }
    LightType type;
    Vec3 pos;
    Triangle *d_triangle;
    union {
        Vec3 I;
        Vec3 L;
    };
};

struct Primitive {
    Primitive() {}
    Primitive(Triangle *d_triangle, Material *d_mat, Light *d_area_light = NULL)
        : d_triangle(d_triangle), d_mat(d_mat), d_area_light(d_area_light) {}
    Triangle *d_triangle;
    Material *d_mat;
    Light *d_area_light;
};

static_assert(sizeof(Vec3) == 12 && sizeof(Triangle) == 48 && sizeof(Material) == 20 && sizeof(Light) == 40 && sizeof(Primitive) == 24,
              "layouts must match the reference (SURVEY.md 8a)");

// Bvh keeps the reference's constructor signature and its public fields (bvh.cuh:17-27).  Like the reference's, the
// constructor BUILDS the tree (on the GPU, through rtb_scene_create_from_primitives with the lights still to come:
// `Scene` is filled in afterwards, main.cu:151-156), so num_nodes / max_depth are valid when it returns; the light
// array is attached at the first render().
//
// Several GPUs: with the environment variable RTB_DEVICES="0,1,2,3" (first entry = the current device, where the
// caller's device arrays live) render() splits the samples over those GPUs (rtb_multi_render: scene copied device to
// device, per-GPU accumulation, one NCCL reduction, tonemap on the first GPU).  Unset: one GPU, as in the reference.
struct Bvh {
    struct Impl {
        rtb_context *ctx = nullptr;
        rtb_scene *scene = nullptr;
        rtb_multi *multi = nullptr;
        rtb_multi_scene *replicas = nullptr;
        const void *lights_attached = nullptr;
        int num_lights_attached = -1;
    };
    Bvh() {}
    Bvh(const std::vector<Triangle> &triangles, const std::vector<Primitive> &primitives)
        : num_primitives((int)triangles.size()), impl(new Impl()) {
        if (triangles.size() != primitives.size()) { fprintf(stderr, "Bvh: triangles/primitives size mismatch\n"); exit(EXIT_FAILURE); }
        int dev = 0;
        CHECK_CUDA(cudaGetDevice(&dev));
        std::vector<int32_t> devices = rtcuda_compat::devices_from_env(dev);
        if (devices.size() > 1) {
            rtcuda_compat::check_rtb(rtb_multi_create(devices.data(), (int32_t)devices.size(), &impl->multi), "rtb_multi_create");
            rtcuda_compat::check_rtb(rtb_multi_context(impl->multi, 0, &impl->ctx), "rtb_multi_context");
        } else {
            rtcuda_compat::check_rtb(rtb_context_create(dev, &impl->ctx), "rtb_context_create");
        }
        const Triangle *tri_base = nullptr;
        const Material *mat_lo = nullptr, *mat_hi = nullptr;
        for (const Primitive &p : primitives) {
            if (!tri_base || p.d_triangle < tri_base) tri_base = p.d_triangle;
            if (!mat_lo || p.d_mat < mat_lo) mat_lo = p.d_mat;
            if (!mat_hi || p.d_mat > mat_hi) mat_hi = p.d_mat;
        }
        const int num_materials = mat_lo ? (int)(mat_hi - mat_lo) + 1 : 0;
        rtcuda_compat::check_rtb(rtb_scene_create_from_primitives(impl->ctx, primitives.data(), (int64_t)primitives.size(), tri_base, mat_lo,
                                                                  num_materials, nullptr, -1, nullptr, &impl->scene),
                                 "rtb_scene_create_from_primitives");
        rtb_bvh_stats st;
        rtcuda_compat::check_rtb(rtb_scene_stats(impl->scene, &st), "rtb_scene_stats");
        num_nodes = (int)st.num_nodes;           // 80-byte 8-wide nodes (the reference counts its 32-byte binary nodes)
        max_depth = (int)st.collapse_levels;     // levels of the 8-wide tree
    }
    // the device scene with Scene::d_lights attached (scene.cuh:4-8)
    rtb_scene *device_scene(int num_lights, Light *d_lights) const {
        if (impl->lights_attached != d_lights || impl->num_lights_attached != num_lights) {
            if (impl->replicas) { rtb_multi_scene_destroy(impl->replicas); impl->replicas = nullptr; }
            rtcuda_compat::check_rtb(rtb_scene_attach_lights(impl->scene, d_lights, num_lights), "rtb_scene_attach_lights");
            impl->lights_attached = d_lights; impl->num_lights_attached = num_lights;
        }
        return impl->scene;
    }
    rtb_multi_scene *device_replicas(int num_lights, Light *d_lights) const {
        rtb_scene *s = device_scene(num_lights, d_lights);
        if (!impl->replicas) rtcuda_compat::check_rtb(rtb_multi_scene_replicate(impl->multi, s, &impl->replicas), "rtb_multi_scene_replicate");
        return impl->replicas;
    }
    int num_primitives = 0;
    int num_nodes = 0;
    int max_depth = 0;
    Impl *impl = nullptr;
};

struct Scene {
    Bvh bvh;
    int num_lights;
    Light *d_lights;
};

struct Camera {
    Camera() {}
    Camera(Vec3 lookfrom, Vec3 lookat, Vec3 up, float vfov, float aspect_ratio) {
        rtb_camera c;
        rtcuda_compat::check_rtb(rtb_camera_look_at(&lookfrom.x, &lookat.x, &up.x, vfov, aspect_ratio, &c), "rtb_camera_look_at");
        this->lookfrom = Vec3(c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]);
        upper_left = Vec3(c.upper_left[0], c.upper_left[1], c.upper_left[2]);
        horizontal = Vec3(c.horizontal[0], c.horizontal[1], c.horizontal[2]);
        vertical = Vec3(c.vertical[0], c.vertical[1], c.vertical[2]);
    }
    Vec3 lookfrom, upper_left, horizontal, vertical;
};
static_assert(sizeof(Camera) == sizeof(rtb_camera), "Camera layout");

inline int clamp(int value, int low, int high) { return value < low ? low : (value > high ? high : value); }

// render(), render.cuh:366-367: same signature, same framebuffer contract
// (sqrt(mean radiance), row 0 = top, W*H Vec3).
inline void render(int width, int height, int num_samples, int max_bounces, Camera camera, Scene scene, std::vector<Vec3> &framebuffer) {
    rtb_scene *s = scene.bvh.device_scene(scene.num_lights, scene.d_lights);
    rtb_render_params p;
    rtb_render_params_default(&p);
    p.width = width; p.height = height; p.spp = num_samples; p.max_bounces = max_bounces;
    framebuffer.resize((size_t)width * (size_t)height);
    if (scene.bvh.impl->multi) {
        rtb_multi_scene *ms = scene.bvh.device_replicas(scene.num_lights, scene.d_lights);
        rtcuda_compat::check_rtb(rtb_multi_render(ms, reinterpret_cast<const rtb_camera *>(&camera), &p, &framebuffer.data()->x, nullptr), "rtb_multi_render");
        return;
    }
    rtcuda_compat::check_rtb(rtb_render(s, reinterpret_cast<const rtb_camera *>(&camera), &p, &framebuffer.data()->x, nullptr), "rtb_render");
}

#endif  // RTCUDA_COMPAT_CUH
