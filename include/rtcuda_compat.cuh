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

// Bvh keeps the reference's constructor signature.  The device BVH is built on
// the GPU the first time the scene is rendered (the light array is only known
// once `Scene` is filled in, scene.cuh:4-8).
struct Bvh {
    struct Impl {
        std::vector<Primitive> primitives;
        rtb_context *ctx = nullptr;
        rtb_scene *scene = nullptr;
        const void *built_for_lights = nullptr;
    };
    Bvh() {}
    Bvh(const std::vector<Triangle> &triangles, const std::vector<Primitive> &primitives)
        : num_primitives((int)triangles.size()), impl(new Impl()) {
        if (triangles.size() != primitives.size()) { fprintf(stderr, "Bvh: triangles/primitives size mismatch\n"); exit(EXIT_FAILURE); }
        impl->primitives = primitives;
    }
    rtb_scene *device_scene(int num_lights, Light *d_lights) const {
        if (impl->scene && impl->built_for_lights == d_lights) return impl->scene;
        if (impl->scene) rtb_scene_destroy(impl->scene);
        if (!impl->ctx) {
            int dev = 0;
            CHECK_CUDA(cudaGetDevice(&dev));
            rtcuda_compat::check_rtb(rtb_context_create(dev, &impl->ctx), "rtb_context_create");
        }
        const Triangle *tri_base = nullptr;
        const Material *mat_lo = nullptr, *mat_hi = nullptr;
        for (const Primitive &p : impl->primitives) {
            if (!tri_base || p.d_triangle < tri_base) tri_base = p.d_triangle;
            if (!mat_lo || p.d_mat < mat_lo) mat_lo = p.d_mat;
            if (!mat_hi || p.d_mat > mat_hi) mat_hi = p.d_mat;
        }
        const int num_materials = mat_lo ? (int)(mat_hi - mat_lo) + 1 : 0;
        rtcuda_compat::check_rtb(rtb_scene_create_from_primitives(impl->ctx, impl->primitives.data(), (int64_t)impl->primitives.size(), tri_base,
                                                                  mat_lo, num_materials, d_lights, num_lights, nullptr, &impl->scene),
                                 "rtb_scene_create_from_primitives");
        impl->built_for_lights = d_lights;
        return impl->scene;
    }
    int num_primitives = 0;
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
    rtcuda_compat::check_rtb(rtb_render(s, reinterpret_cast<const rtb_camera *>(&camera), &p, &framebuffer.data()->x, nullptr), "rtb_render");
}

#endif  // RTCUDA_COMPAT_CUH
