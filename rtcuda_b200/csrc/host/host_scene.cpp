// host_scene.cpp — host-side scene preparation (no GPU work): mesh I/O, the
// reference's default scene and the procedural benchmark scenes, camera
// set-up, PPM output.  Compiled with -ffp-contract=off: this code must round
// like the reference's HOST code (g++ does not fuse on x86-64).
//
// Restates main.cu:41-166 (scene + camera), transform.hpp:13-33 and
// matrix4x4.hpp:22-34 (vertex transform arithmetic), camera.cuh:15-29 and the
// PPM writer main.cu:178-191.  The PLY reader covers what main.cu:60-62 uses
// of happly.h (ASCII, float vertices, uchar-count face lists).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "rtb.h"
#include "host_util.h"

namespace {

struct M4 {
    float m[4][4];
};
M4 translate(float x, float y, float z) {  // matrix4x4.hpp:22-27
    return M4{{{1, 0, 0, x}, {0, 1, 0, y}, {0, 0, 1, z}, {0, 0, 0, 1}}};
}
M4 scale(float x, float y, float z) {  // matrix4x4.hpp:29-34
    return M4{{{x, 0, 0, 0}, {0, y, 0, 0}, {0, 0, z, 0}, {0, 0, 0, 1}}};
}
M4 rotate_y(float theta) {  // matrix4x4.hpp:36-56 with axis (0,1,0)
    float c = cosf(theta), s = sinf(theta);
    return M4{{{c, 0, s, 0}, {0, 1, 0, 0}, {-s, 0, c, 0}, {0, 0, 0, 1}}};
}
// Transform::composite, transform.hpp:13-24: result = other x current, float accumulation
M4 composite(const M4 &cur, const M4 &other) {
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            r.m[i][j] = 0;
            for (int k = 0; k < 4; ++k) r.m[i][j] += other.m[i][k] * cur.m[k][j];
        }
    return r;
}
// Transform::apply, transform.hpp:26-33: double arithmetic on a float matrix;
// x and y pass through a float temporary, z stays double until Vec3(...)
void apply(const M4 &t, double v[3]) {
    float nx = t.m[0][0] * v[0] + t.m[0][1] * v[1] + t.m[0][2] * v[2] + t.m[0][3];
    float ny = t.m[1][0] * v[0] + t.m[1][1] * v[1] + t.m[1][2] * v[2] + t.m[1][3];
    v[2] = t.m[2][0] * v[0] + t.m[2][1] * v[1] + t.m[2][2] * v[2] + t.m[2][3];
    v[0] = nx;
    v[1] = ny;
}

struct Builder {
    std::vector<float> verts;
    std::vector<int32_t> mat, light;
    std::vector<rtb_material> materials;
    std::vector<rtb_light> lights;
    void tri(const float a[3], const float b[3], const float c[3], int m) {
        verts.insert(verts.end(), a, a + 3);
        verts.insert(verts.end(), b, b + 3);
        verts.insert(verts.end(), c, c + 3);
        mat.push_back(m);
        light.push_back(-1);
    }
    void tri(float ax, float ay, float az, float bx, float by, float bz, float cx, float cy, float cz, int m) {
        float a[3] = {ax, ay, az}, b[3] = {bx, by, bz}, c[3] = {cx, cy, cz};
        tri(a, b, c, m);
    }
    int add_material(int type, float r, float g, float b, float ior) {
        rtb_material m;
        m.albedo[0] = r; m.albedo[1] = g; m.albedo[2] = b; m.ior = ior; m.type = type;
        materials.push_back(m);
        return (int)materials.size() - 1;
    }
    void make_area_light(int64_t tri_index, float L) {  // Light::make_area_light, light.cuh:78-84
        rtb_light l;
        memset(&l, 0, sizeof l);
        l.type = RTB_AREA_LIGHT;
        l.triangle = tri_index;
        l.L[0] = l.L[1] = l.L[2] = L;
        lights.push_back(l);
        light[tri_index] = (int)lights.size() - 1;
    }
};

// walls + the two emissive triangles, main.cu:88-116
void cornell_shell(Builder &b, int red, int green, int white) {
    b.tri(0, 0, 0, 0, 0, -1, 0, 1, -1, red);
    b.tri(0, 0, 0, 0, 1, 0, 0, 1, -1, red);
    b.tri(1, 0, 0, 1, 0, -1, 1, 1, -1, green);
    b.tri(1, 0, 0, 1, 1, 0, 1, 1, -1, green);
    b.tri(0, 0, 0, 1, 0, 0, 1, 0, -1, white);
    b.tri(0, 0, 0, 0, 0, -1, 1, 0, -1, white);
    b.tri(0, 1, 0, 1, 1, 0, 1, 1, -1, white);
    b.tri(0, 1, 0, 0, 1, -1, 1, 1, -1, white);
    b.tri(0, 0, -1, 1, 0, -1, 1, 1, -1, white);
    b.tri(0, 0, -1, 0, 1, -1, 1, 1, -1, white);
    b.tri(0.4f, 0.999f, -0.4f, 0.6f, 0.999f, -0.4f, 0.6f, 0.999f, -0.6f, white);
    b.make_area_light((int64_t)b.mat.size() - 1, 15.f);
    b.tri(0.4f, 0.999f, -0.4f, 0.4f, 0.999f, -0.6f, 0.6f, 0.999f, -0.6f, white);
    b.make_area_light((int64_t)b.mat.size() - 1, 15.f);
}

void add_mesh(Builder &b, const M4 &t, const float *mv, int64_t nv, const int32_t *mf, int64_t nf, int mat) {
    std::vector<float> tv((size_t)nv * 3);
    for (int64_t i = 0; i < nv; ++i) {
        double v[3] = {mv[3 * i], mv[3 * i + 1], mv[3 * i + 2]};
        apply(t, v);
        tv[3 * i] = (float)v[0]; tv[3 * i + 1] = (float)v[1]; tv[3 * i + 2] = (float)v[2];
    }
    for (int64_t f = 0; f < nf; ++f)
        b.tri(&tv[3 * (size_t)mf[3 * f]], &tv[3 * (size_t)mf[3 * f + 1]], &tv[3 * (size_t)mf[3 * f + 2]], mat);
}

uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s; }
float lcg01(uint32_t &s) { return (float)(lcg(s) >> 8) * (1.0f / 16777216.0f); }

}  // namespace

struct rtb_host_scene {
    Builder b;
    float lookfrom[3], lookat[3], up[3], vfov;
    // instanced form (rtb_host_scene_build_instanced): b holds the meshes in object space
    std::vector<int64_t> mesh_first;
    std::vector<rtb_instance> instances;
};

namespace {
void default_camera(rtb_host_scene *hs) {  // main.cu:162-166
    hs->lookfrom[0] = 0.5f; hs->lookfrom[1] = 0.5f; hs->lookfrom[2] = 1.5f;
    hs->lookat[0] = 0.5f; hs->lookat[1] = 0.5f; hs->lookat[2] = 0.f;
    hs->up[0] = 0.f; hs->up[1] = 1.f; hs->up[2] = 0.f;
    hs->vfov = 37.8f;
}
rtb_instance make_instance(int mesh, int material, const M4 &t) {
    rtb_instance in;
    in.mesh = mesh; in.material = material;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) in.xform[4 * r + c] = t.m[r][c];
    return in;
}
// the placement of bunny (gx, gz) of RTB_SCENE_S2: shared by the flat and the instanced generator
M4 s2_placement(const M4 &base, uint32_t &s, int gx, int gz, int grid) {
    const float cx = 0.1557f, cz = -0.12065f, foot = 0.3114f;
    const float cell = 1.f / (float)grid;
    const float ang = 6.2831853f * lcg01(s);
    const float sc = (0.70f + 0.25f * lcg01(s)) * cell / foot;
    const float jx = (lcg01(s) - 0.5f) * 0.15f * cell, jz = (lcg01(s) - 0.5f) * 0.15f * cell;
    M4 t = composite(base, translate(-cx, 0.f, -cz));
    t = composite(t, rotate_y(ang));
    t = composite(t, scale(sc, sc, sc));
    t = composite(t, translate(((float)gx + 0.5f) * cell + jx, 0.f, -((float)gz + 0.5f) * cell + jz));
    return t;
}
}  // namespace

extern "C" {

int rtb_host_scene_build(int32_t kind, const float *mv, int64_t nv, const int32_t *mf, int64_t nf, int32_t grid,
                         uint32_t seed, rtb_host_scene **out) {
    if (!out || !mv || !mf || nv <= 0 || nf <= 0) return rtb::set_error(RTB_ERR_INVALID, "rtb_host_scene_build: null mesh");
    *out = nullptr;
    for (int64_t i = 0; i < 3 * nf; ++i)
        if (mf[i] < 0 || mf[i] >= nv) return rtb::set_error(RTB_ERR_INVALID, "rtb_host_scene_build: face index out of range");
    return rtb::host_guarded(RTB_ERR_INVALID, [&]() -> int {
    rtb_host_scene *hs = new rtb_host_scene();
    Builder &b = hs->b;
    // materials, main.cu:42-45
    const int red = b.add_material(RTB_MATTE, 0.65f, 0.05f, 0.05f, 0.f);
    const int green = b.add_material(RTB_MATTE, 0.12f, 0.45f, 0.15f, 0.f);
    const int white = b.add_material(RTB_MATTE, 0.73f, 0.73f, 0.73f, 0.f);
    const int brown = b.add_material(RTB_MATTE, 0.62f, 0.57f, 0.54f, 0.f);
    // bunny placement, main.cu:67-71
    M4 base = translate(0.0946899f, -0.0329874f, -0.0587997f);
    base = composite(base, scale(2.f, 2.f, 2.f));
    if (kind == RTB_SCENE_S1 || kind == RTB_SCENE_S1_MIXED || kind == RTB_SCENE_S1_GLOSSY) {
        M4 t = composite(base, translate(0.3f, 0.f, -0.5f));
        add_mesh(b, t, mv, nv, mf, nf, brown);  // triangles [0, nf)
        cornell_shell(b, red, green, white);    // walls [nf, nf+10), lights nf+10, nf+11
        if (kind == RTB_SCENE_S1_MIXED) {
            // config C4: round-robin MATTE / mirror(0.9) / glass(1.5) by triangle
            // index over bunny + walls; the emissive triangles stay matte
            const int mirror = b.add_material(RTB_MIRROR, 0.9f, 0.9f, 0.9f, 0.f);
            const int glass = b.add_material(RTB_GLASS, 0.f, 0.f, 0.f, 1.5f);
            for (int64_t i = 0; i < nf + 10; ++i) {
                if (i % 3 == 1) b.mat[i] = mirror;
                else if (i % 3 == 2) b.mat[i] = glass;
            }
        }
        if (kind == RTB_SCENE_S1_GLOSSY) {
            // beyond the reference: the bunny a broad glossy lobe, the x = 1 wall (triangles nf+2, nf+3) a sharp one
            const int gloss_bunny = b.add_material(RTB_GLOSSY, 0.75f, 0.65f, 0.45f, 50.f);
            const int gloss_wall = b.add_material(RTB_GLOSSY, 0.85f, 0.85f, 0.85f, 400.f);
            for (int64_t i = 0; i < nf; ++i) b.mat[i] = gloss_bunny;
            b.mat[nf + 2] = gloss_wall; b.mat[nf + 3] = gloss_wall;
        }
    } else if (kind == RTB_SCENE_S2) {
        if (grid <= 0) grid = 12;
        // After `base` the bunny spans about [0,0.3114]x[0,0.3087]x[-0.2413,0];
        // each instance is rotated about Y around its footprint centre, scaled
        // into its floor cell and dropped on a jittered grid over the floor.
        uint32_t s = seed ? seed : 1234u;
        for (int gz = 0; gz < grid; ++gz)
            for (int gx = 0; gx < grid; ++gx) add_mesh(b, s2_placement(base, s, gx, gz, grid), mv, nv, mf, nf, brown);
        cornell_shell(b, red, green, white);
    } else {
        delete hs;
        return rtb::set_error(RTB_ERR_INVALID, "rtb_host_scene_build: unknown scene kind");
    }
    default_camera(hs);
    *out = hs;
    return RTB_OK;
    });
}

// Instanced form: the meshes stay in object space, the placements become instance transforms.
int rtb_host_scene_build_instanced(int32_t kind, const float *mv, int64_t nv, const int32_t *mf, int64_t nf, int32_t grid,
                                   uint32_t seed, rtb_host_scene **out) {
    if (!out || !mv || !mf || nv <= 0 || nf <= 0) return rtb::set_error(RTB_ERR_INVALID, "rtb_host_scene_build_instanced: null mesh");
    if (kind != RTB_SCENE_S2 && kind != RTB_SCENE_S1) return rtb::set_error(RTB_ERR_INVALID, "rtb_host_scene_build_instanced: RTB_SCENE_S1 or RTB_SCENE_S2");
    *out = nullptr;
    for (int64_t i = 0; i < 3 * nf; ++i)
        if (mf[i] < 0 || mf[i] >= nv) return rtb::set_error(RTB_ERR_INVALID, "rtb_host_scene_build_instanced: face index out of range");
    return rtb::host_guarded(RTB_ERR_INVALID, [&]() -> int {
    rtb_host_scene *hs = new rtb_host_scene();
    Builder &b = hs->b;
    const int red = b.add_material(RTB_MATTE, 0.65f, 0.05f, 0.05f, 0.f);
    const int green = b.add_material(RTB_MATTE, 0.12f, 0.45f, 0.15f, 0.f);
    const int white = b.add_material(RTB_MATTE, 0.73f, 0.73f, 0.73f, 0.f);
    const int brown = b.add_material(RTB_MATTE, 0.62f, 0.57f, 0.54f, 0.f);
    // mesh 0: the bunny as loaded
    for (int64_t f = 0; f < nf; ++f) b.tri(&mv[3 * (size_t)mf[3 * f]], &mv[3 * (size_t)mf[3 * f + 1]], &mv[3 * (size_t)mf[3 * f + 2]], brown);
    hs->mesh_first.push_back(0);
    hs->mesh_first.push_back((int64_t)b.mat.size());
    // mesh 1: walls + emitters
    cornell_shell(b, red, green, white);
    hs->mesh_first.push_back((int64_t)b.mat.size());
    M4 base = translate(0.0946899f, -0.0329874f, -0.0587997f);
    base = composite(base, scale(2.f, 2.f, 2.f));
    if (kind == RTB_SCENE_S1) {
        hs->instances.push_back(make_instance(0, -1, composite(base, translate(0.3f, 0.f, -0.5f))));
    } else {
        if (grid <= 0) grid = 12;
        uint32_t s = seed ? seed : 1234u;
        for (int gz = 0; gz < grid; ++gz)
            for (int gx = 0; gx < grid; ++gx) hs->instances.push_back(make_instance(0, -1, s2_placement(base, s, gx, gz, grid)));
    }
    hs->instances.push_back(make_instance(1, -1, translate(0.f, 0.f, 0.f)));
    default_camera(hs);
    *out = hs;
    return RTB_OK;
    });
}

int rtb_host_scene_instanced_desc(const rtb_host_scene *hs, rtb_instanced_scene_desc *d) {
    if (!hs || !d) return rtb::set_error(RTB_ERR_INVALID, "rtb_host_scene_instanced_desc: null");
    if (hs->instances.empty()) return rtb::set_error(RTB_ERR_INVALID, "rtb_host_scene_instanced_desc: not an instanced scene");
    rtb_host_scene_desc(hs, &d->geometry);
    d->num_meshes = (int32_t)hs->mesh_first.size() - 1;
    d->mesh_first = hs->mesh_first.data();
    d->num_instances = (int32_t)hs->instances.size();
    d->instances = hs->instances.data();
    return RTB_OK;
}

// Every instance's triangles through its transform with the reference's vertex arithmetic (Transform::apply,
// transform.hpp:26-33), instance after instance; lights follow their triangles.
int rtb_instanced_flatten(const rtb_instanced_scene_desc *D, rtb_host_scene **out) {
    if (!D || !out || !D->mesh_first || !D->instances || D->num_meshes <= 0 || D->num_instances <= 0 || !D->geometry.vertices ||
        !D->geometry.material_ids)
        return rtb::set_error(RTB_ERR_INVALID, "rtb_instanced_flatten: incomplete description");
    const rtb_scene_desc &g = D->geometry;
    *out = nullptr;
    if (g.num_triangles < 0 || g.num_materials <= 0 || !g.materials || (g.num_lights > 0 && !g.lights) || D->mesh_first[0] != 0 ||
        D->mesh_first[D->num_meshes] != g.num_triangles)
        return rtb::set_error(RTB_ERR_INVALID, "rtb_instanced_flatten: inconsistent description");
    for (int m = 0; m < D->num_meshes; ++m)
        if (D->mesh_first[m + 1] < D->mesh_first[m]) return rtb::set_error(RTB_ERR_INVALID, "rtb_instanced_flatten: mesh_first must ascend");
    return rtb::host_guarded(RTB_ERR_INVALID, [&]() -> int {
    rtb_host_scene *hs = new rtb_host_scene();
    Builder &b = hs->b;
    b.materials.assign(g.materials, g.materials + g.num_materials);
    if (g.num_lights) b.lights.assign(g.lights, g.lights + g.num_lights);
    for (int i = 0; i < D->num_instances; ++i) {
        const rtb_instance &in = D->instances[i];
        if (in.mesh < 0 || in.mesh >= D->num_meshes) { delete hs; return rtb::set_error(RTB_ERR_INVALID, "rtb_instanced_flatten: mesh out of range"); }
        M4 t = translate(0.f, 0.f, 0.f);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) t.m[r][c] = in.xform[4 * r + c];
        for (int64_t k = D->mesh_first[in.mesh]; k < D->mesh_first[in.mesh + 1]; ++k) {
            float p[3][3];
            for (int v = 0; v < 3; ++v) {
                double w[3] = {g.vertices[9 * k + 3 * v], g.vertices[9 * k + 3 * v + 1], g.vertices[9 * k + 3 * v + 2]};
                apply(t, w);
                p[v][0] = (float)w[0]; p[v][1] = (float)w[1]; p[v][2] = (float)w[2];
            }
            b.tri(p[0], p[1], p[2], in.material >= 0 ? in.material : g.material_ids[k]);
            const int l = g.light_ids ? g.light_ids[k] : -1;
            if (l >= 0) {
                b.light.back() = l;
                if (l < g.num_lights && b.lights[(size_t)l].type == RTB_AREA_LIGHT) b.lights[(size_t)l].triangle = (int64_t)b.mat.size() - 1;
            }
        }
    }
    default_camera(hs);
    *out = hs;
    return RTB_OK;
    });
}

int rtb_host_scene_desc(const rtb_host_scene *hs, rtb_scene_desc *d) {
    if (!hs || !d) return rtb::set_error(RTB_ERR_INVALID, "rtb_host_scene_desc: null");
    const Builder &b = hs->b;
    d->num_triangles = (int64_t)b.mat.size();
    d->vertices = b.verts.data();
    d->material_ids = b.mat.data();
    d->light_ids = b.light.data();
    d->num_materials = (int32_t)b.materials.size();
    d->materials = b.materials.data();
    d->num_lights = (int32_t)b.lights.size();
    d->lights = b.lights.data();
    return RTB_OK;
}

int rtb_host_scene_camera(const rtb_host_scene *hs, float aspect, rtb_camera *out) {
    if (!hs || !out) return rtb::set_error(RTB_ERR_INVALID, "rtb_host_scene_camera: null");
    return rtb_camera_look_at(hs->lookfrom, hs->lookat, hs->up, hs->vfov, aspect, out);
}

int rtb_host_scene_destroy(rtb_host_scene *hs) {
    delete hs;
    return RTB_OK;
}

// Camera::Camera, camera.cuh:15-29 (host float arithmetic, no contraction)
int rtb_camera_look_at(const float from[3], const float at[3], const float up[3], float vfov_deg, float aspect,
                       rtb_camera *c) {
    if (!c) return rtb::set_error(RTB_ERR_INVALID, "rtb_camera_look_at: null");
    const float PI = 3.14159265358979323846f;
    float vfov_rad = vfov_deg * (PI / 180.f);  // deg_to_rad, utility.cuh:15-17
    float vh = 2.f * tanf(vfov_rad * 0.5f);
    float vw = vh * aspect;
    float w[3] = {from[0] - at[0], from[1] - at[1], from[2] - at[2]};
    float il = 1.f / sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    for (int i = 0; i < 3; ++i) w[i] *= il;
    float d = up[0] * w[0] + up[1] * w[1] + up[2] * w[2];
    float v[3] = {up[0] - d * w[0], up[1] - d * w[1], up[2] - d * w[2]};
    il = 1.f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int i = 0; i < 3; ++i) v[i] *= il;
    float u[3] = {v[1] * w[2] - v[2] * w[1], v[2] * w[0] - v[0] * w[2], v[0] * w[1] - v[1] * w[0]};
    for (int i = 0; i < 3; ++i) {
        c->lookfrom[i] = from[i];
        c->horizontal[i] = u[i] * vw;
        c->vertical[i] = v[i] * -vh;
    }
    for (int i = 0; i < 3; ++i) {
        // lookfrom - w - 0.5f*horizontal - 0.5f*vertical, left to right
        float t = from[i] - w[i];
        t = t - c->horizontal[i] * 0.5f;
        t = t - c->vertical[i] * 0.5f;
        c->upper_left[i] = t;
    }
    return RTB_OK;
}

// ------------------------------------------------------------------ mesh I/O
// size in bytes of a PLY scalar type name, 0 = unknown
static int ply_type_size(const std::string &t) {
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    return 0;
}
static bool ply_type_signed(const std::string &t) { return t == "char" || t == "int8" || t == "short" || t == "int16" || t == "int" || t == "int32"; }
// one little-endian integer of `size` bytes at p (the library is built for little-endian hosts only)
static int64_t ply_read_int(const unsigned char *p, int size, bool is_signed) {
    uint64_t v = 0;
    for (int k = 0; k < size; ++k) v |= (uint64_t)p[k] << (8 * k);
    if (is_signed && size < 8 && (v >> (8 * size - 1) & 1u)) v |= ~0ull << (8 * size);
    return (int64_t)v;
}
// ASCII (the form of the reference's bun_zipper.ply, read there by happly: main.cu:60) and binary_little_endian PLY:
// vertex element with x y z as its first three properties (float32, or float64 in a binary file), any further scalar
// properties skipped; face element whose first property is the vertex index list; polygons are fanned into triangles
int rtb_mesh_load_ply(const char *path, float **verts_out, int64_t *nv_out, int32_t **faces_out, int64_t *nf_out) {
    if (!path || !verts_out || !nv_out || !faces_out || !nf_out) return rtb::set_error(RTB_ERR_INVALID, "rtb_mesh_load_ply: null argument");
    return rtb::host_guarded(RTB_ERR_IO, [&]() -> int {
    std::ifstream in(path, std::ios::binary);
    if (!in) return rtb::set_error(RTB_ERR_IO, std::string("cannot open ") + path);
    std::string line, tok;
    int64_t nv = -1, nf = -1;
    int vprops = 0;
    bool ascii = false, binary_le = false;
    enum { kNone, kVertex, kFace, kOther } cur = kNone;
    std::vector<int> vsizes;                 // byte sizes of the vertex element's scalar properties
    int cnt_size = 0, idx_size = 0;          // face list: count type, index type
    bool idx_signed = true;
    int face_extra = 0;                      // bytes of scalar face properties behind the list
    bool face_list_first = false, unsupported = false;
    auto strip_cr = [](std::string &l) { if (!l.empty() && l.back() == '\r') l.pop_back(); };
    std::getline(in, line);
    strip_cr(line);
    if (line.substr(0, 3) != "ply") return rtb::set_error(RTB_ERR_IO, "not a PLY file");
    bool ended = false;
    while (std::getline(in, line)) {
        strip_cr(line);
        std::istringstream ss(line);
        tok.clear();
        ss >> tok;
        if (tok == "format") { ss >> tok; ascii = (tok == "ascii"); binary_le = (tok == "binary_little_endian"); }
        else if (tok == "element") {
            std::string name; int64_t cnt = -1;
            ss >> name >> cnt;
            cur = name == "vertex" ? kVertex : (name == "face" ? kFace : kOther);
            if (cur == kVertex) nv = cnt;
            else if (cur == kFace) { if (nv < 0) unsupported = true; nf = cnt; }  // (vertices must come first)
            else if (cnt != 0 && nf < 0) unsupported = true;  // an unknown element in front of the ones we read
        } else if (tok == "property") {
            std::string t;
            ss >> t;
            if (cur == kVertex) {
                if (t == "list") unsupported = true;
                vsizes.push_back(ply_type_size(t));
                vprops++;
            } else if (cur == kFace) {
                if (t == "list") {
                    std::string ct, it;
                    ss >> ct >> it;
                    if (cnt_size == 0 && face_extra == 0) { cnt_size = ply_type_size(ct); idx_size = ply_type_size(it); idx_signed = ply_type_signed(it); face_list_first = true; }
                    else unsupported = true;  // a second list per face
                } else {
                    if (!face_list_first) unsupported = true;
                    face_extra += ply_type_size(t);
                }
            }
        } else if (tok == "end_header") { ended = true; break; }
    }
    if (!ended || !(ascii || binary_le) || nv < 0 || nf < 0 || vprops < 3 || unsupported)
        return rtb::set_error(RTB_ERR_IO, "unsupported PLY (need ascii or binary_little_endian, vertex x y z first, then face lists)");
    if (nv > 0x7fffffff || nf > 0x7fffffff) return rtb::set_error(RTB_ERR_IO, "PLY: element count out of range");
    std::vector<float> verts(3 * (size_t)nv);
    std::vector<int32_t> faces;
    if (ascii) {
        for (int64_t i = 0; i < nv; ++i) {
            if (!std::getline(in, line)) return rtb::set_error(RTB_ERR_IO, "PLY: file ends inside the vertex list");
            std::istringstream ss(line);
            if (!(ss >> verts[3 * i] >> verts[3 * i + 1] >> verts[3 * i + 2]))  // x y z are the first three properties
                return rtb::set_error(RTB_ERR_IO, "PLY: bad vertex line");
        }
        for (int64_t i = 0; i < nf; ++i) {
            if (!std::getline(in, line)) return rtb::set_error(RTB_ERR_IO, "PLY: file ends inside the face list");
            std::istringstream ss(line);
            int k = -1;
            if (!(ss >> k) || k < 3 || k > 64) return rtb::set_error(RTB_ERR_IO, "PLY: bad face vertex count");
            int32_t idx[64];
            for (int j = 0; j < k; ++j)
                if (!(ss >> idx[j]) || idx[j] < 0 || idx[j] >= nv) return rtb::set_error(RTB_ERR_IO, "PLY: face index out of range");
            for (int j = 1; j + 1 < k; ++j) { faces.push_back(idx[0]); faces.push_back(idx[j]); faces.push_back(idx[j + 1]); }
        }
    } else {
        int stride = 0;
        for (int sz : vsizes) { if (sz == 0) return rtb::set_error(RTB_ERR_IO, "PLY: unknown vertex property type"); stride += sz; }
        const int csz = vsizes[0];
        if ((csz != 4 && csz != 8) || vsizes[1] != csz || vsizes[2] != csz) return rtb::set_error(RTB_ERR_IO, "PLY: x y z must be float32 or float64");
        if (cnt_size == 0 || idx_size == 0 || idx_size > 4) return rtb::set_error(RTB_ERR_IO, "PLY: unsupported face list types");
        std::vector<unsigned char> row((size_t)stride);
        for (int64_t i = 0; i < nv; ++i) {
            if (!in.read((char *)row.data(), stride)) return rtb::set_error(RTB_ERR_IO, "PLY: file ends inside the vertex list");
            for (int a = 0; a < 3; ++a) {
                if (csz == 4) { float f; memcpy(&f, row.data() + 4 * a, 4); verts[3 * i + a] = f; }
                else { double d; memcpy(&d, row.data() + 8 * a, 8); verts[3 * i + a] = (float)d; }
            }
        }
        unsigned char buf[64 * 4 + 8];
        std::vector<char> skip((size_t)face_extra);
        for (int64_t i = 0; i < nf; ++i) {
            if (!in.read((char *)buf, cnt_size)) return rtb::set_error(RTB_ERR_IO, "PLY: file ends inside the face list");
            const int64_t k = ply_read_int(buf, cnt_size, false);
            if (k < 3 || k > 64) return rtb::set_error(RTB_ERR_IO, "PLY: bad face vertex count");
            if (!in.read((char *)buf, (std::streamsize)(k * idx_size))) return rtb::set_error(RTB_ERR_IO, "PLY: file ends inside the face list");
            int32_t idx[64];
            for (int j = 0; j < k; ++j) {
                const int64_t v = ply_read_int(buf + j * idx_size, idx_size, idx_signed);
                if (v < 0 || v >= nv) return rtb::set_error(RTB_ERR_IO, "PLY: face index out of range");
                idx[j] = (int32_t)v;
            }
            for (int j = 1; j + 1 < k; ++j) { faces.push_back(idx[0]); faces.push_back(idx[j]); faces.push_back(idx[j + 1]); }
            if (face_extra && !in.read(skip.data(), face_extra)) return rtb::set_error(RTB_ERR_IO, "PLY: file ends inside the face list");
        }
    }
    float *v = (float *)malloc(sizeof(float) * (verts.size() ? verts.size() : 1));
    int32_t *f = (int32_t *)malloc(sizeof(int32_t) * (faces.size() ? faces.size() : 1));
    if (!v || !f) { free(v); free(f); return rtb::set_error(RTB_ERR_OOM, "out of host memory"); }
    memcpy(v, verts.data(), sizeof(float) * verts.size());
    memcpy(f, faces.data(), sizeof(int32_t) * faces.size());
    *verts_out = v; *nv_out = nv; *faces_out = f; *nf_out = (int64_t)faces.size() / 3;
    return RTB_OK;
    });
}

int rtb_mesh_save_bin(const char *path, const float *verts, int64_t nv, const int32_t *faces, int64_t nf) {
    FILE *f = fopen(path, "wb");
    if (!f) return rtb::set_error(RTB_ERR_IO, std::string("cannot write ") + path);
    uint32_t hdr[3] = {0x4d425452u /* "RTBM" */, (uint32_t)nv, (uint32_t)nf};
    fwrite(hdr, 4, 3, f);
    fwrite(verts, 4, 3 * (size_t)nv, f);
    fwrite(faces, 4, 3 * (size_t)nf, f);
    fclose(f);
    return RTB_OK;
}

int rtb_mesh_load_bin(const char *path, float **verts_out, int64_t *nv_out, int32_t **faces_out, int64_t *nf_out) {
    if (!path || !verts_out || !nv_out || !faces_out || !nf_out) return rtb::set_error(RTB_ERR_INVALID, "rtb_mesh_load_bin: null argument");
    FILE *f = fopen(path, "rb");
    if (!f) return rtb::set_error(RTB_ERR_IO, std::string("cannot open ") + path);
    uint32_t hdr[3];
    if (fread(hdr, 4, 3, f) != 3 || hdr[0] != 0x4d425452u) { fclose(f); return rtb::set_error(RTB_ERR_IO, "bad mesh file"); }
    if (hdr[1] > 0x7fffffffu || hdr[2] > 0x7fffffffu) { fclose(f); return rtb::set_error(RTB_ERR_IO, "bad mesh file: counts out of range"); }
    float *v = (float *)malloc(12 * (size_t)(hdr[1] ? hdr[1] : 1));
    int32_t *fc = (int32_t *)malloc(12 * (size_t)(hdr[2] ? hdr[2] : 1));
    if (!v || !fc) { free(v); free(fc); fclose(f); return rtb::set_error(RTB_ERR_OOM, "out of host memory"); }
    bool ok = fread(v, 12, hdr[1], f) == hdr[1] && fread(fc, 12, hdr[2], f) == hdr[2];
    fclose(f);
    if (!ok) { free(v); free(fc); return rtb::set_error(RTB_ERR_IO, "truncated mesh file"); }
    for (size_t i = 0; i < 3 * (size_t)hdr[2]; ++i)
        if (fc[i] < 0 || (uint32_t)fc[i] >= hdr[1]) { free(v); free(fc); return rtb::set_error(RTB_ERR_IO, "bad mesh file: face index out of range"); }
    *verts_out = v; *nv_out = hdr[1]; *faces_out = fc; *nf_out = hdr[2];
    return RTB_OK;
}

void rtb_free(void *p) { free(p); }

// scene file shared with the reference harness (oracle/ref_harness.cu):
//   "RTBS" i64 num_triangles i32 num_materials i32 num_lights
//   f32 vertices[9n] i32 material_ids[n] i32 light_ids[n]
//   rtb_material[num_materials] rtb_light[num_lights]
int rtb_scene_desc_save(const char *path, const rtb_scene_desc *d) {
    if (!path || !d || d->num_triangles < 0 || (d->num_triangles > 0 && (!d->vertices || !d->material_ids || !d->light_ids)))
        return rtb::set_error(RTB_ERR_INVALID, "rtb_scene_desc_save: incomplete description (vertices, material_ids and light_ids are all written)");
    FILE *f = fopen(path, "wb");
    if (!f) return rtb::set_error(RTB_ERR_IO, std::string("cannot write ") + path);
    uint32_t magic = 0x53425452u;
    fwrite(&magic, 4, 1, f);
    fwrite(&d->num_triangles, 8, 1, f);
    fwrite(&d->num_materials, 4, 1, f);
    fwrite(&d->num_lights, 4, 1, f);
    fwrite(d->vertices, 4, 9 * (size_t)d->num_triangles, f);
    fwrite(d->material_ids, 4, (size_t)d->num_triangles, f);
    fwrite(d->light_ids, 4, (size_t)d->num_triangles, f);
    fwrite(d->materials, sizeof(rtb_material), (size_t)d->num_materials, f);
    fwrite(d->lights, sizeof(rtb_light), (size_t)d->num_lights, f);
    fclose(f);
    return RTB_OK;
}

int rtb_host_scene_load(const char *path, rtb_host_scene **out) {
    if (!path || !out) return rtb::set_error(RTB_ERR_INVALID, "rtb_host_scene_load: null argument");
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return rtb::set_error(RTB_ERR_IO, std::string("cannot open ") + path);
    uint32_t magic; int64_t n; int32_t nm, nl;
    bool ok = fread(&magic, 4, 1, f) == 1 && magic == 0x53425452u && fread(&n, 8, 1, f) == 1 &&
              fread(&nm, 4, 1, f) == 1 && fread(&nl, 4, 1, f) == 1;
    if (!ok) { fclose(f); return rtb::set_error(RTB_ERR_IO, "bad scene file"); }
    // the counts must be plausible for the file that holds them before anything is sized from them
    long here = ftell(f);
    fseek(f, 0, SEEK_END);
    const long long remaining = (long long)ftell(f) - here;
    fseek(f, here, SEEK_SET);
    if (n < 0 || nm < 0 || nl < 0 || n > 0x3fffffff || (long long)n * 44 + (long long)nm * 20 + (long long)nl * 40 != remaining) {
        fclose(f);
        return rtb::set_error(RTB_ERR_IO, "bad scene file: counts do not match the file size");
    }
    return rtb::host_guarded(RTB_ERR_IO, [&]() -> int {
    rtb_host_scene *hs = nullptr;
    try { hs = new rtb_host_scene(); } catch (...) { fclose(f); throw; }
    Builder &b = hs->b;
    try {
        b.verts.resize(9 * (size_t)n); b.mat.resize((size_t)n); b.light.resize((size_t)n);
        b.materials.resize((size_t)nm); b.lights.resize((size_t)nl);
    } catch (...) { fclose(f); delete hs; throw; }
    ok = fread(b.verts.data(), 4, b.verts.size(), f) == b.verts.size() && fread(b.mat.data(), 4, (size_t)n, f) == (size_t)n &&
         fread(b.light.data(), 4, (size_t)n, f) == (size_t)n &&
         fread(b.materials.data(), sizeof(rtb_material), (size_t)nm, f) == (size_t)nm &&
         fread(b.lights.data(), sizeof(rtb_light), (size_t)nl, f) == (size_t)nl;
    fclose(f);
    if (!ok) { delete hs; return rtb::set_error(RTB_ERR_IO, "truncated scene file"); }
    for (int64_t i = 0; i < n; ++i)
        if (b.mat[(size_t)i] < 0 || b.mat[(size_t)i] >= nm || b.light[(size_t)i] >= nl) { delete hs; return rtb::set_error(RTB_ERR_IO, "bad scene file: material / light id out of range"); }
    default_camera(hs);
    *out = hs;
    return RTB_OK;
    });
}

// main.cu:178-191
int rtb_write_ppm(const char *path, const float *rgb, int32_t w, int32_t h) {
    if (!path || !rgb || w <= 0 || h <= 0) return rtb::set_error(RTB_ERR_INVALID, "rtb_write_ppm: bad arguments");
    FILE *f = fopen(path, "w");
    if (!f) return rtb::set_error(RTB_ERR_IO, std::string("cannot write ") + path);
    fprintf(f, "P3\n%d %d\n255\n", w, h);
    for (int64_t i = 0; i < (int64_t)w * h; ++i) {
        int c[3];
        for (int k = 0; k < 3; ++k) {
            int v = (int)(256.f * rgb[3 * i + k]);
            c[k] = v < 0 ? 0 : (v > 255 ? 255 : v);
        }
        fprintf(f, "%d %d %d\n", c[0], c[1], c[2]);
    }
    fclose(f);
    return RTB_OK;
}

}  // extern "C"
