// compat_main.cu — TEST: a host program in the style of the reference's
// main.cu (hand-made device arrays of Triangle / Material / Light, host
// Primitives holding device pointers, Bvh, Scene, Camera, render()) compiled
// against include/rtcuda_compat.cuh + librtb.so.  It renders a small image of
// the scene file it is given and writes the raw float framebuffer, so the test
// can compare it with rtb_render() on the flat description of the same scene.
//   compat_main <scene.rtbs> <W> <H> <spp> <bounces> <out.f32>            main.cu-style program; RTB_DEVICES="0,1,.." spans GPUs
//   compat_main <scene.rtbs> <W> <H> <spp> <bounces> <out.f32> multi      rtb_render_multi() on every GPU of the box (C ABI, C++ caller)
#include <cstring>
#include <string>
#include <unordered_map>

#include "rtcuda_compat.cuh"

int main(int argc, char **argv) {
    if (argc < 7) return 2;
    rtb_host_scene *hs = nullptr;
    rtcuda_compat::check_rtb(rtb_host_scene_load(argv[1], &hs), "rtb_host_scene_load");
    rtb_scene_desc d;
    rtb_host_scene_desc(hs, &d);
    const int W = atoi(argv[2]), H = atoi(argv[3]), spp = atoi(argv[4]), bounces = atoi(argv[5]);

    if (argc >= 8 && std::string(argv[7]) == "multi") {
        // the one-call C entry on all GPUs of the box: contexts, concurrent scene builds, sample shards, NCCL reduce, tonemap
        int ndev = 0;
        CHECK_CUDA(cudaGetDeviceCount(&ndev));
        std::vector<int32_t> devs;
        for (int i = 0; i < ndev; ++i) devs.push_back(i);
        rtb_camera cam;
        const float from[3] = {0.5f, 0.5f, 1.5f}, at[3] = {0.5f, 0.5f, 0.0f}, up[3] = {0.0f, 1.0f, 0.0f};
        rtcuda_compat::check_rtb(rtb_camera_look_at(from, at, up, 37.8f, (float)W / (float)H, &cam), "rtb_camera_look_at");
        rtb_render_params p;
        rtb_render_params_default(&p);
        p.width = W; p.height = H; p.spp = spp; p.max_bounces = bounces;
        std::vector<float> fb(3 * (size_t)W * H);
        rtb_render_stats st;
        rtcuda_compat::check_rtb(rtb_render_multi(devs.data(), ndev, &d, nullptr, &cam, &p, fb.data(), &st), "rtb_render_multi");
        FILE *f = fopen(argv[6], "wb");
        fwrite(fb.data(), sizeof(float), fb.size(), f);
        fclose(f);
        printf("compat_main multi: %d GPU(s), %dx%d, %d spp, %llu paths, %.2f ms\n", ndev, W, H, spp, (unsigned long long)st.paths, st.ms_total);
        return 0;
    }

    std::vector<Material> materials;
    for (int i = 0; i < d.num_materials; ++i) {
        const rtb_material &m = d.materials[i];
        if (m.type == RTB_MATTE) materials.push_back(Material::make_matte(Vec3(m.albedo[0], m.albedo[1], m.albedo[2])));
        else if (m.type == RTB_MIRROR) materials.push_back(Material::make_mirror(Vec3(m.albedo[0], m.albedo[1], m.albedo[2])));
        else materials.push_back(Material::make_glass(m.ior));
    }
    Material *d_materials;
    CHECK_CUDA(cudaMalloc(&d_materials, materials.size() * sizeof(Material)));
    CHECK_CUDA(cudaMemcpy(d_materials, materials.data(), materials.size() * sizeof(Material), cudaMemcpyHostToDevice));

    std::vector<Triangle> triangles;
    for (int64_t i = 0; i < d.num_triangles; ++i) {
        const float *v = d.vertices + 9 * i;
        triangles.emplace_back(Vec3(v[0], v[1], v[2]), Vec3(v[3], v[4], v[5]), Vec3(v[6], v[7], v[8]));
    }
    Triangle *d_triangles;
    CHECK_CUDA(cudaMalloc(&d_triangles, triangles.size() * sizeof(Triangle)));
    CHECK_CUDA(cudaMemcpy(d_triangles, triangles.data(), triangles.size() * sizeof(Triangle), cudaMemcpyHostToDevice));

    std::vector<Light> lights;
    for (int i = 0; i < d.num_lights; ++i) {
        const rtb_light &l = d.lights[i];
        if (l.type == RTB_AREA_LIGHT) lights.push_back(Light::make_area_light(&d_triangles[l.triangle], Vec3(l.L[0], l.L[1], l.L[2])));
        else lights.push_back(Light::make_point_light(Vec3(l.pos[0], l.pos[1], l.pos[2]), Vec3(l.L[0], l.L[1], l.L[2])));
    }
    Light *d_lights = nullptr;
    int num_lights = (int)lights.size();
    if (num_lights) {
        CHECK_CUDA(cudaMalloc(&d_lights, lights.size() * sizeof(Light)));
        CHECK_CUDA(cudaMemcpy(d_lights, lights.data(), lights.size() * sizeof(Light), cudaMemcpyHostToDevice));
    }

    std::vector<Primitive> primitives;
    for (int64_t i = 0; i < d.num_triangles; ++i) {
        if (d.light_ids[i] >= 0) primitives.emplace_back(&d_triangles[i], &d_materials[d.material_ids[i]], &d_lights[d.light_ids[i]]);
        else primitives.emplace_back(&d_triangles[i], &d_materials[d.material_ids[i]]);
    }

    Bvh bvh(triangles, primitives);
    printf("compat_main: Bvh built: %d primitives, %d nodes, depth %d\n", bvh.num_primitives, bvh.num_nodes, bvh.max_depth);  // bvh.cuh:203-204
    if (bvh.num_nodes <= 0 || bvh.max_depth <= 0) return 3;
    Scene scene = {bvh, num_lights, d_lights};
    Camera camera(Vec3(0.5f, 0.5f, 1.5f), Vec3(0.5f, 0.5f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), 37.8f, (float)W / (float)H);
    std::vector<Vec3> framebuffer;
    render(W, H, spp, bounces, camera, scene, framebuffer);

    FILE *f = fopen(argv[6], "wb");
    fwrite(framebuffer.data(), sizeof(Vec3), framebuffer.size(), f);
    fclose(f);
    printf("compat_main: %dx%d, %d spp, first pixel %g %g %g\n", W, H, spp, framebuffer[0].x, framebuffer[0].y, framebuffer[0].z);
    return 0;
}
