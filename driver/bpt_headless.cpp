// Headless driver: what is left of SDL_main (Raytracer/raytracer.cpp:1560-2390) once the Win32/SDL/microui shell is
// removed -- build a scene, run progressive passes through the C ABI, "Take picture" (resolve + write_bitmap).
//
//   bpt_headless --tables sampler_tables.bin [--scene week3|icosphere|obj] [--obj file.obj [--cw]] [--hdr sky.hdr]
//                [--integrator "Whitted"|...] [--filter "Gaussian 3"|...] [--gpu-bvh] [--w W --h H] [--spp N] [--passes P]
//                [--level L] [--device D] [--gpus N] [--out file.bmp]
//
// --gpus N: the frame is split into interleaved 8-row blocks over N GPUs of this box (one host thread and one bpt_ctx per
// GPU, scene replicated), every progressive pass ends with ONE NCCL reduce of the partial films to GPU 0
// (bpt_reduce_film), and GPU 0 resolves the summed film -- the multi-GPU analogue of the reference's worker threads
// sharing one AccumulationBuffer (raytracer.cpp:551-603, :692-757).
//
// It only speaks include/bpt.h (plain C), exactly like a binding inside the reference would.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "../include/bpt.h"

static const float kDegToRad = 6.28318530717f / 360.0f;   // my_math.h:17

static bpt_m4x4inv translate(float x, float y, float z, float s = 1.0f) {
    bpt_m4x4inv m;
    memset(&m, 0, sizeof(m));
    for (int i = 0; i < 4; ++i) { m.forward.e[i][i] = i < 3 ? s : 1.0f; m.inverse.e[i][i] = i < 3 ? 1.0f/s : 1.0f; }
    m.forward.e[0][3] = x; m.forward.e[1][3] = y; m.forward.e[2][3] = z;
    m.inverse.e[0][3] = -x/s; m.inverse.e[1][3] = -y/s; m.inverse.e[2][3] = -z/s;   // (T*S)^-1 = S^-1 * T^-1
    return m;
}

static void set_camera(bpt_scene* s, uint32_t w, uint32_t h, float px, float py, float pz, float vfov_deg) {
    bpt_camera cam;
    bpt_get_camera(s, &cam);
    cam.vfov = kDegToRad*vfov_deg;
    cam.aspect_ratio = (float)w / (float)h;
    cam.lens_radius = 0.0f;
    cam.focus_distance = 1.0f;
    cam.p[0] = px; cam.p[1] = py; cam.p[2] = pz;
    bpt_set_camera(s, &cam);
}

static void use_advanced_integrator(bpt_scene* s) {
    bpt_settings st;
    bpt_get_settings(s, &st);
    st.integrator = bpt_find_integrator("Advanced Pathtracer");
    st.lens_distortion = 0.0f;
    bpt_set_settings(s, &st);
    bpt_load_reconstruction_kernel(s, "Mitchell Netravali");
}

// week_3_scene (raytracer.cpp:840-861) with the advanced integrator = BASELINE config 1
static void build_week3(bpt_scene* s, uint32_t w, uint32_t h) {
    set_camera(s, w, h, 0, 4, -10, 60.0f);
    const float d[3] = {0, 0, -1};
    bpt_aim_camera(s, d);
    use_advanced_integrator(s);
    const float white[3] = {1, 1, 1}, black[3] = {0, 0, 0}, red[3] = {1, 0, 0}, grey[3] = {0.1f, 0.1f, 0.1f}, e[3] = {12500, 12500, 12500};
    uint32_t ground = bpt_add_diffuse_material(s, white, 1.0f, 0.0f, 1, black);
    uint32_t sphere = bpt_add_diffuse_material(s, red, 1.0f, 0.0f, 0, grey);
    uint32_t light = bpt_add_emissive_material(s, e);
    const float up[3] = {0, 1, 0};
    bpt_add_plane(s, ground, up, 0.0f);
    bpt_m4x4inv t1 = translate(0, 4, 0), t2 = translate(8, 16, -8);
    bpt_add_sphere(s, sphere, 4.0f, &t1);
    bpt_add_sphere(s, light, 0.1f, &t2);
}

// displaced icosphere under the TLAS = BASELINE config 2 (level 8 = 1,310,720 triangles)
static void build_icosphere(bpt_scene* s, uint32_t w, uint32_t h, uint32_t level) {
    set_camera(s, w, h, 0, 4.5f, -11, 50.0f);
    const float at[3] = {0, 3.5f, 0};
    bpt_aim_camera_at(s, at);
    bpt_camera cam; bpt_get_camera(s, &cam); cam.focus_distance = 1.0f; bpt_set_camera(s, &cam);
    use_advanced_integrator(s);
    const float sky[3] = {0.45f, 0.6f, 0.9f};
    bpt_set_sky(s, sky, sky);
    const float g0[3] = {0.8f, 0.8f, 0.8f}, g1[3] = {0.25f, 0.25f, 0.25f}, clay[3] = {0.85f, 0.55f, 0.35f}, grey[3] = {0.1f, 0.1f, 0.1f}, e[3] = {4000, 3800, 3500};
    uint32_t ground = bpt_add_diffuse_material(s, g0, 1.0f, 0.0f, 1, g1);
    uint32_t mat = bpt_add_diffuse_material(s, clay, 1.0f, 0.0f, 0, grey);
    uint32_t light = bpt_add_emissive_material(s, e);
    const float up[3] = {0, 1, 0};
    bpt_add_plane(s, ground, up, 0.0f);
    uint32_t n = bpt_make_displaced_icosphere(level, 0.08f, nullptr);
    std::vector<float> tris((size_t)n*9);
    bpt_make_displaced_icosphere(level, 0.08f, tris.data());
    uint32_t mesh = bpt_create_mesh(s, n, tris.data(), nullptr);
    bpt_m4x4inv xf = translate(0, 3.6f, 0, 3.5f), lt = translate(9, 14, -9);
    bpt_add_mesh(s, mat, mesh, &xf);
    bpt_add_sphere(s, light, 0.5f, &lt);
}

// a mesh from an .obj file (parse_obj, assets.cpp:187-400) on the icosphere scene's stage, scaled into a 7-unit box
static int build_obj_scene(bpt_scene* s, bpt_ctx* ctx, uint32_t w, uint32_t h, const char* path, int winding, bool gpu_bvh) {
    bpt_obj* obj = bpt_load_obj(path, winding);
    if (!obj) { fprintf(stderr, "%s\n", bpt_last_error()); return 1; }
    uint32_t n = bpt_obj_triangle_count(obj);
    if (!n) { fprintf(stderr, "%s: no triangles\n", path); return 1; }
    const float* pos = bpt_obj_positions(obj);
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (size_t i = 0; i < (size_t)n*3; ++i)
        for (int k = 0; k < 3; ++k) { float v = pos[i*3 + k]; if (v < lo[k]) lo[k] = v; if (v > hi[k]) hi[k] = v; }
    float ext = hi[0] - lo[0];
    for (int k = 1; k < 3; ++k) if (hi[k] - lo[k] > ext) ext = hi[k] - lo[k];
    float scale = ext > 0 ? 7.0f/ext : 1.0f;
    set_camera(s, w, h, 0, 4.5f, -11, 50.0f);
    const float at[3] = {0, 3.5f, 0};
    bpt_aim_camera_at(s, at);
    bpt_camera cam; bpt_get_camera(s, &cam); cam.focus_distance = 1.0f; bpt_set_camera(s, &cam);
    use_advanced_integrator(s);
    const float sky[3] = {0.45f, 0.6f, 0.9f};
    bpt_set_sky(s, sky, sky);
    const float g0[3] = {0.8f, 0.8f, 0.8f}, g1[3] = {0.25f, 0.25f, 0.25f}, clay[3] = {0.85f, 0.55f, 0.35f}, grey[3] = {0.1f, 0.1f, 0.1f}, e[3] = {4000, 3800, 3500};
    uint32_t ground = bpt_add_diffuse_material(s, g0, 1.0f, 0.0f, 1, g1);
    uint32_t mat = bpt_add_diffuse_material(s, clay, 1.0f, 0.0f, 0, grey);
    uint32_t light = bpt_add_emissive_material(s, e);
    const float up[3] = {0, 1, 0};
    bpt_add_plane(s, ground, up, 0.0f);
    uint32_t mesh;
    if (gpu_bvh) {      // create_bvh_for_mesh(BVH_MidpointSplit) like load_mesh (raytracer.cpp:154), on the device: same arrays as the host build
        std::vector<bpt_bvh_node> nodes((size_t)2*n + 2);
        std::vector<uint32_t> order(n);
        uint32_t node_count = 0; float ms = 0;
        if (bpt_build_mesh_bvh_device(ctx, n, pos, BPT_BVH_MIDPOINT_SPLIT, nodes.data(), (uint32_t)nodes.size(), &node_count, order.data(), &ms) != BPT_OK) { fprintf(stderr, "%s\n", bpt_last_error()); return 1; }
        printf("device BVH build: %u triangles -> %u nodes in %.2f ms\n", n, node_count, ms);
        mesh = bpt_create_mesh_with_bvh(s, n, pos, bpt_obj_normals(obj), nodes.data(), node_count, order.data());
    } else {
        mesh = bpt_create_mesh_from_obj(s, obj);
    }
    bpt_m4x4inv xf = translate(-0.5f*(lo[0] + hi[0])*scale, 0.05f - lo[1]*scale, -0.5f*(lo[2] + hi[2])*scale, scale), lt = translate(9, 14, -9);
    bpt_add_mesh(s, mat, mesh, &xf);
    bpt_add_sphere(s, light, 0.5f, &lt);
    bpt_obj_free(obj);
    return 0;
}

#define CHECK(call) do { int rc_ = (call); if (rc_ != BPT_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, bpt_last_error()); return 1; } } while (0)

static const int kBlockRows = 8;      // interleave granularity of the row partition (block b -> GPU b % N)

// one GPU's share of a multi-GPU render: its own context, the replicated scene, its row blocks, one reduce per pass
static int render_rank(int rank, int n_gpus, int device, void* comm, const bpt_scene* scene, const uint8_t* blob,
                       uint32_t w, uint32_t h, uint32_t spp, uint32_t passes, bpt_ctx** out_ctx, bpt_stats* out_stats) {
    bpt_ctx* ctx = nullptr;
    CHECK(bpt_create(device, &ctx));
    *out_ctx = ctx;
    CHECK(bpt_set_sampler_tables(ctx, blob, blob + 16384, blob + 81920, blob + 212992));
    CHECK(bpt_upload_scene(ctx, scene));
    CHECK(bpt_film_resize(ctx, w, h));
    std::vector<int32_t> bands;
    for (uint32_t b = 0, y0 = 0; y0 < h; ++b, y0 += kBlockRows)
        if ((int)(b % (uint32_t)n_gpus) == rank) { bands.push_back((int32_t)y0); bands.push_back((int32_t)(y0 + kBlockRows < h ? y0 + kBlockRows : h)); }
    for (uint32_t p = 0; p < passes; ++p) {
        if (!bands.empty()) CHECK(bpt_render_pass_bands(ctx, 0, (int32_t)w, (uint32_t)bands.size()/2, bands.data(), p*spp, spp, BPT_SEED_PER_PIXEL, p));
        CHECK(bpt_reduce_film(ctx, comm, 0));          // the partial films keep accumulating; the root's sum is refreshed
    }
    CHECK(bpt_sync(ctx));
    CHECK(bpt_get_stats(ctx, out_stats, 0));
    return 0;
}

int main(int argc, char** argv) {
    const char* scene_name = "week3"; const char* out = "render.bmp"; const char* tables = nullptr;
    const char* obj_path = nullptr; const char* hdr_path = nullptr; const char* integrator = nullptr; const char* filter = nullptr;
    int winding = BPT_WINDING_COUNTER_CLOCKWISE; bool gpu_bvh = false;
    uint32_t w = 640, h = 360, spp = 16, passes = 1, level = 6; int device = 0; int gpus = 1;
    for (int i = 1; i < argc; ++i) {
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (!strcmp(argv[i], "--scene")) scene_name = next();
        else if (!strcmp(argv[i], "--w")) w = (uint32_t)atoi(next());
        else if (!strcmp(argv[i], "--h")) h = (uint32_t)atoi(next());
        else if (!strcmp(argv[i], "--spp")) spp = (uint32_t)atoi(next());
        else if (!strcmp(argv[i], "--passes")) passes = (uint32_t)atoi(next());
        else if (!strcmp(argv[i], "--level")) level = (uint32_t)atoi(next());
        else if (!strcmp(argv[i], "--device")) device = atoi(next());
        else if (!strcmp(argv[i], "--gpus")) gpus = atoi(next());
        else if (!strcmp(argv[i], "--out")) out = next();
        else if (!strcmp(argv[i], "--tables")) tables = next();
        else if (!strcmp(argv[i], "--obj")) { obj_path = next(); scene_name = "obj"; }
        else if (!strcmp(argv[i], "--cw")) winding = BPT_WINDING_CLOCKWISE;
        else if (!strcmp(argv[i], "--hdr")) hdr_path = next();
        else if (!strcmp(argv[i], "--integrator")) integrator = next();
        else if (!strcmp(argv[i], "--filter")) filter = next();
        else if (!strcmp(argv[i], "--gpu-bvh")) gpu_bvh = true;
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    if (!tables) { fprintf(stderr, "--tables <sampler_tables.bin> is required (the sampler lookup tables, see INTEGRATION.md)\n"); return 2; }
    if (gpus < 1 || gpus > 64) { fprintf(stderr, "--gpus must be 1..64\n"); return 2; }
    std::vector<uint8_t> blob(16384 + 65536 + 131072 + 131072);
    FILE* tf = fopen(tables, "rb");
    if (!tf || fread(blob.data(), 1, blob.size(), tf) != blob.size()) { fprintf(stderr, "cannot read %s\n", tables); return 2; }
    fclose(tf);

    bpt_ctx* ctx = nullptr;
    CHECK(bpt_create(device, &ctx));                       // fails when there is no GPU: no CPU fallback
    bpt_scene* scene = bpt_scene_create();
    auto t0 = std::chrono::steady_clock::now();
    if (!strcmp(scene_name, "week3")) build_week3(scene, w, h);
    else if (!strcmp(scene_name, "icosphere")) build_icosphere(scene, w, h, level);
    else if (!strcmp(scene_name, "obj") && obj_path) { if (build_obj_scene(scene, ctx, w, h, obj_path, winding, gpu_bvh)) return 1; }
    else { fprintf(stderr, "unknown scene %s\n", scene_name); return 2; }
    if (hdr_path) CHECK(bpt_load_skydome_hdr(scene, hdr_path));
    if (integrator) {
        int32_t id = bpt_find_integrator(integrator);     // unknown names select the default, like find_integrator (integrators.cpp:832-838)
        bpt_settings st; bpt_get_settings(scene, &st); st.integrator = id; bpt_set_settings(scene, &st);
    }
    if (filter) bpt_load_reconstruction_kernel(scene, filter);   // unknown names select Box (reconstruction_filters.cpp:112)
    CHECK(bpt_create_scene_bvh(scene));
    printf("Scene + BVH construction took: %fs\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());

    bpt_post_settings post = {0.0f, 1, 1, 0.5f, 0.0f};      // init_scene defaults (raytracer.cpp:1450-1452)
    std::vector<uint32_t> pixels((size_t)w*h);
    if (gpus > 1) {
        // one host thread + one context per GPU (devices device .. device+gpus-1), ncclCommInitAll, reduce to rank 0
        bpt_destroy(ctx); ctx = nullptr;
        std::vector<int> devs(gpus);
        for (int i = 0; i < gpus; ++i) devs[i] = device + i;
        std::vector<void*> comms(gpus, nullptr);
        CHECK(bpt_nccl_comm_init_all(gpus, devs.data(), comms.data()));
        std::vector<bpt_ctx*> ctxs(gpus, nullptr);
        std::vector<bpt_stats> stats(gpus);
        std::vector<int> rcs(gpus, 0);
        std::vector<std::thread> threads;
        t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < gpus; ++r)
            threads.emplace_back([&, r]() {
                rcs[r] = render_rank(r, gpus, devs[r], comms[r], scene, blob.data(), w, h, spp, passes, &ctxs[r], &stats[r]);
                if (rcs[r]) { fprintf(stderr, "GPU %d failed; aborting (the other ranks would wait in the collective)\n", devs[r]); _Exit(1); }
            });
        for (auto& t : threads) t.join();
        double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        uint64_t rays = 0, samples = 0;
        for (int r = 0; r < gpus; ++r) { rays += stats[r].rays; samples += stats[r].samples; }
        printf("Took %ux%u %uspp image in %f seconds on %d GPUs incl. context setup and scene upload.  (%.1f Mrays/s, %.1f Msamples/s)\n",
               w, h, spp*passes, sec, gpus, (double)rays/sec/1e6, (double)samples/sec/1e6);
        CHECK(bpt_resolve_reduced_bgra8(ctxs[0], &post, nullptr, 0, 0, pixels.data()));
        for (int r = 0; r < gpus; ++r) { bpt_nccl_comm_destroy(comms[r]); bpt_destroy(ctxs[r]); }
    } else {
        CHECK(bpt_set_sampler_tables(ctx, blob.data(), blob.data() + 16384, blob.data() + 81920, blob.data() + 212992));
        CHECK(bpt_upload_scene(ctx, scene));
        CHECK(bpt_film_resize(ctx, w, h));
        t0 = std::chrono::steady_clock::now();
        for (uint32_t p = 0; p < passes; ++p) CHECK(bpt_render_pass(ctx, 0, 0, (int32_t)w, (int32_t)h, p*spp, spp, BPT_SEED_PER_PIXEL, p));
        CHECK(bpt_sync(ctx));
        double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        bpt_stats st;
        CHECK(bpt_get_stats(ctx, &st, 0));
        printf("Took %ux%u %uspp image in %f seconds.  (%.1f Mrays/s, %.1f Msamples/s)\n", w, h, spp*passes, sec,
               (double)st.rays/sec/1e6, (double)st.samples/sec/1e6);
        CHECK(bpt_resolve_bgra8(ctx, &post, nullptr, 0, 0, pixels.data()));
        bpt_destroy(ctx);
    }
    CHECK(bpt_write_bitmap(out, pixels.data(), w, h));
    printf("wrote %s\n", out);
    bpt_scene_destroy(scene);
    return 0;
}
