// ORACLE-ONLY (test infrastructure).  Wrapper TU: compiles the reference's raytracer.cpp UNMODIFIED by
// #include (SDL_main is renamed to an unused static so the Win32/SDL/microui shell drops out), which makes its
// static functions -- init_scene, load_scene, aim_camera, render_tile, splat_filter, init_work_queue,
// render_all_tiles ... -- callable from the C entry points below.  No reference code is restated here: the
// parity render calls the reference's own render_tile once per sample on a 1x1 rect (see ref_render_parity).
#include "precomp.h"             // the shim (pulls the SDL declarations in before SDL_main is renamed)
#include <sys/mman.h>
#define SDL_main static __attribute__((unused)) ref_SDL_main_unused
#include "raytracer.cpp"         // -> /root/reference/Raytracer/raytracer.cpp via the symlink farm
#undef SDL_main

#include "oracle_internal.h"
#include <new>
#include <unistd.h>

extern "C" Uint32 SDL_GetTicks(void) { return 1u; }   // nested_dielectrics_scene seeds from it (raytracer.cpp:1372)

static void ensure_platform() {
    static bool done = false;
    if (!done) { platform_init(); done = true; }
}

static void latch(ref_scene* s) {
    // what render_all_tiles does when a render (re)starts (raytracer.cpp:711-720)
    Scene* scene = &s->scene;
    recompute_camera(&scene->new_camera);
    scene->camera   = scene->new_camera;
    scene->settings = scene->new_settings;
    s->latched = true;
}

static void install_filter(ref_scene* s) {
    FilterCache* cache = &g_filter_cache;
    cache->kernel_size = s->filter.kernel_size;
    cache->cache_size  = s->filter.cache_size;
    memcpy(cache->cache, s->filter.cache, sizeof(cache->cache));
}

static void save_filter(ref_scene* s) {
    s->filter.kernel_size = g_filter_cache.kernel_size;
    s->filter.cache_size  = g_filter_cache.cache_size;
    memcpy(s->filter.cache, g_filter_cache.cache, sizeof(s->filter.cache));
}

extern "C" {

BPT_API ref_scene* ref_scene_create(void) {
    ensure_platform();
    ref_scene* s = new ref_scene();
    memset(&s->scene, 0, sizeof(s->scene));
    memset(&s->temp, 0, sizeof(s->temp));
    memset(&s->skydome, 0, sizeof(s->skydome));
    s->latched = false;
    clear_scene(&s->scene);
    init_scene(&s->scene);        // raytracer.cpp:1424-1453 (loads Mitchell-Netravali into g_filter_cache)
    save_filter(s);
    return s;
}

BPT_API void ref_scene_destroy(ref_scene* s) {
    if (!s) return;
    s->scene.materials.free(); s->scene.lights.free(); s->scene.planes.free(); s->scene.primitives.free();
    // arenas are 128 GiB PROT_NONE reservations; unmap them
    if (s->scene.arena.base) munmap(s->scene.arena.base, s->scene.arena.capacity);
    if (s->temp.base) munmap(s->temp.base, s->temp.capacity);
    delete s;
}

BPT_API uint32_t ref_add_material(ref_scene* s, const bpt_material* m) {
    Material mat; memcpy(&mat, m, sizeof(mat));
    return add_material(&s->scene, mat);
}
BPT_API uint32_t ref_add_diffuse_material(ref_scene* s, const float c[3], float ior, float roughness, int32_t checkers, const float cc[3]) {
    return add_diffuse_material(&s->scene, to_v3(c), ior, roughness, checkers, to_v3(cc));
}
BPT_API uint32_t ref_add_translucent_material(ref_scene* s, const float absorb[3], float ior, float roughness) {
    return add_translucent_material(&s->scene, to_v3(absorb), ior, roughness);
}
BPT_API uint32_t ref_add_emissive_material(ref_scene* s, const float e[3]) {
    return add_emissive_material(&s->scene, to_v3(e));
}
BPT_API uint32_t ref_add_plane(ref_scene* s, uint32_t mat, const float n[3], float d) {
    return add_plane(&s->scene, MaterialID::from(mat), to_v3(n), d);
}
BPT_API uint32_t ref_add_sphere(ref_scene* s, uint32_t mat, float r, const bpt_m4x4inv* xf) {
    if (xf) return add_sphere(&s->scene, MaterialID::from(mat), r, to_ref(xf));
    return add_sphere(&s->scene, MaterialID::from(mat), r, (M4x4Inv*)nullptr);
}
BPT_API uint32_t ref_add_box(ref_scene* s, uint32_t mat, const float r[3], const bpt_m4x4inv* xf) {
    if (xf) return add_box(&s->scene, MaterialID::from(mat), to_v3(r), to_ref(xf));
    return add_box(&s->scene, MaterialID::from(mat), to_v3(r), (M4x4Inv*)nullptr);
}

BPT_API uint32_t ref_create_mesh(ref_scene* s, uint32_t triangle_count, const float* positions, const float* normals) {
    return ref_create_mesh_ex(s, triangle_count, positions, normals, (int32_t)BVH_SAHBinned);
}

BPT_API uint32_t ref_create_mesh_ex(ref_scene* s, uint32_t triangle_count, const float* positions, const float* normals, int32_t method) {
    Scene* scene = &s->scene;
    Mesh mesh = {};
    mesh.has_normals = normals ? 1 : 0;
    mesh.triangle_count = triangle_count;
    usize blocks = normals ? 2 : 1;
    mesh.triangles = push_array(&scene->arena, blocks*(usize)triangle_count, Triangle, no_clear());
    memcpy(mesh.triangles, positions, sizeof(Triangle)*(usize)triangle_count);
    if (normals) memcpy(mesh.triangles + triangle_count, normals, sizeof(Triangle)*(usize)triangle_count);
    // same call create_scene_bvh would make lazily (scene.cpp:208); done up front so instances share it,
    // like load_mesh does for the dragon (raytracer.cpp:150-159)
    mesh.bvh = create_bvh_for_mesh(mesh.triangle_count, mesh.triangles, &scene->arena, &s->temp, (BVHConstructionMethod)method, BVHStorage_Scalar);
    s->meshes.push_back(mesh);
    return (uint32_t)(s->meshes.size() - 1);
}

BPT_API uint32_t ref_add_mesh(ref_scene* s, uint32_t mat, uint32_t mesh, const bpt_m4x4inv* xf) {
    if (mesh >= s->meshes.size()) return 0xFFFFFFFFu;
    if (xf) return add_mesh(&s->scene, MaterialID::from(mat), &s->meshes[mesh], to_ref(xf));
    return add_mesh(&s->scene, MaterialID::from(mat), &s->meshes[mesh], (M4x4Inv*)nullptr);
}

BPT_API int ref_set_sky(ref_scene* s, const float top[3], const float bot[3]) {
    s->scene.top_sky_color = to_v3(top);
    s->scene.bot_sky_color = to_v3(bot);
    return 0;
}

BPT_API int ref_set_ambient_light(ref_scene* s, const float rgb[3]) {
    s->scene.ambient_light = to_v3(rgb);
    return 0;
}

BPT_API int ref_set_skydome(ref_scene* s, uint32_t w, uint32_t h, const float* pixels) {
    if (!pixels) { s->scene.skydome = nullptr; return 0; }
    s->skydome.w = w; s->skydome.h = h;
    s->skydome.pixels = push_array(&s->scene.arena, (usize)w*h, V3, no_clear());
    memcpy(s->skydome.pixels, pixels, sizeof(V3)*(usize)w*h);
    s->scene.skydome = &s->skydome;
    return 0;
}

BPT_API int ref_get_camera(const ref_scene* s, bpt_camera* out) { memcpy(out, &s->scene.new_camera, sizeof(*out)); return 0; }
BPT_API int ref_set_camera(ref_scene* s, const bpt_camera* c) { memcpy(&s->scene.new_camera, c, sizeof(*c)); s->latched = false; return 0; }
BPT_API int ref_aim_camera(ref_scene* s, const float d[3]) { aim_camera(&s->scene.new_camera, to_v3(d)); s->latched = false; return 0; }
BPT_API int ref_aim_camera_at(ref_scene* s, const float at[3]) { aim_camera_at(&s->scene.new_camera, to_v3(at)); s->latched = false; return 0; }

BPT_API int ref_find_integrator(const char* name) { return (int)(find_integrator(name) - g_integrators); }

BPT_API int ref_get_settings(const ref_scene* s, bpt_settings* out) {
    memcpy(out, &s->scene.new_settings, offsetof(bpt_settings, integrator));
    out->integrator = (int)(s->scene.new_settings.integrator - g_integrators);
    return 0;
}
BPT_API int ref_set_settings(ref_scene* s, const bpt_settings* in) {
    memcpy(&s->scene.new_settings, in, offsetof(bpt_settings, integrator));
    int idx = in->integrator;
    if (idx < 0 || idx >= (int)g_integrator_count) idx = 0;
    s->scene.new_settings.integrator = &g_integrators[idx];
    s->latched = false;
    return 0;
}

BPT_API int ref_load_reconstruction_kernel(ref_scene* s, const char* filter_name) {
    FilterKernelOption* f = find_filter(filter_name);
    load_reconstruction_kernel(f);
    save_filter(s);
    return (int)(f - g_filters);
}
BPT_API int ref_get_filter_cache(const ref_scene* s, bpt_filter_cache* out) { *out = s->filter; return 0; }
BPT_API int ref_set_filter_cache(ref_scene* s, const bpt_filter_cache* in) { s->filter = *in; return 0; }

BPT_API int ref_create_scene_bvh(ref_scene* s) {
    create_scene_bvh(&s->scene, &s->temp);
    // meshes that create_scene_bvh built lazily live in the primitives' copies; ours were built up front
    return 0;
}

BPT_API int ref_get_scene_bvh(const ref_scene* s, const bpt_bvh_node** nodes, uint32_t* node_count, const uint32_t** indices, uint32_t* index_count) {
    BVH* bvh = s->scene.bvh;
    if (!bvh) return -2;
    *nodes = (const bpt_bvh_node*)bvh->nodes; *node_count = bvh->node_count;
    *indices = bvh->indices; *index_count = bvh->index_count;
    return 0;
}

BPT_API int ref_get_mesh_bvh(const ref_scene* s, uint32_t mesh, const bpt_bvh_node** nodes, uint32_t* node_count, const uint32_t** indices, uint32_t* index_count, const float** tris) {
    if (mesh >= s->meshes.size()) return -1;
    MeshBVH* bvh = s->meshes[mesh].bvh;
    if (!bvh) return -2;
    *nodes = (const bpt_bvh_node*)bvh->nodes; *node_count = bvh->node_count;
    *indices = bvh->indices; *index_count = bvh->index_count;
    if (tris) *tris = (const float*)bvh->triangles;
    return 0;
}

BPT_API int ref_get_counts(const ref_scene* s, uint32_t* materials, uint32_t* primitives, uint32_t* planes, uint32_t* lights, uint32_t* meshes) {
    if (materials)  *materials  = s->scene.materials.count;
    if (primitives) *primitives = s->scene.primitives.count;
    if (planes)     *planes     = s->scene.planes.count;
    if (lights)     *lights     = s->scene.lights.count;
    if (meshes)     *meshes     = (uint32_t)s->meshes.size();
    return 0;
}

BPT_API int ref_load_builtin_scene(ref_scene* s, const char* name, uint32_t w, uint32_t h) {
    for (usize i = 0; i < ArrayCount(g_scenes); ++i) {
        if (0 == strcmp(g_scenes[i].name, name)) {
            load_scene(&s->scene, &g_scenes[i], w, h, &s->temp);   // raytracer.cpp:1455-1470
            save_filter(s);
            s->latched = false;
            return 0;
        }
    }
    return -1;
}

BPT_API int ref_sizeof(const char* n) {
    if (!strcmp(n, "Material"))  return (int)sizeof(Material);
    if (!strcmp(n, "Camera"))    return (int)sizeof(Camera);
    if (!strcmp(n, "SceneSettings")) return (int)sizeof(SceneSettings);
    if (!strcmp(n, "BVHNode"))   return (int)sizeof(BVHNode);
    if (!strcmp(n, "Triangle"))  return (int)sizeof(Triangle);
    if (!strcmp(n, "Primitive")) return (int)sizeof(Primitive);
    if (!strcmp(n, "M4x4Inv"))   return (int)sizeof(M4x4Inv);
    if (!strcmp(n, "Ray"))       return (int)sizeof(Ray);
    if (!strcmp(n, "V4"))        return (int)sizeof(V4);
    if (!strcmp(n, "RandomSeries")) return (int)sizeof(RandomSeries);
    if (!strcmp(n, "FilterCache")) return (int)sizeof(FilterCache);
    return -1;
}

BPT_API float ref_kat_filter(const char* name, float x) {
    FilterKernelOption* f = find_filter(name);
    return f->f ? f->f(x) : 0.0f;
}

// ---- parity render -------------------------------------------------------------------------------------------

static Integrator* g_real_integrator;
static bpt_sample_record* g_record_slot;

static INTEGRATOR(recording_integrator) {
    u64 rays0 = g_ref_rays;
    V3 r = g_real_integrator(state);
    if (g_record_slot) {
        bpt_sample_record* rec = g_record_slot;
        rec->ray_o[0] = state->in_ray_o.x; rec->ray_o[1] = state->in_ray_o.y; rec->ray_o[2] = state->in_ray_o.z;
        rec->ray_d[0] = state->in_ray_d.x; rec->ray_d[1] = state->in_ray_d.y; rec->ray_d[2] = state->in_ray_d.z;
        rec->radiance[0] = r.x; rec->radiance[1] = r.y; rec->radiance[2] = r.z;
        rec->rays = (uint32_t)(g_ref_rays - rays0);
    }
    return r;
}

BPT_API int ref_render_parity(ref_scene* s, float* film, uint32_t w, uint32_t h,
                              int32_t x0, int32_t y0, int32_t x1, int32_t y1,
                              uint32_t frame_count, uint32_t spp, uint32_t salt, bpt_sample_record* records) {
    Scene* scene = &s->scene;
    if (!scene->bvh) return -2;
    if (!s->latched) latch(s);
    install_filter(s);

    AccumulationBuffer buffer = {};
    buffer.w = w; buffer.h = h; buffer.pixels = (V4*)film;

    WorkQueue queue;
    memset(&queue, 0, sizeof(queue));
    queue.parameters.scene = scene;
    queue.parameters.backbuffer = &buffer;
    queue.parameters.frontbuffer = &buffer;
    queue.discard_render = false;

    static char* task_memory = (char*)malloc(THREAD_TASK_ARENA_SIZE + 64);
    Arena task_arena;
    task_arena.init_with_memory(THREAD_TASK_ARENA_SIZE, task_memory);

    u32 saved_spp = scene->settings.samples_per_pixel;
    IntegratorOption* saved_integrator = scene->settings.integrator;
    static IntegratorOption recording_option = { "Recording", recording_integrator };
    g_real_integrator = saved_integrator->f;
    scene->settings.integrator = &recording_option;
    scene->settings.samples_per_pixel = 1;
    int saved_count = g_ref_count_rays;
    g_ref_count_rays = 1;

    bpt_sample_record* rec = records;
    for (s32 y = y0; y < y1; ++y)
    for (s32 x = x0; x < x1; ++x)
    for (u32 si = 0; si < spp; ++si) {
        u32 sample_index = frame_count + si;
        buffer.frame_count = sample_index;
        RandomSeries entropy = random_seed(hash_coordinate((u32)x, (u32)y, sample_index) ^ salt);
        g_record_slot = rec;
        task_arena.clear();
        render_tile(&queue, entropy, x, y, x + 1, y + 1, &task_arena);   // the reference's own per-sample body
        if (rec) ++rec;
    }

    g_record_slot = nullptr;
    g_ref_count_rays = saved_count;
    scene->settings.integrator = saved_integrator;
    scene->settings.samples_per_pixel = saved_spp;
    return 0;
}

// ---- verbatim tile-multithreaded render (CPU baseline) ---------------------------------------------------------

BPT_API int ref_render_threaded(ref_scene* s, uint32_t w, uint32_t h, uint32_t spp, uint32_t threads,
                                float* film, double* seconds, bpt_stats* stats) {
    Scene* scene = &s->scene;
    if (!scene->bvh) return -2;
    install_filter(s);
    scene->new_settings.samples_per_pixel = spp;

    // Everything the queue and its (never-exiting, raytracer.cpp:610-627) workers touch is leaked on purpose.
    Arena* permanent = new Arena(); memset(permanent, 0, sizeof(*permanent));
    Arena* transient = new Arena(); memset(transient, 0, sizeof(*transient));
    (void)push_size(permanent, 64);          // Arena capacity is set lazily on first push (memory_arena.cpp:6-9)

    RenderParameters parameters = {};
    parameters.scene       = scene;
    parameters.backbuffer  = allocate_accumulation_buffer(transient, w, h);
    parameters.frontbuffer = allocate_accumulation_buffer(transient, w, h);
    parameters.path_guide  = allocate_path_guide(transient, w, h);

    WorkQueue* queue = (WorkQueue*)aligned_alloc(64, (sizeof(WorkQueue) + 63) & ~(size_t)63);
    init_work_queue(queue, threads, permanent, parameters, 64, 64);   // raytracer.cpp:1656-1661
    WRITE_BARRIER;

    int saved_count = g_ref_count_rays;
    g_ref_count_rays = 0;                    // no shared counter traffic inside the timed region
    zero_struct(&g_stats);

    // first call: discard_render is true -> reset buffers, latch camera/settings, kick the workers (:705-749)
    f64 elapsed = 0.0;
    TraversalStats ts = {};
    b32 done = render_all_tiles(queue, &elapsed, &ts);
    while (!done) {
        usleep(200);
        done = render_all_tiles(queue, &elapsed, &ts);   // true once every tile of the pass has retired
    }
    // the finished pass is in frontbuffer (copy + Swap, :708-709); a new pass was kicked -> throw it away
    if (film) memcpy(film, queue->parameters.frontbuffer->pixels, sizeof(V4)*(usize)w*h);
    discard_current_render(queue);
    while (queue->tiles_retired < (s32)queue->total_tile_count) usleep(200);

    if (seconds) *seconds = elapsed;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->mesh_intersection_count = ts.mesh_intersection_count;
        stats->mesh_bvh_traversals     = ts.mesh_bvh_traversals;
        stats->mesh_node_traversals    = ts.mesh_node_traversals;
        stats->mesh_leaf_traversals    = ts.mesh_leaf_traversals;
        stats->samples = (uint64_t)w*h*spp;
    }
    g_ref_count_rays = saved_count;
    s->latched = true;   // render_all_tiles latched camera/settings itself
    return 0;
}

// The reference's display-loop resolve.  The loop below is the reference's OWN text, cut out of raytracer.cpp:2103-2173
// at build time (oracle/tools/slice_resolve.py -> oracle/_ref/resolve_slice.inc, generated and git-ignored); the locals
// here only give it the names it uses inside SDL_main.
BPT_API int ref_resolve_bgra8(const float* film_rgba, uint32_t w_, uint32_t h_, const bpt_post_settings* post,
                              const uint8_t* dither_rgb8, uint32_t dither_w, uint32_t dither_h, uint32_t* out_pixels) {
    if (!film_rgba || !post || !dither_rgb8 || !out_pixels || dither_w == 0 || dither_h == 0) return -1;
    AccumulationBuffer buffer = {};
    buffer.w = w_; buffer.h = h_; buffer.pixels = (V4*)film_rgba;
    AccumulationBuffer* hdr_buffer = &buffer;
    PostProcessSettings settings_copy;
    settings_copy.exposure = post->exposure; settings_copy.tonemapping = post->tonemapping;
    settings_copy.srgb_transform = post->srgb_transform; settings_copy.midpoint = post->midpoint; settings_copy.contrast = post->contrast;
    PostProcessSettings* post_settings = &settings_copy;
    Image_R8G8B8 noise = {};
    noise.w = dither_w; noise.h = dither_h; noise.pixels = (Color_R8G8B8*)dither_rgb8;
    Image_R8G8B8* dither_noise = &noise;
    void* pixels = out_pixels;
    int w = (int)w_, h = (int)h_;
    {
#include "resolve_slice.inc"
    }
    return 0;
}

// write_bitmap (assets.cpp:693-724), the reference's own
BPT_API int ref_write_bitmap(const char* file_name, const uint32_t* pixels, uint32_t w, uint32_t h) {
    ensure_platform();
    static Arena arena = {};
    write_bitmap((u32*)pixels, w, h, file_name, &arena);
    return 0;
}

} // extern "C"

#ifdef ORACLE_WITH_INTEGRATION
// ---- INTEGRATION.md's binding, compiled as written against the reference's own Scene (oracle/_ref/libbpt_integration.so
//      only: the one artefact in which reference code and the product library meet; nothing else loads it) ----------------
#include "integration/raytracer_bpt.inl"

extern "C" int ref_get_sampler_tables(uint8_t* perm, uint8_t* sobol, uint8_t* scramble, uint8_t* rank);   // tu_samplers.cpp

// load_scene(g_scenes[name]) exactly like SDL_main (raytracer.cpp:1455-1470, :1630) -> bpt_mirror_scene -> `passes` calls
// of render_all_tiles_bpt -> the front buffer.  `scene_out` (nullable) receives the ref_scene so the caller can render the
// very same Scene with the reference's own CPU path.
extern "C" BPT_API int integration_render_builtin(ref_scene* s, const char* name, uint32_t w, uint32_t h, uint32_t spp,
                                                  uint32_t passes, int device, float* film_out) {
    if (ref_load_builtin_scene(s, name, w, h) != 0) return -1;
    Scene* scene = &s->scene;
    scene->new_settings.samples_per_pixel = spp;
    static uint8_t perm[16384], sobol[65536], scr[131072], rank[131072];
    if (ref_get_sampler_tables(perm, sobol, scr, rank) != 0) return -2;

    BptBridge bridge = {};
    if (bpt_create(device, &bridge.ctx) != BPT_OK) { fprintf(stderr, "bpt: %s\n", bpt_last_error()); return -3; }
    int rc = 0;
    AccumulationBuffer back = {}, front = {};
    std::vector<V4> back_px((size_t)w*h), front_px((size_t)w*h);
    back.w = front.w = w; back.h = front.h = h; back.pixels = back_px.data(); front.pixels = front_px.data();
    RenderParameters params = {};
    params.scene = scene; params.backbuffer = &back; params.frontbuffer = &front;
    if (bpt_set_sampler_tables(bridge.ctx, perm, sobol, scr, rank) != BPT_OK) rc = -4;
    if (!rc) {
        bpt_mirror_scene(&bridge, scene);
        scene->camera = scene->new_camera; scene->settings = scene->new_settings;      // the first latch of render_all_tiles (:711-720)
        recompute_camera(&scene->camera);
        bpt_latch(&bridge, scene);
        if (bpt_upload_scene(bridge.ctx, bridge.mirror) != BPT_OK || bpt_film_resize(bridge.ctx, w, h) != BPT_OK) rc = -5;
    }
    for (uint32_t p = 0; !rc && p < passes; ++p)
        if (!render_all_tiles_bpt(&bridge, &params)) rc = -6;
    if (!rc) memcpy(film_out, front.pixels, sizeof(V4)*(size_t)w*h);
    if (rc) fprintf(stderr, "integration: rc %d: %s\n", rc, bpt_last_error());
    if (bridge.mirror) bpt_scene_destroy(bridge.mirror);
    bpt_destroy(bridge.ctx);
    s->latched = false;
    return rc;
}
#endif

// ---- the reference's asset parsers (assets.cpp, compiled unmodified into assets.o) ---------------------------------
#include <sys/mman.h>
static void release_arena(Arena* a) { if (a->base) munmap(a->base, a->capacity); delete a; }
BPT_API int ref_parse_obj(const char* text, int winding, uint32_t* triangle_count, int* has_normals, int* has_texcoords,
                          float* positions, float* normals, float* texcoords) {
    ensure_platform();
    Arena* arena = new Arena(); memset(arena, 0, sizeof(*arena));
    Arena* temp = new Arena(); memset(temp, 0, sizeof(*temp));
    (void)push_size(arena, 64); (void)push_size(temp, 64);     // Arena capacity is set lazily on first push (memory_arena.cpp:6-9)
    size_t len = strlen(text);
    char* input = (char*)malloc(len + 1);
    memcpy(input, text, len + 1);
    Mesh mesh;
    b32 ok = parse_obj(arena, temp, input, &mesh, winding == 0 ? MeshWinding_Clockwise : MeshWinding_CounterClockwise);
    free(input);
    if (!ok) { release_arena(arena); release_arena(temp); return -1; }
    *triangle_count = mesh.triangle_count;
    *has_normals = mesh.has_normals ? 1 : 0;
    *has_texcoords = mesh.has_texture_coordinates ? 1 : 0;
    if (positions) memcpy(positions, mesh.triangles, (size_t)mesh.triangle_count*sizeof(Triangle));
    if (normals && mesh.has_normals) memcpy(normals, get_normals(&mesh), (size_t)mesh.triangle_count*sizeof(Triangle));
    if (texcoords && mesh.has_texture_coordinates) memcpy(texcoords, get_texture_coordinates(&mesh), (size_t)mesh.triangle_count*sizeof(Triangle));
    release_arena(arena); release_arena(temp);
    return 0;
}

BPT_API int ref_parse_hdr(const char* data, size_t size, uint32_t* w, uint32_t* h, float* pixels) {
    ensure_platform();
    Arena* arena = new Arena(); memset(arena, 0, sizeof(*arena));
    Arena* temp = new Arena(); memset(temp, 0, sizeof(*temp));
    (void)push_size(arena, 64); (void)push_size(temp, 64);
    char* input = (char*)malloc(size + 1);
    memcpy(input, data, size);
    input[size] = 0;
    Image_V3 image;
    b32 ok = parse_hdr(arena, temp, input, &image);
    free(input);
    if (!ok) { release_arena(arena); release_arena(temp); return -1; }
    *w = image.w; *h = image.h;
    if (pixels) memcpy(pixels, image.pixels, (size_t)image.w*image.h*sizeof(V3));
    release_arena(arena); release_arena(temp);
    return 0;
}
