// raytracer_bpt.inl -- the binding a maintainer of the reference adds to raytracer.cpp (which quotes this
// file): it replaces the render_all_tiles / render_tile call chain (Raytracer/raytracer.cpp:366-495, :692-757) with calls into
// the C ABI of include/bpt.h.  Include it after scene.h / Raytracer.h, i.e. anywhere below raytracer.cpp:20.
//
// This file is compiled as written, against the reference's own Scene, by oracle/Makefile (`make integration`), and run
// by tests/test_integration_binding.py: a g_scenes[] entry loaded by the reference's load_scene is mirrored, rendered on
// the GPU through render_all_tiles_bpt and compared with the reference's own CPU render of the same Scene.
extern "C" {
#include "bpt.h"
}

struct BptBridge { bpt_ctx* ctx; bpt_scene* mirror; };

// Mirror the host Scene into the library's host scene model.  Primitive order (hence PrimitiveIDs, hence the TLAS)
// is preserved, and bpt_create_scene_bvh() reproduces create_bvh*/create_scene_bvh bit for bit, so hit ids agree.
static void bpt_mirror_scene(BptBridge* b, Scene* scene) {
    if (b->mirror) bpt_scene_destroy(b->mirror);
    b->mirror = bpt_scene_create();                 // already holds the null material / null primitive (init_scene)
    for (u32 i = 1; i < scene->materials.count; ++i)
        bpt_add_material(b->mirror, (const bpt_material*)&scene->materials[i]);        // layout-compatible
    for (u32 i = 0; i < scene->planes.count; ++i) {
        Primitive* p = &scene->planes[i];
        bpt_add_plane(b->mirror, p->material_id, p->plane.n.e, p->plane.d);
    }
    // meshes shared between instances alias `triangles`/`bvh` (scene.cpp:146-154): dedupe on the triangle pointer
    std::vector<std::pair<Triangle*, u32>> seen;
    for (u32 i = 1; i < scene->primitives.count; ++i) {
        Primitive* p = &scene->primitives[i];
        const bpt_m4x4inv* xf = (const bpt_m4x4inv*)p->transform;                      // M4x4Inv is 2x float[4][4]
        switch (p->type) {
            case Primitive_Sphere: bpt_add_sphere(b->mirror, p->material_id, p->sphere.r, xf); break;
            case Primitive_Box:    bpt_add_box(b->mirror, p->material_id, p->box.r.e, xf); break;
            case Primitive_Mesh: {
                u32 mesh = ~0u;
                for (auto& s : seen) if (s.first == p->mesh.triangles) mesh = s.second;
                if (mesh == ~0u) {
                    mesh = bpt_create_mesh(b->mirror, p->mesh.triangle_count, (const float*)p->mesh.triangles,
                                           p->mesh.has_normals ? (const float*)get_normals(&p->mesh) : nullptr);
                    seen.push_back({p->mesh.triangles, mesh});
                }
                bpt_add_mesh(b->mirror, p->material_id, mesh, xf);
            } break;
            default: break;                                                            // CSG nodes are never intersected
        }
    }
    bpt_set_sky(b->mirror, scene->top_sky_color.e, scene->bot_sky_color.e);
    bpt_set_ambient_light(b->mirror, scene->ambient_light.e);                          // read by "Whitted" only
    if (scene->skydome) bpt_set_skydome(b->mirror, scene->skydome->w, scene->skydome->h, (const float*)scene->skydome->pixels);
    bpt_create_scene_bvh(b->mirror);
}

static void bpt_latch(BptBridge* b, Scene* scene) {            // what render_all_tiles does at :711-720
    bpt_set_camera(b->mirror, (const bpt_camera*)&scene->new_camera);                  // layout-compatible
    bpt_settings st;
    memcpy(&st, &scene->new_settings, offsetof(bpt_settings, integrator));             // common prefix
    st.integrator = (int32_t)(scene->new_settings.integrator - g_integrators);
    bpt_set_settings(b->mirror, &st);
    bpt_filter_cache fc = { g_filter_cache.kernel_size, g_filter_cache.cache_size, {0} };
    memcpy(fc.cache, g_filter_cache.cache, sizeof(fc.cache));
    bpt_set_filter_cache(b->mirror, &fc);
    bpt_update_settings(b->ctx, b->mirror);
}

// once, after load_scene():
//   bpt_create(0, &bridge.ctx);
//   u8 sobol[65536], scr[131072], rank[131072];                    // the vendored tables are `int`, values 0..255
//   for (...) sobol[i] = (u8)sobol_256spp_256d[i]; ...             // (needs a 3-line accessor in samplers.cpp)
//   bpt_set_sampler_tables(bridge.ctx, &g_strata_permutation_sets[0][0], sobol, scr, rank);
//   bpt_mirror_scene(&bridge, &scene);  bpt_upload_scene(bridge.ctx, bridge.mirror);
//   bpt_film_resize(bridge.ctx, w, h);

// replaces the body of render_all_tiles + every render_tile of the pass:
static b32 render_all_tiles_bpt(BptBridge* b, RenderParameters* params) {
    Scene* scene = params->scene;
    AccumulationBuffer* back = params->backbuffer;
    if (!structs_are_equal(&scene->settings, &scene->new_settings) ||
        !structs_are_equal(&scene->camera, &scene->new_camera)) {
        scene->camera = scene->new_camera; scene->settings = scene->new_settings;
        bpt_latch(b, scene); bpt_film_clear(b->ctx); back->frame_count = 0;
    }
    u32 spp = scene->settings.samples_per_pixel;
    if (bpt_render_pass(b->ctx, 0, 0, back->w, back->h, back->frame_count, spp, BPT_SEED_PER_PIXEL,
                        scene->total_frame_index) != BPT_OK) {
        fprintf(stderr, "bpt: %s\n", bpt_last_error());
        return false;
    }
    bpt_download_film(b->ctx, (float*)params->frontbuffer->pixels);   // V4 == float[4]; feeds the existing tonemap/present
    params->frontbuffer->frame_count = back->frame_count += spp;
    scene->total_frame_index += 1;
    return true;
}
