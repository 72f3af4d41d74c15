/*
 * bpt.h -- C ABI of the B200-native path-tracing core for BUAS-Pathtracer.
 *
 * This is the drop-in boundary.  The reference has no FFI layer;
 * its seam is the per-pass call chain
 *     render_all_tiles (Raytracer/raytracer.cpp:692-757)
 *       -> try_render_next_tile (:551-603) -> render_tile (:366-495)
 * i.e. "(immutable Scene + latched camera/settings, filter LUT, frame_count, pixel rect, spp)
 * -> accumulate into AccumulationBuffer.pixels".  Every entry point below cites the
 * reference interface it replaces.  All types are plain C PODs; the layouts marked
 * "layout-compatible" can be filled by memcpy from the reference's own structs
 * (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Two halves:
 *   1. host scene model  (bpt_scene_*, bpt_add_*, bpt_create_scene_bvh) -- mirrors
 *      Raytracer/scene.h:134-149 and bvh.h:62-63, rebuilt from scratch; produces BVH
 *      arrays bit-identical to the reference's (parity-tested against oracle/_ref).
 *   2. device renderer   (bpt_create, bpt_upload_*, bpt_render_pass, bpt_trace, ...)
 *      -- hand-written CUDA for sm_100a.  There is NO CPU fallback: every device entry
 *      point returns BPT_ERR_CUDA when no usable GPU/driver is present.
 *
 * Error convention: functions returning int return BPT_OK (0) or a negative bpt_status;
 * bpt_last_error() gives a thread-local human-readable message.  The library never
 * aborts (the reference asserts; Raytracer/common.h:37-38).
 */
#ifndef BPT_H
#define BPT_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPT_API __attribute__((visibility("default")))

typedef enum bpt_status {
    BPT_OK            =  0,
    BPT_ERR_ARG       = -1,   /* null / out-of-range argument */
    BPT_ERR_STATE     = -2,   /* call order violated (e.g. render before upload) */
    BPT_ERR_CUDA      = -3,   /* CUDA runtime error or no device: there is no CPU fallback */
    BPT_ERR_NOMEM     = -4,
    BPT_ERR_UNSUPPORTED = -5  /* feature of the reference that is out of scope */
} bpt_status;

/* ---------------------------------------------------------------------------------
 * PODs (layout-compatible with the reference where noted)
 * ------------------------------------------------------------------------------- */

/* MathLib/math_types.h:42-49 -- row-major e[row][col]; translation in e[r][3]. */
typedef struct bpt_m4x4    { float e[4][4]; } bpt_m4x4;
typedef struct bpt_m4x4inv { bpt_m4x4 forward, inverse; } bpt_m4x4inv;   /* 128 B */

/* Raytracer/scene.h:9-29 (Material, 68 B) -- layout-compatible. */
enum { BPT_MATERIAL_MIRROR = 0x1, BPT_MATERIAL_CHECKERS = 0x2, BPT_MATERIAL_EMISSIVE = 0x4 };
typedef struct bpt_material {
    uint32_t flags;
    float albedo[3];
    float checker_color[3];
    float emission_color[3];
    float ior;
    float metallic;
    float roughness;
    int32_t is_participating_medium;
    float absorb[3];                 /* Medium::absorb */
} bpt_material;

/* Raytracer/scene.h:31-46 (Camera, 76 B) -- layout-compatible. */
typedef struct bpt_camera {
    float p[3], x[3], y[3], z[3];
    float vfov, aspect_ratio;
    float lens_radius, focus_distance;
    float film_distance, half_film_w, half_film_h;
} bpt_camera;

/* Raytracer/samplers.h:110-117 */
enum { BPT_SAMPLING_UNIFORM = 0, BPT_SAMPLING_OPTIMIZED_BLUE_NOISE = 1, BPT_SAMPLING_STRATIFIED = 2 };

/* Raytracer/integrators.cpp:823-830 (g_integrators order) */
enum { BPT_INTEGRATOR_ADVANCED = 0, BPT_INTEGRATOR_WHITTED = 1, BPT_INTEGRATOR_GT_RECURSIVE = 2,
       BPT_INTEGRATOR_GT_ITERATIVE = 3, BPT_INTEGRATOR_NORMALS = 4, BPT_INTEGRATOR_DISTANCES = 5 };

/* Raytracer/scene.h:64-82 (SceneSettings).  Field-for-field the same up to and including
 * max_bounce_count (60 B, layout-compatible prefix); the trailing IntegratorOption* of the
 * reference becomes an index into g_integrators (find_integrator, integrators.cpp:834). */
typedef struct bpt_settings {
    int32_t next_event_estimation;
    int32_t importance_sample_lights;
    int32_t importance_sample_diffuse;
    int32_t use_mis;
    int32_t russian_roulette;
    int32_t caustics;
    int32_t sampling_strategy;
    int32_t use_path_guide;           /* ignored: the reference never reads it on the path */
    float   vignette_strength;
    float   lens_distortion;
    float   f_factor;
    float   diaphragm_edges;
    float   phi_shutter_max;
    uint32_t samples_per_pixel;
    uint32_t max_bounce_count;
    int32_t integrator;               /* BPT_INTEGRATOR_* */
} bpt_settings;

/* Raytracer/bvh.h:31-37 (BVHNode, 32 B) -- layout-compatible, and the unit the
 * bit-exactness test memcmp()s. Leaf iff count != 0. */
typedef struct bpt_bvh_node {
    float    bv_p[3];
    float    bv_r[3];
    uint32_t left_first;
    uint16_t count;
    uint16_t split_axis;
} bpt_bvh_node;

/* Raytracer/primitives.h:3-10 */
enum { BPT_PRIM_NONE = 0, BPT_PRIM_PLANE = 1, BPT_PRIM_SPHERE = 2, BPT_PRIM_BOX = 3, BPT_PRIM_MESH = 4 };

/* Raytracer/Raytracer.h:34-40 (FilterCache minus the option struct). */
typedef struct bpt_filter_cache {
    uint32_t kernel_size;             /* radius; 0 = Box (no splat) */
    uint32_t cache_size;              /* 256, or 0 for Box */
    float    cache[512];
} bpt_filter_cache;

/* Raytracer/intersection.h:5-11 (Ray) as the caller supplies it: make_ray() is re-done on device. */
typedef struct bpt_ray {
    float o[3];
    float d[3];
    float max_t;
} bpt_ray;

/* What intersect_scene_internal (Raytracer/intersection.cpp:411-598) knows at return. */
typedef struct bpt_hit {
    float    t;                       /* == ray.max_t when nothing was hit */
    uint32_t primitive;               /* index into Scene::primitives; planes: 0x80000000|plane index; miss: 0xFFFFFFFF */
    uint32_t triangle;                /* original triangle index (MeshBVH::indices), 0xFFFFFFFF if not a mesh */
    float    n[3];                    /* world normal (BPT_TRACE_CLOSEST only) */
    float    p[3];                    /* world hit point (BPT_TRACE_CLOSEST only) */
} bpt_hit;
#define BPT_HIT_MISS  0xFFFFFFFFu
#define BPT_HIT_PLANE 0x80000000u

enum { BPT_TRACE_CLOSEST = 0,   /* intersect_scene,      intersection.cpp:606-610 */
       BPT_TRACE_OCCLUSION = 1  /* intersect_shadow_ray, intersection.cpp:600-604 */ };

/* Counters in the units of the REFERENCE's binary-BVH traversal: node pops include
 * failed box tests.  mesh_* match TraversalStats (Raytracer/intersection.h:33-38) exactly. */
typedef struct bpt_stats {
    uint64_t rays;                    /* intersect_scene + intersect_shadow_ray calls */
    uint64_t shadow_rays;             /* subset of rays */
    uint64_t tlas_node_pops;
    uint64_t instances_visited;       /* TLAS leaf primitives tested (transform_ray executed) */
    uint64_t mesh_intersection_count; /* g_stats.mesh_intersection_count */
    uint64_t mesh_bvh_traversals;     /* g_stats.mesh_bvh_traversals  (BLAS node pops) */
    uint64_t mesh_node_traversals;    /* g_stats.mesh_node_traversals (BLAS inner nodes entered) */
    uint64_t mesh_leaf_traversals;    /* g_stats.mesh_leaf_traversals */
    uint64_t triangles_tested;
    uint64_t samples;                 /* integrator invocations */
    /* the same traversal counters restricted to shadow rays (so closest-hit = total - shadow) */
    uint64_t shadow_tlas_node_pops, shadow_instances_visited, shadow_mesh_intersection_count,
             shadow_mesh_bvh_traversals, shadow_mesh_node_traversals, shadow_mesh_leaf_traversals,
             shadow_triangles_tested;
} bpt_stats;

/* Per-sample record for parity tests (bpt_render_pass with a record buffer attached). */
typedef struct bpt_sample_record {
    float ray_o[3];
    float ray_d[3];
    float radiance[3];                /* integrator result before vignette (raytracer.cpp:467) */
    uint32_t rays;                    /* closest + shadow rays this sample traced */
} bpt_sample_record;

enum { BPT_SEED_PER_PIXEL = 0 };      /* entropy = random_seed(hash_coordinate(x, y, sample_index) ^ salt)
                                         (samplers.h:14-18, :92-108); see SURVEY 8b "RNG contract" */

/* ---------------------------------------------------------------------------------
 * 1. Host scene model (no GPU needed)
 * ------------------------------------------------------------------------------- */
typedef struct bpt_scene bpt_scene;

BPT_API const char* bpt_last_error(void);
BPT_API const char* bpt_version(void);

/* clear_scene + init_scene (scene.cpp:244-254, raytracer.cpp:1424-1453): null material and null
 * primitive at index 0, reference default settings (Appendix C), Mitchell-Netravali filter. */
BPT_API bpt_scene* bpt_scene_create(void);
BPT_API void       bpt_scene_destroy(bpt_scene* scene);

/* scene.cpp:9-62.  Return the new MaterialID. */
BPT_API uint32_t bpt_add_material(bpt_scene* s, const bpt_material* source_material);
BPT_API uint32_t bpt_add_diffuse_material(bpt_scene* s, const float diffuse_color[3], float ior, float roughness,
                                          int32_t checkers, const float checker_color[3]);
BPT_API uint32_t bpt_add_translucent_material(bpt_scene* s, const float absorb[3], float ior, float roughness);
BPT_API uint32_t bpt_add_emissive_material(bpt_scene* s, const float emission_color[3]);

/* scene.cpp:107-159.  transform may be NULL (identity).  Return the new PrimitiveID
 * (planes index Scene::planes, everything else Scene::primitives). */
BPT_API uint32_t bpt_add_plane (bpt_scene* s, uint32_t material_id, const float n[3], float d);
BPT_API uint32_t bpt_add_sphere(bpt_scene* s, uint32_t material_id, float r, const bpt_m4x4inv* transform);
BPT_API uint32_t bpt_add_box   (bpt_scene* s, uint32_t material_id, const float r[3], const bpt_m4x4inv* transform);

/* A Mesh value (primitives.h:63-73): triangle_count Triangles of 9 floats (a,b,c), optional
 * per-vertex normals in the same shape (has_normals).  The data is copied.  Instances made with
 * bpt_add_mesh share the mesh and its BLAS exactly like Mesh copies share `bvh` (scene.cpp:146-154). */
BPT_API uint32_t bpt_create_mesh(bpt_scene* s, uint32_t triangle_count, const float* positions, const float* normals);
BPT_API uint32_t bpt_add_mesh  (bpt_scene* s, uint32_t material_id, uint32_t mesh, const bpt_m4x4inv* transform);

/* Scene::top_sky_color / bot_sky_color / skydome (scene.h:92-96, assets.h:29-37). pixels = w*h*3 floats, copied. */
BPT_API int bpt_set_sky(bpt_scene* s, const float top[3], const float bot[3]);
/* Scene::ambient_light (scene.h:102), read by the "Whitted" integrator only (integrators.cpp:371); default 0 */
BPT_API int bpt_set_ambient_light(bpt_scene* s, const float rgb[3]);
BPT_API int bpt_set_skydome(bpt_scene* s, uint32_t w, uint32_t h, const float* pixels);

/* Scene::new_camera / new_settings latched as render_all_tiles does (raytracer.cpp:711-720,
 * including recompute_camera :50-59). */
BPT_API int bpt_get_camera(const bpt_scene* s, bpt_camera* out);
BPT_API int bpt_set_camera(bpt_scene* s, const bpt_camera* camera);
/* aim_camera / aim_camera_at (raytracer.cpp:26-48) on the scene's camera. */
BPT_API int bpt_aim_camera(bpt_scene* s, const float camera_d[3]);
BPT_API int bpt_aim_camera_at(bpt_scene* s, const float at[3]);
BPT_API int bpt_get_settings(const bpt_scene* s, bpt_settings* out);
BPT_API int bpt_set_settings(bpt_scene* s, const bpt_settings* settings);

/* find_integrator / find_filter + load_reconstruction_kernel (integrators.cpp:834-845,
 * reconstruction_filters.cpp:110-121, raytracer.cpp:164-185).  Unknown names fall back to the
 * reference's defaults (integrator 0 / Box filter) exactly as the reference does. */
BPT_API int bpt_find_integrator(const char* name);
BPT_API int bpt_load_reconstruction_kernel(bpt_scene* s, const char* filter_name);
BPT_API int bpt_get_filter_cache(const bpt_scene* s, bpt_filter_cache* out);
BPT_API int bpt_set_filter_cache(bpt_scene* s, const bpt_filter_cache* cache);

/* create_scene_bvh (scene.cpp:173-242): per-mesh create_bvh_for_mesh(BVH_SAHBinned, Scalar)
 * (bvh.cpp:342-426) then the TLAS create_bvh (bvh.cpp:328-340).  Output is bit-identical. */
BPT_API int bpt_create_scene_bvh(bpt_scene* s);

/* Read-only views for parity checks; pointers stay valid until the scene is modified/destroyed. */
BPT_API int bpt_get_scene_bvh(const bpt_scene* s, const bpt_bvh_node** nodes, uint32_t* node_count,
                              const uint32_t** indices, uint32_t* index_count);
BPT_API int bpt_get_mesh_bvh(const bpt_scene* s, uint32_t mesh, const bpt_bvh_node** nodes, uint32_t* node_count,
                             const uint32_t** indices, uint32_t* index_count,
                             const float** leaf_order_triangles /* 9 floats each */);
/* The same trees in the layout the device traverses (csrc/wide_bvh.h: 64-byte sibling pairs grouped into two-level
 * records; every child carries the reference node's box verbatim plus a packed reference).  mesh < 0 selects the TLAS.
 * pairs points at 2*pair_count children; big_leaves at big_leaf_count {first, count} pairs.  Read-only, for parity checks
 * of the re-layout (tests/test_wide_layout.py decodes it back into the reference node array). */
typedef struct bpt_wide_child { float bv_p[3]; float bv_r[3]; uint32_t ref; uint32_t aux; } bpt_wide_child;
BPT_API int bpt_get_wide_bvh(const bpt_scene* s, int32_t mesh, const bpt_wide_child** pairs, uint32_t* pair_count,
                             bpt_wide_child* root, const uint32_t** big_leaves, uint32_t* big_leaf_count, uint32_t* depth);
/* ---- asset readers: the reference's parse_obj (Raytracer/assets.cpp:187-400) and parse_hdr (:411-600),
 * same input -> same triangles / texels, quirks included (see csrc/obj_hdr_readers.cpp).  Malformed input that would make the
 * reference read out of bounds or loop forever returns an error here. */
typedef struct bpt_obj bpt_obj;
enum { BPT_WINDING_CLOCKWISE = 0, BPT_WINDING_COUNTER_CLOCKWISE = 1 };        /* MeshWinding, assets.h:74-77 */
BPT_API bpt_obj* bpt_parse_obj(const char* text, int32_t winding);           /* NUL-terminated text; NULL on error */
BPT_API bpt_obj* bpt_load_obj(const char* path, int32_t winding);
BPT_API void bpt_obj_free(bpt_obj* obj);
BPT_API uint32_t bpt_obj_triangle_count(const bpt_obj* obj);
BPT_API const float* bpt_obj_positions(const bpt_obj* obj);                  /* 9 floats per triangle */
BPT_API const float* bpt_obj_normals(const bpt_obj* obj);                    /* 9 floats per triangle, or NULL */
BPT_API const float* bpt_obj_texcoords(const bpt_obj* obj);                  /* 9 floats per triangle, or NULL */
BPT_API uint32_t bpt_create_mesh_from_obj(bpt_scene* s, const bpt_obj* obj); /* load_mesh, raytracer.cpp:148-158: midpoint-split BVH */
/* Radiance .hdr (32-bit_rle_rgbe): pixels == NULL only reports the size; otherwise w*h*3 floats, row order as the
 * reference stores them (Image_V3, assets.h:29-32). */
BPT_API int bpt_parse_hdr(const char* data, size_t size, uint32_t* w, uint32_t* h, float* pixels);
BPT_API int bpt_load_skydome_hdr(bpt_scene* s, const char* path);            /* parse + bpt_set_skydome */

/* bpt_create_mesh with the construction method spelled out (BVHConstructionMethod, bvh.h:7-11).  bpt_create_mesh uses
 * BPT_BVH_SAH_BINNED like create_scene_bvh (scene.cpp:208); the reference's load_mesh builds OBJ meshes with
 * BPT_BVH_MIDPOINT_SPLIT (raytracer.cpp:154).  BPT_BVH_SAH_FULL is O(n^2), as in the reference. */
enum { BPT_BVH_MIDPOINT_SPLIT = 0, BPT_BVH_SAH_BINNED = 1, BPT_BVH_SAH_FULL = 2 };
BPT_API uint32_t bpt_create_mesh_ex(bpt_scene* s, uint32_t triangle_count, const float* positions, const float* normals,
                                    int32_t method);

/* bpt_create_mesh with a BVH supplied by the caller (e.g. from bpt_build_mesh_bvh_device) instead of the host build. */
BPT_API uint32_t bpt_create_mesh_with_bvh(bpt_scene* s, uint32_t triangle_count, const float* positions, const float* normals,
                                          const bpt_bvh_node* nodes, uint32_t node_count, const uint32_t* indices);
BPT_API int bpt_get_counts(const bpt_scene* s, uint32_t* materials, uint32_t* primitives, uint32_t* planes,
                           uint32_t* lights, uint32_t* meshes);

/* Procedural inputs for the BASELINE.json configs.  Write triangle_count*9 floats;
 * a NULL `positions` just returns the count.  level-L icosphere has 20*4^L triangles; each vertex is
 * normalised then pushed out by 1 + amplitude*sin(9x)*sin(7y)*sin(11z). */
BPT_API uint32_t bpt_make_displaced_icosphere(uint32_t level, float amplitude, float* positions);
/* Closed-form HDR environment (gradient + sun disc), w*h*3 floats. */
BPT_API int bpt_make_procedural_skydome(uint32_t w, uint32_t h, float* pixels);

/* ---------------------------------------------------------------------------------
 * 2. Device renderer (sm_100a; no CPU fallback)
 * ------------------------------------------------------------------------------- */
typedef struct bpt_ctx bpt_ctx;

BPT_API int  bpt_create(int device, bpt_ctx** out_ctx);
BPT_API void bpt_destroy(bpt_ctx* ctx);

/* The tables behind get_next_sample_1d/2d (samplers.cpp:140-397 g_strata_permutation_sets[256][64];
 * blue_noise_samplers/..._256spp.cpp sobol_256spp_256d[65536], scramblingTile[131072],
 * rankingTile[131072], values 0..255).  Passed in by the caller as bytes; copied to the device. */
BPT_API int bpt_set_sampler_tables(bpt_ctx* ctx, const uint8_t* strata_permutation_sets,
                                   const uint8_t* sobol, const uint8_t* scrambling_tile, const uint8_t* ranking_tile);

/* Flatten + re-lay-out + upload: materials, planes, primitives with their transforms, lights,
 * TLAS, one BLAS per unique mesh (child-pair nodes, 48-B triangles), skydome, camera, settings,
 * filter LUT.  Replaces handing `Scene*` to the worker threads (Raytracer.h:50-55). */
BPT_API int bpt_upload_scene(bpt_ctx* ctx, const bpt_scene* scene);
/* The same without waiting: the upload runs on a copy stream into the scene buffer that is NOT being rendered (the device
 * holds two), so passes enqueued earlier keep running over the previous scene; the next bpt_render_pass / bpt_trace
 * switches to the new one (ordered on the device, no host wait).  `scene` must stay unmodified until then.  This is how a
 * host loop that re-submits its scene every frame overlaps the transfer with rendering (bench.py's e2e leg). */
BPT_API int bpt_upload_scene_async(bpt_ctx* ctx, const bpt_scene* scene);
/* Re-latch camera/settings/filter only (render_all_tiles :700-720) without touching geometry. */
BPT_API int bpt_update_settings(bpt_ctx* ctx, const bpt_scene* scene);

/* AccumulationBuffer (Raytracer.h:44-48; raytracer.cpp:501-522): V4{sum w*rgb, sum w} per pixel. */
BPT_API int bpt_film_resize(bpt_ctx* ctx, uint32_t w, uint32_t h);       /* allocates + zeroes */
BPT_API int bpt_film_clear(bpt_ctx* ctx);
BPT_API int bpt_film_use_external(bpt_ctx* ctx, void* device_ptr, uint32_t w, uint32_t h); /* caller-owned device buffer (e.g. a torch tensor for the NCCL reduce) */
BPT_API int bpt_film_device_ptr(bpt_ctx* ctx, void** out_device_ptr);
BPT_API int bpt_download_film(bpt_ctx* ctx, float* out_rgba /* w*h*4 floats, host */);
/* Snapshot + asynchronous read-back: the film (reduced != 0: the root's reduced film) is copied into a front buffer behind
 * the passes enqueued so far -- what render_all_tiles does when a pass completes (copy to the front buffer, raytracer.cpp:
 * 705-709) -- and that snapshot travels to `out_rgba` on a copy stream while later passes already run.  `out_rgba` should
 * be page-locked (bpt_host_register) for the copy to be asynchronous; it is valid after bpt_wait_download. */
BPT_API int bpt_download_film_async(bpt_ctx* ctx, float* out_rgba, int reduced);
BPT_API int bpt_wait_download(bpt_ctx* ctx);
BPT_API int bpt_host_register(void* host_ptr, size_t bytes);       /* cudaHostRegister / cudaHostUnregister */
BPT_API int bpt_host_unregister(void* host_ptr);

/* One progressive pass over the pixel rect [x0,x1) x [y0,y1): what render_all_tiles + every
 * render_tile call of the pass do (raytracer.cpp:366-495, :692-757).  `frame_count` is
 * AccumulationBuffer::frame_count (sample_index = frame_count + s).  Asynchronous; bpt_sync() or
 * bpt_download_film() waits.  Passes enqueued without a wait in between overlap on the device (the
 * next pass starts under the kernel tails of the previous one); every reader / clear of the film
 * issued between two passes still sees exactly the passes issued before it. */
BPT_API int bpt_render_pass(bpt_ctx* ctx, int32_t x0, int32_t y0, int32_t x1, int32_t y1,
                            uint32_t frame_count, uint32_t spp, uint32_t seed_mode, uint32_t seed_salt);
/* Same pass over a set of row bands [y0,y1) x [x0,x1) (n_bands pairs in y0y1) treated as ONE workload: the rows a
 * rank owns under the multi-GPU row-block partition. */
BPT_API int bpt_render_pass_bands(bpt_ctx* ctx, int32_t x0, int32_t x1, uint32_t n_bands, const int32_t* y0y1,
                                  uint32_t frame_count, uint32_t spp, uint32_t seed_mode, uint32_t seed_salt);
BPT_API int bpt_sync(bpt_ctx* ctx);

/* ---- resolve + post-process + "Render to bitmap" (raytracer.cpp:2103-2185, assets.cpp:671-724) ---- */
/* PostProcessSettings (Raytracer/scene.h:84-90) -- layout-compatible. */
typedef struct bpt_post_settings {
    float   exposure;
    int32_t tonemapping;
    int32_t srgb_transform;
    float   midpoint;
    float   contrast;
} bpt_post_settings;
/* film -> 0xAARRGGBB pixels exactly as the display loop does: NaN -> cyan, /w, max(0), 2^exposure, 1-exp(-x),
 * pow(x, 1/2.23333), sigmoidal contrast, *255, + TPDF dither from an RGB8 blue-noise tile (power-of-two w/h; NULL
 * skips the dither term), clamp, pack.  out = w*h uint32 on the host. */
BPT_API int bpt_resolve_bgra8(bpt_ctx* ctx, const bpt_post_settings* post, const uint8_t* dither_rgb8,
                              uint32_t dither_w, uint32_t dither_h, uint32_t* out_pixels);
/* write_bitmap (assets.cpp:693-724): 32-bit top-down BMP, byte-identical header. Host only. */
BPT_API int bpt_write_bitmap(const char* file_name, const uint32_t* pixels, uint32_t w, uint32_t h);

/* Parity / diagnostics. */
BPT_API int bpt_trace(bpt_ctx* ctx, uint32_t n, const bpt_ray* rays, int mode, uint32_t ignored_primitive,
                      bpt_hit* out_hits);
/* Attach a host buffer of (x1-x0)*(y1-y0)*spp records, filled (pixel-major, sample-minor) by the next
 * bpt_render_pass; NULL detaches. */
BPT_API int bpt_set_sample_records(bpt_ctx* ctx, bpt_sample_record* host_records, uint64_t capacity);
/* rays / shadow_rays / samples are always counted; the traversal counters need the counting instantiations. */
BPT_API int bpt_stats_enable(bpt_ctx* ctx, int enable);
BPT_API int bpt_get_stats(bpt_ctx* ctx, bpt_stats* out, int reset);
/* Shadow rays that never reach a mesh BLAS (planes / TLAS root / the spheres, boxes and mesh root boxes of a one-leaf TLAS
 * decide them: the head of intersect_shadow_ray, intersection.cpp:424-520) are settled inside the shading kernel and never
 * enter the traversal kernels.  bpt_set_ray_prefilter(0) turns that off.  A pass rendered with bpt_stats_enable(1)
 * traces every shadow ray in the traversal kernels (so that bpt_stats stays in the reference's units) and only counts the
 * rays a normal pass would settle early: out[0] = their number, out[1] = their algorithmic bytes (SURVEY 8d units). */
BPT_API int bpt_set_ray_prefilter(bpt_ctx* ctx, int enable);
BPT_API int bpt_get_ray_prefilter_stats(bpt_ctx* ctx, uint64_t out[2]);
/* GPU time of the last render pass by stage, measured with CUDA events on the context's stream (ms). */
typedef struct bpt_pass_timing {
    float total_ms, raygen_ms, trace_ms, shade_ms, shadow_ms, splat_ms;
    uint32_t kernel_launches;
    uint32_t trace_launches;
} bpt_pass_timing;
BPT_API int bpt_get_pass_timing(bpt_ctx* ctx, bpt_pass_timing* out);
/* per-stage CUDA-event timing (two event records around every kernel); off by default */
BPT_API int bpt_set_detailed_timing(bpt_ctx* ctx, int enable);
/* Scheduling knob (results do not depend on it): once at most `paths` paths of a batch survive, they finish inside one
 * fused launch instead of one wavefront round per bounce (the analogue of a render_tile worker simply looping on,
 * raytracer.cpp:409-494).  0 = always wavefront.  Default 65536. */
BPT_API int bpt_set_tail_threshold(bpt_ctx* ctx, uint32_t paths);
/* create_bvh_for_mesh (Raytracer/bvh.cpp:342-391, BVH_SAHBinned) built ON THE DEVICE: same node array (numbering
 * included) and same leaf-order indices as bpt_create_mesh / the reference produce on the host.  nodes_out needs room for
 * 2*triangle_count + 2 nodes.  build_ms (nullable) receives the device time of the build (entries resident). */
BPT_API int bpt_build_mesh_bvh_device(bpt_ctx* ctx, uint32_t triangle_count, const float* positions,
                                      int32_t method /* BPT_BVH_SAH_BINNED or BPT_BVH_MIDPOINT_SPLIT */, bpt_bvh_node* nodes_out, uint32_t node_capacity, uint32_t* node_count,
                                      uint32_t* indices_out, float* build_ms);

/* ---------------------------------------------------------------------------------
 * 3. Multi-GPU: the image is split into row bands over the GPUs of one box, the scene is replicated, every
 *    rank accumulates its bands (plus the filter's halo rows) into its own full-frame film, and ONE NCCL reduce per
 *    progressive pass sums the partial films on the root.  The reference's analogue is its worker threads sharing one
 *    AccumulationBuffer (raytracer.cpp:551-603); (sum w*rgb, sum w) is linear, so the sum of the partial films is the
 *    single-GPU film up to float-add order.  One context per GPU; one host thread (or process) per context.
 *    NCCL (libnccl.so.2) is loaded on first use of these entry points; the rest of the library does not need it.
 * ------------------------------------------------------------------------------- */
typedef struct bpt_nccl_id { char internal[128]; } bpt_nccl_id;               /* ncclUniqueId */
BPT_API int bpt_nccl_get_unique_id(bpt_nccl_id* out);                         /* rank 0 calls this and hands the id to the others */
/* ncclCommInitRank on the context's device; *out_comm is an ncclComm_t (usable with NCCL directly). */
BPT_API int bpt_nccl_comm_init_rank(bpt_ctx* ctx, const bpt_nccl_id* id, int nranks, int rank, void** out_comm);
/* ncclCommInitAll: one process driving ndev devices (the headless driver's --gpus N); out_comms[ndev]. */
BPT_API int bpt_nccl_comm_init_all(int ndev, const int* devices, void** out_comms);
BPT_API int bpt_nccl_comm_destroy(void* comm);
/* reduced film (root only) := sum over ranks of the ranks' films; enqueued on the context's stream behind the passes
 * rendered so far, no host synchronisation.  The partial films are NOT modified, so progressive passes keep accumulating
 * into them and the next bpt_reduce_film gives the sum of everything rendered so far (no double counting).
 * `nccl_comm` is an ncclComm_t (from bpt_nccl_comm_init_* or the caller's own NCCL of the same process).
 * Every rank of the communicator must make the call (it is a collective). */
BPT_API int bpt_reduce_film(bpt_ctx* ctx, void* nccl_comm, int root);
BPT_API int bpt_reduced_film_device_ptr(bpt_ctx* ctx, void** out_device_ptr); /* root: valid after the first bpt_reduce_film */
BPT_API int bpt_download_reduced_film(bpt_ctx* ctx, float* out_rgba /* w*h*4 floats, host */);
/* bpt_resolve_bgra8 on the root's reduced film */
BPT_API int bpt_resolve_reduced_bgra8(bpt_ctx* ctx, const bpt_post_settings* post, const uint8_t* dither_rgb8,
                                      uint32_t dither_w, uint32_t dither_h, uint32_t* out_pixels);

/* cumulative bytes this context copied host->device / device->host (scene uploads, rays, film, records) */
BPT_API int bpt_get_transfer_bytes(bpt_ctx* ctx, uint64_t* h2d, uint64_t* d2h, int reset);

#ifdef __cplusplus
}
#endif
#endif /* BPT_H */
