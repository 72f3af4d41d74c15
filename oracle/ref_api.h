/*
 * ORACLE-ONLY (test infrastructure).  C ABI of oracle/_ref/libbpt_ref.so: the REFERENCE's own
 * translation units (compiled unmodified from /root/reference by oracle/Makefile) behind the same
 * scene-building calls as include/bpt.h, plus entry points that run the reference's renderer.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product library (libbpt.so) never links, loads or calls anything declared here.
 */
#ifndef BPT_REF_API_H
#define BPT_REF_API_H
#include "../include/bpt.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ref_scene ref_scene;

BPT_API ref_scene* ref_scene_create(void);
BPT_API void       ref_scene_destroy(ref_scene* s);
BPT_API uint32_t ref_add_material(ref_scene* s, const bpt_material* m);
BPT_API uint32_t ref_add_diffuse_material(ref_scene* s, const float c[3], float ior, float roughness, int32_t checkers, const float cc[3]);
BPT_API uint32_t ref_add_translucent_material(ref_scene* s, const float absorb[3], float ior, float roughness);
BPT_API uint32_t ref_add_emissive_material(ref_scene* s, const float e[3]);
BPT_API uint32_t ref_add_plane (ref_scene* s, uint32_t mat, const float n[3], float d);
BPT_API uint32_t ref_add_sphere(ref_scene* s, uint32_t mat, float r, const bpt_m4x4inv* xf);
BPT_API uint32_t ref_add_box   (ref_scene* s, uint32_t mat, const float r[3], const bpt_m4x4inv* xf);
BPT_API uint32_t ref_create_mesh(ref_scene* s, uint32_t triangle_count, const float* positions, const float* normals);
BPT_API uint32_t ref_create_mesh_ex(ref_scene* s, uint32_t triangle_count, const float* positions, const float* normals, int32_t method);
BPT_API uint32_t ref_add_mesh  (ref_scene* s, uint32_t mat, uint32_t mesh, const bpt_m4x4inv* xf);
BPT_API int ref_set_sky(ref_scene* s, const float top[3], const float bot[3]);
BPT_API int ref_set_ambient_light(ref_scene* s, const float rgb[3]);
BPT_API int ref_set_skydome(ref_scene* s, uint32_t w, uint32_t h, const float* pixels);
BPT_API int ref_get_camera(const ref_scene* s, bpt_camera* out);
BPT_API int ref_set_camera(ref_scene* s, const bpt_camera* c);
BPT_API int ref_aim_camera(ref_scene* s, const float d[3]);
BPT_API int ref_aim_camera_at(ref_scene* s, const float at[3]);
BPT_API int ref_get_settings(const ref_scene* s, bpt_settings* out);
BPT_API int ref_set_settings(ref_scene* s, const bpt_settings* in);
BPT_API int ref_find_integrator(const char* name);
BPT_API int ref_load_reconstruction_kernel(ref_scene* s, const char* filter_name);
BPT_API int ref_get_filter_cache(const ref_scene* s, bpt_filter_cache* out);
BPT_API int ref_set_filter_cache(ref_scene* s, const bpt_filter_cache* in);
BPT_API int ref_create_scene_bvh(ref_scene* s);
BPT_API int ref_get_scene_bvh(const ref_scene* s, const bpt_bvh_node** nodes, uint32_t* node_count, const uint32_t** indices, uint32_t* index_count);
BPT_API int ref_get_mesh_bvh(const ref_scene* s, uint32_t mesh, const bpt_bvh_node** nodes, uint32_t* node_count, const uint32_t** indices, uint32_t* index_count, const float** tris);
BPT_API int ref_get_counts(const ref_scene* s, uint32_t* materials, uint32_t* primitives, uint32_t* planes, uint32_t* lights, uint32_t* meshes);

/* Builds one of the reference's own SCENE_DESCRIPTIONs by name (g_scenes, raytracer.cpp:1409-1422)
 * through load_scene (:1455-1470).  Only scenes that need no LFS blob work (Week 1..5). */
BPT_API int ref_load_builtin_scene(ref_scene* s, const char* name, uint32_t w, uint32_t h);

/* sizeof() of the reference structs the ABI claims layout compatibility with. */
BPT_API int ref_sizeof(const char* type_name);

/* The sampler tables as bytes (samplers.cpp:140-397 and the vendored 256spp blue-noise file). */
BPT_API int ref_get_sampler_tables(uint8_t* perm /*16384*/, uint8_t* sobol /*65536*/, uint8_t* scramble /*131072*/, uint8_t* rank /*131072*/);

/* intersect_scene / intersect_shadow_ray (intersection.cpp:600-610) on a ray batch. */
BPT_API int ref_trace(ref_scene* s, uint32_t n, const bpt_ray* rays, int mode, uint32_t ignored_primitive, bpt_hit* out);
/* g_stats (intersection.h:33-40) since the last reset; only mesh_* and rays are filled. */
BPT_API int ref_get_stats(bpt_stats* out, int reset);

/* Single-threaded parity render: for y, for x, for s: entropy = random_seed(hash_coordinate(x,y,frame_count+s) ^ salt);
 * then the reference's own render_tile (raytracer.cpp:366-495) on the 1x1 rect {x,y} with spp = 1 and
 * AccumulationBuffer::frame_count = frame_count+s.  film = w*h*4 floats, accumulated into (not cleared).
 * records (nullable): (x1-x0)*(y1-y0)*spp entries, pixel-major / sample-minor. */
BPT_API int ref_render_parity(ref_scene* s, float* film, uint32_t w, uint32_t h,
                              int32_t x0, int32_t y0, int32_t x1, int32_t y1,
                              uint32_t frame_count, uint32_t spp, uint32_t salt, bpt_sample_record* records);

/* The reference's tile-multithreaded renderer, verbatim: init_work_queue / render_all_tiles / thread_proc /
 * render_tile with per-tile seeding (raytracer.cpp:551-757), `threads` workers + the calling thread helping
 * only to drain.  Renders ONE pass of `spp` samples; film (nullable) receives the front buffer.
 * seconds = the reference's own pass timer (:733-738). */
BPT_API int ref_render_threaded(ref_scene* s, uint32_t w, uint32_t h, uint32_t spp, uint32_t threads,
                                float* film, double* seconds, bpt_stats* stats);

/* The reference's display-loop resolve (raytracer.cpp:2103-2173: /w, exposure, 1-exp(-x), gamma, sigmoidal contrast, TPDF
 * blue-noise dither, BGRA8 pack), compiled from the reference's own text (oracle/tools/slice_resolve.py cuts the loop out
 * of raytracer.cpp at build time), and its write_bitmap (assets.cpp:693-724).  dither_rgb8: a power-of-two RGB8 tile
 * (the reference always dithers; there is no NULL path). */
BPT_API int ref_resolve_bgra8(const float* film_rgba, uint32_t w, uint32_t h, const bpt_post_settings* post,
                              const uint8_t* dither_rgb8, uint32_t dither_w, uint32_t dither_h, uint32_t* out_pixels);
BPT_API int ref_write_bitmap(const char* file_name, const uint32_t* pixels, uint32_t w, uint32_t h);

/* The reference's own parse_obj / parse_hdr (assets.cpp:187-400, :423-600).  Two-call pattern: NULL outputs = report sizes. */
BPT_API int ref_parse_obj(const char* text, int winding, uint32_t* triangle_count, int* has_normals, int* has_texcoords,
                          float* positions, float* normals, float* texcoords);
BPT_API int ref_parse_hdr(const char* data, size_t size, uint32_t* w, uint32_t* h, float* pixels);

/* Known-answer helpers straight from the reference's static functions (for device-math unit parity). */
BPT_API void ref_kat_cosine_hemisphere(const float n[3], const float u[2], float out[3]);   /* integrators.cpp:107-119 */
BPT_API void ref_kat_hemisphere(const float n[3], const float u[2], float out[3]);          /* integrators.cpp:93-105 */
BPT_API float ref_kat_fresnel(float cos_i, float eta_i, float eta_t, float* cos_t);         /* integrators.cpp:235-258 */
BPT_API void ref_kat_random_seed(uint32_t seed, uint32_t out_state[4]);                     /* samplers.h:92-108 */
BPT_API void ref_kat_sample_2d(uint32_t state[4], int strategy, uint32_t index, uint32_t x, uint32_t y, int dim, uint32_t bounce, float out[2]); /* samplers.cpp:18-90 */
BPT_API float ref_kat_sample_1d(uint32_t state[4], int strategy, uint32_t index, uint32_t x, uint32_t y, int dim, uint32_t bounce);              /* samplers.cpp:92-138 */
BPT_API float ref_kat_filter(const char* name, float x);                                    /* reconstruction_filters.cpp */

#ifdef __cplusplus
}
#endif
#endif
