// ORACLE-ONLY (test infrastructure).  Wrapper TU: compiles the reference's intersection.cpp UNMODIFIED
// (by #include, so its static functions are reachable) and adds the batch-trace entry point.
#include "intersection.cpp"      // -> /root/reference/Raytracer/intersection.cpp via the symlink farm
#include "oracle_internal.h"

volatile int g_ref_count_rays = 0;
volatile u64 g_ref_rays = 0;
volatile u64 g_ref_shadow_rays = 0;

// Linked with -Wl,--wrap=<mangled intersect_scene / intersect_shadow_ray>: every call the reference's
// integrators make goes through these, which only count and forward.
extern "C" {
Primitive* __real__Z15intersect_sceneP5SceneRK3RayPfPN4math2V3ES7_(Scene*, const Ray&, f32*, V3*, V3*);
b32 __real__Z20intersect_shadow_rayP5SceneRK3Ray10TypesafeIDI12PrimitiveTagjE(Scene*, const Ray&, PrimitiveID);

__attribute__((visibility("default")))
Primitive* __wrap__Z15intersect_sceneP5SceneRK3RayPfPN4math2V3ES7_(Scene* scene, const Ray& ray, f32* t, V3* p, V3* n) {
    if (g_ref_count_rays) __atomic_add_fetch(&g_ref_rays, 1, __ATOMIC_RELAXED);
    return __real__Z15intersect_sceneP5SceneRK3RayPfPN4math2V3ES7_(scene, ray, t, p, n);
}
__attribute__((visibility("default")))
b32 __wrap__Z20intersect_shadow_rayP5SceneRK3Ray10TypesafeIDI12PrimitiveTagjE(Scene* scene, const Ray& ray, PrimitiveID ignored) {
    if (g_ref_count_rays) { __atomic_add_fetch(&g_ref_rays, 1, __ATOMIC_RELAXED); __atomic_add_fetch(&g_ref_shadow_rays, 1, __ATOMIC_RELAXED); }
    return __real__Z20intersect_shadow_rayP5SceneRK3Ray10TypesafeIDI12PrimitiveTagjE(scene, ray, ignored);
}
}

static u32 primitive_to_id(Scene* scene, Primitive* p) {
    if (!p) return BPT_HIT_MISS;
    Primitive* planes = scene->planes.data;
    if (planes && p >= planes && p < planes + scene->planes.count) return BPT_HIT_PLANE | (u32)(p - planes);
    return (u32)(p - scene->primitives.data);
}

extern "C" BPT_API int
ref_trace(ref_scene* s, uint32_t n, const bpt_ray* rays, int mode, uint32_t ignored_primitive, bpt_hit* out) {
    Scene* scene = &s->scene;
    for (uint32_t i = 0; i < n; ++i) {
        const bpt_ray* r = &rays[i];
        Ray ray = make_ray(to_v3(r->o), to_v3(r->d), r->max_t);
        bpt_hit h;
        memset(&h, 0, sizeof(h));
        h.triangle = 0xFFFFFFFFu;
        if (mode == BPT_TRACE_CLOSEST) {
            f32 t = 0; V3 I = {}, N = {};
            Primitive* prim = intersect_scene(scene, ray, &t, &I, &N);
            h.t = t;
            h.primitive = primitive_to_id(scene, prim);
            if (prim) {
                h.n[0] = N.x; h.n[1] = N.y; h.n[2] = N.z;
                h.p[0] = I.x; h.p[1] = I.y; h.p[2] = I.z;
                if (prim->type == Primitive_Mesh) {
                    // The reference keeps hit_triangle_index in a local (intersection.cpp:435); recover it by
                    // re-running its own intersect_mesh on the winning instance with the same object-space ray.
                    Ray obj = transform_ray(ray, prim->transform->inverse);
                    f32 t2 = ray.max_t; u32 tri = 0xFFFFFFFFu; V3 uvw, a, b, c;
                    u64 s0 = g_stats.mesh_intersection_count, s1 = g_stats.mesh_bvh_traversals,
                        s2 = g_stats.mesh_node_traversals,    s3 = g_stats.mesh_leaf_traversals;
                    if (intersect_mesh(&prim->mesh, obj, Intersect_Full, &t2, &tri, &uvw, &a, &b, &c)) {
                        h.triangle = tri;
                        if (t2 != t) h.triangle = 0xFFFFFFFEu;   // would flag an inconsistency
                    }
                    g_stats.mesh_intersection_count = s0; g_stats.mesh_bvh_traversals = s1;   // do not let the
                    g_stats.mesh_node_traversals = s2;    g_stats.mesh_leaf_traversals = s3;  // re-run skew g_stats
                }
            }
        } else {
            // intersect_shadow_ray returns only b32; call the internal to learn WHICH primitive occluded.
            Primitive* prim = intersect_scene_internal(scene, ray, Intersect_Occlusion, PrimitiveID::from(ignored_primitive));
            if (g_ref_count_rays) { __atomic_add_fetch(&g_ref_rays, 1, __ATOMIC_RELAXED); __atomic_add_fetch(&g_ref_shadow_rays, 1, __ATOMIC_RELAXED); }
            h.t = r->max_t;
            h.primitive = primitive_to_id(scene, prim);
        }
        out[i] = h;
    }
    return 0;
}

extern "C" BPT_API int
ref_get_stats(bpt_stats* out, int reset) {
    memset(out, 0, sizeof(*out));
    out->rays = g_ref_rays;
    out->shadow_rays = g_ref_shadow_rays;
    out->mesh_intersection_count = g_stats.mesh_intersection_count;
    out->mesh_bvh_traversals     = g_stats.mesh_bvh_traversals;
    out->mesh_node_traversals    = g_stats.mesh_node_traversals;
    out->mesh_leaf_traversals    = g_stats.mesh_leaf_traversals;
    if (reset) {
        g_ref_rays = 0; g_ref_shadow_rays = 0;
        g_stats.mesh_intersection_count = 0; g_stats.mesh_bvh_traversals = 0;
        g_stats.mesh_node_traversals = 0;    g_stats.mesh_leaf_traversals = 0;
    }
    return 0;
}
