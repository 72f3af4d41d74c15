// Device-resident scene + wavefront state layouts (HBM), shared by the kernels and the uploader.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#include "../../include/bpt.h"
#include "wide_bvh.h"

namespace bpt {

// raised by kernels in DScene::error_flag, reported by the next synchronising entry point
enum { BPT_DEVERR_NONE = 0, BPT_DEVERR_STACK_OVERFLOW = 1, BPT_DEVERR_TRIP_LIMIT = 2 };

// ---- acceleration structures -------------------------------------------------------------------------------------
// Two-level pair records (wide_bvh.h): the reference's binary BVH re-laid-out on the host.  A child is two 128-bit
// words, a sibling pair four, a record (pair of X, pair of X's child 0, pair of X's child 1) twelve:
//   q0 = {bv_p.x, bv_p.y, bv_p.z, bv_r.x}   q1 = {bv_r.y, bv_r.z, ref, aux}
struct DChild { float4 q0; float4 q1; };
struct DPair { DChild c[2]; };
static_assert(sizeof(DChild) == sizeof(WChild) && sizeof(DPair) == sizeof(WPair), "device / host layouts of the pair records differ");

// Leaf triangles, leaf order, 48 bytes = 3 x 128-bit:  {a.xyz, original index}  {b-a, 0}  {c-a, 0}.
// b-a / c-a are the first two operations of ray_intersect_triangle (intersection.cpp:145-146); hoisting them to
// upload time is exact (same IEEE subtraction, once instead of per test).
struct DTriangle { float4 a_idx; float4 e1; float4 e2; };

struct DMesh {                 // 64 bytes
    float4   root_q0, root_q1; // the BLAS root as a child record: box + ref
    uint32_t pair_base;        // index of this BLAS's pair 0 in DScene::pairs
    uint32_t tri_base;         // index of its first DTriangle in DScene::triangles (also indexes normals)
    uint32_t triangle_count;
    uint32_t has_normals;
    uint32_t big_base;         // its first entry in DScene::big_leaves
    uint32_t pad[3];
};

// Primitive (primitives.h:92-106) flattened; rows 0..2 of the inverse / forward matrices (row 3 is never read
// on the path: transform() my_math.h:947-954, transform_normal :956-963, translation :975-983).
struct DPrimitive {
    float4 inv[3];
    float4 fwd[3];
    uint32_t type;
    uint32_t material;
    uint32_t mesh;             // index into DScene::meshes (Primitive_Mesh)
    float    sphere_r;
    float    box_r[3];
    uint32_t pad;
};

struct DPlane { float n[3]; float d; uint32_t material; uint32_t pad[3]; };

struct DMaterial {             // Material (scene.h:15-29) + padding to 80 bytes
    uint32_t flags;
    float albedo[3];
    float checker_color[3];
    float emission_color[3];
    float ior, metallic, roughness;
    int32_t is_participating_medium;
    float absorb[3];
    uint32_t pad[3];
};

struct DScene {
    const DPair*      pairs;           // TLAS pairs (from index 0) then every BLAS's pairs, back to back
    const uint2*      big_leaves;      // {first, count} of leaves with more than 7 items; TLAS entries first
    const uint32_t*   tlas_indices;
    float4            tlas_root_q0, tlas_root_q1;   // the TLAS root as a child record (in the kernel's constant bank)
    uint32_t*         error_flag;      // device-visible word the kernels raise on an internal limit (BPT_DEVERR_*)
    const DTriangle*  triangles;       // all meshes back to back, leaf order
    const float4*     normals;         // 3 x float4 per triangle (leaf order), only for has_normals meshes; may be null
    const DMesh*      meshes;
    const DPrimitive* primitives;
    const DPlane*     planes;
    const DMaterial*  materials;       // [material_count] + one extra "air" entry (integrators.cpp:597-599)
    const uint32_t*   lights;
    const float4*     skydome;         // w*h texels {r,g,b,0}; null -> gradient sky
    const uint8_t*    strata_perm;     // [256][64]
    const uint8_t*    bn_sobol;        // [256*256]
    const uint8_t*    bn_scramble;     // [128*128*8]
    const uint8_t*    bn_rank;         // [128*128*8]
    const float*      filter_lut;      // [512]

    uint32_t plane_count, primitive_count, material_count, light_count;
    uint32_t air_material;             // == material_count
    uint32_t skydome_w, skydome_h;
    float top_sky[3], bot_sky[3];
    float ambient_light[3];

    bpt_camera   camera;               // latched Scene::camera
    bpt_settings settings;             // latched Scene::settings
    uint32_t filter_radius;            // FilterCache::kernel_size
    uint32_t filter_lut_size;          // FilterCache::cache_size (0 = Box)
    uint32_t film_w, film_h;
    uint32_t tame_bounds;              // every TLAS/BLAS node box lies inside |x| < 1e15 (enables the FMNMX slab test)
    uint32_t prefilter;                // k_shade settles the shadow rays that never reach a BLAS itself: 0 off, 1 on, 2 count them only (shadow_tlas_head)
};

// ---- wavefront path state (SoA, one entry per path slot of the current batch) ------------------------------------
#define BPT_MATERIAL_STACK_DEPTH 64    // integrators.cpp:602

struct DPathState {
    float4*   ray_o;        // {o.xyz, max_t}
    float4*   ray_d;        // {d.xyz, unused}
    float4*   hit;          // {t, prim (bits), tri slot (bits), v}
    float*    hit_w;        // barycentric w
    float4*   throughput;   // {xyz, unused}
    float4*   radiance;     // {total_color.xyz, vignette}  (first written by the first bounce's shading)
    uint4*    rng;          // RandomSeries (samplers.h:29-34)
    float4*   prev_n;       // {prev_N.xyz, bits: flags}   flags: bit0 = is_specular_bounce
    float2*   jitter;       // AA jitter (raytracer.cpp:444-446)
    uint8_t*  mstack_at;    // material_stack_at
    uint16_t* mstack;       // [63][slots] level-major material ids of stack levels 1..63 (level 0 is always the integrator's "air")
    float4*   primary_d;    // {primary ray d.xyz, ray count} (records / vignette)
    float4*   primary_o;    // only written when records are requested
};

struct DShadowItem {        // one NEE shadow ray awaiting its occlusion test
    float4 o_maxt;          // {o.xyz, max_t}
    float4 d_light;         // {d.xyz, bits: ignored light primitive}
    float4 contrib_slot;    // {contribution.xyz, bits: slot}
};

struct DQueues {
    uint32_t* active[2];    // slot ids for the current / next bounce
    DShadowItem* shadow;
    uint32_t* counters;     // [0]=active in, [1]=active out, [2]=shadow count, [3]=trace fetch cursor, [4]=shadow fetch cursor
};

struct DStats {             // device mirror of bpt_stats
    unsigned long long v[20];   // [0..9] totals as in bpt_stats, [10..16] shadow-only traversal counters, [17] / [18] shadow rays settled
                                // inside k_shade and their algorithmic bytes (SURVEY 8d units), counted when shadow_prefilter == 2
};

} // namespace bpt
