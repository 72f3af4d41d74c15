// Two-level BVH traversal + primitive tests, equivalent hit-for-hit to the reference's
// intersect_scene_internal / intersect_mesh (Raytracer/intersection.cpp:12-182, :243-401, :411-520).
//
// Re-designed for SIMT rather than transcribed:
//   * one flattened state machine walks TLAS and BLAS with a single per-thread stack (the reference nests
//     intersect_mesh's loop inside the TLAS leaf loop, which would serialise a warp);
//   * an inner node fetches BOTH children as one 64-byte sibling record (4 x LDG.128) and slab-tests them
//     together; the near child is entered directly, the far child is pushed with its entry distance.
// Equivalence argument (tie semantics): the reference pushes far then near and
// re-tests each box when popped, against the t of that moment.  (tn < tf && tf > 0) does not depend on t, so
// it is evaluated once at the parent; `tn < t` is evaluated for the near child immediately (nothing happens
// between the reference's push and pop of it) and for the far child when it is popped, with the stored tn.
// Leaves are therefore visited in the reference's order with the reference's culling, so equal-t ties
// (`t <= out_t` accepts, intersection.cpp:174) resolve to the same triangle.
#pragma once
#include "device_math.cuh"
#include "device_scene.cuh"

namespace bpt {

struct RayT {
    V3 o, d, inv;
    uint32_t neg;          // bit k = d_is_negative[k]  (intersection.h:13-24)
};

BPT_D void make_ray(RayT& r, V3 o, V3 d) {
    r.o = o; r.d = d;
    r.inv = 1.0f / d;
    r.neg = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
}

// ray_intersect_bounding_volume (intersection.cpp:107-133) split into its t-independent part and tn
BPT_D bool slab_test(const RayT& r, float px, float py, float pz, float rx, float ry, float rz, float& tn) {
    float nx = r.inv.x*(r.o.x - px), ny = r.inv.y*(r.o.y - py), nz = r.inv.z*(r.o.z - pz);
    float kx = fabsf(r.inv.x)*rx,    ky = fabsf(r.inv.y)*ry,    kz = fabsf(r.inv.z)*rz;
    float t1x = -nx - kx, t1y = -ny - ky, t1z = -nz - kz;
    float t2x = -nx + kx, t2y = -ny + ky, t2z = -nz + kz;
    tn = max_t(max_t(t1x, t1y), t1z);
    float tf = min_t(min_t(t2x, t2y), t2z);
    return (tn < tf) && (tf > 0.0f);
}

// ray_intersect_triangle (intersection.cpp:135-182) with edge1/edge2 hoisted to upload time
BPT_D bool triangle_test(const RayT& r, V3 a, V3 e1, V3 e2, float& t_io, float& out_v, float& out_w) {
    const float epsilon = 0.000000001f;
    V3 pvec = cross(r.d, e2);
    float det = dot(e1, pvec);
    if (det > -epsilon && det < epsilon) return false;
    float inv_det = 1.0f / det;
    V3 tvec = r.o - a;
    float v = dot(tvec, pvec)*inv_det;
    if (v < 0.0f || v > 1.0f) return false;
    V3 qvec = cross(tvec, e1);
    float w = dot(r.d, qvec)*inv_det;
    if (w < 0.0f || v + w > 1.0f) return false;
    float t = dot(e2, qvec)*inv_det;
    if ((t < epsilon) || (t_io < t)) return false;
    t_io = t; out_v = v; out_w = w;
    return true;
}

BPT_D bool sphere_test(const RayT& r, float sphere_r, float& t_io) {      // intersection.cpp:44-74
    float r_sq = sphere_r*sphere_r;
    float b = dot(r.d, r.o);
    float c = dot(r.o, r.o) - r_sq;
    float discr = b*b - c;
    if (discr >= 0.0f) {
        float root = sqrtf(discr);
        float tn = -b - root;
        float tf = -b + root;
        float t = (tn >= 0.0f ? tn : tf);
        if ((t >= kEps) && (t_io > t)) { t_io = t; return true; }
    }
    return false;
}

BPT_D bool box_test(const RayT& r, float rx, float ry, float rz, float& t_io) {   // intersection.cpp:76-105
    float tn;
    float nx = r.inv.x*r.o.x, ny = r.inv.y*r.o.y, nz = r.inv.z*r.o.z;
    float kx = fabsf(r.inv.x)*rx, ky = fabsf(r.inv.y)*ry, kz = fabsf(r.inv.z)*rz;
    float t1x = -nx - kx, t1y = -ny - ky, t1z = -nz - kz;
    float t2x = -nx + kx, t2y = -ny + ky, t2z = -nz + kz;
    tn = max_t(max_t(t1x, t1y), t1z);
    float tf = min_t(min_t(t2x, t2y), t2z);
    if (tn < tf) {
        float t = (tn >= 0.0f ? tn : tf);
        if ((t_io > t) && (t >= kEps)) { t_io = t; return true; }
    }
    return false;
}

BPT_D bool plane_test(const RayT& r, V3 n, float dist, float& t_io) {     // intersection.cpp:12-42
    float denom = dot(n, r.d);
    if (denom < -kEps) {
        float t = (dist - dot(n, r.o)) / denom;
        if ((t >= kEps) && (t < t_io)) { t_io = t; return true; }
    }
    return false;
}

struct HitRecord {
    float    t;
    uint32_t prim;       // BPT_HIT_MISS, BPT_HIT_PLANE|i, or primitive index
    uint32_t tri;        // global DTriangle slot (mesh hits), else 0xFFFFFFFF
    float    v, w;       // barycentrics of the hit triangle
};

struct TraceCounters {   // per-thread, flushed by the caller
    uint32_t tlas_pops, instances, mesh_calls, blas_pops, blas_inner, blas_leaves, tris;
};

#define BPT_STACK_DEPTH 64     // the reference's node_stack[64] (intersection.cpp:261, :445), far children only here

template <bool OCCLUSION, bool STATS>
BPT_D void trace_ray(const DScene& sc, V3 o, V3 d, float max_t, uint32_t ignored, HitRecord& out, TraceCounters& ctr) {
    RayT wray;
    make_ray(wray, o, d);
    float t = max_t;
    uint32_t hit_prim = BPT_HIT_MISS, hit_tri = 0xFFFFFFFFu;
    float hit_v = 0.0f, hit_w = 0.0f;

    // planes first, linearly (intersection.cpp:424-433); in occlusion mode a plane hit does not return early
    for (uint32_t i = 0; i < sc.plane_count; ++i) {
        const DPlane& pl = sc.planes[i];
        if (plane_test(wray, v3(__ldg(&pl.n[0]), __ldg(&pl.n[1]), __ldg(&pl.n[2])), __ldg(&pl.d), t)) hit_prim = BPT_HIT_PLANE | i;
    }

    uint32_t stk_lf[BPT_STACK_DEPTH];
    uint32_t stk_ca[BPT_STACK_DEPTH];
    float    stk_tn[BPT_STACK_DEPTH];
    int sp = 0;

    enum { S_NODE, S_ITEMS, S_POP, S_DONE };
    int state;
    int level = 0;                       // 0 = TLAS, 1 = inside a mesh BLAS
    RayT ray = wray;
    const DNodeHalf* nodes = sc.tlas_nodes;
    uint32_t cur_lf, cur_ca;
    uint32_t leaf_i = 0, leaf_end = 0;   // TLAS leaf items still to test
    int blas_sp = 0;
    uint32_t cur_prim = 0, cur_tri_base = 0;
    uint32_t c_pops = 0, c_inner = 0, c_leaves = 0;   // per intersect_mesh call (see STATS note below)

    {   // TLAS root (popped and box-tested like any node, intersection.cpp:450-454)
        float4 q0 = __ldg(&nodes[0].q0), q1 = __ldg(&nodes[0].q1);
        float tn;
        bool hit = slab_test(ray, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tn) && (tn < t);
        if (STATS) ctr.tlas_pops += 1;
        cur_lf = __float_as_uint(q1.z); cur_ca = __float_as_uint(q1.w);
        state = hit ? S_NODE : S_DONE;
    }

    while (state != S_DONE) {
        if (state == S_NODE) {
            uint32_t count = cur_ca & 0xFFFFu;
            if (count == 0) {
                // inner node: fetch the sibling pair, test both, enter near, push far
                const DNodeHalf* pr = nodes + cur_lf;
                float4 l0 = __ldg(&pr[0].q0), l1 = __ldg(&pr[0].q1);
                float4 r0 = __ldg(&pr[1].q0), r1 = __ldg(&pr[1].q1);
                float tnl, tnr;
                bool hl = slab_test(ray, l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, tnl);
                bool hr = slab_test(ray, r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, tnr);
                if (STATS) { if (level) { c_pops += 2; c_inner += 1; } else ctr.tlas_pops += 2; }
                bool right_first = (ray.neg >> (cur_ca >> 16)) & 1u;     // intersection.cpp:365-373, :509-517
                bool  near_hit = right_first ? hr : hl,            far_hit = right_first ? hl : hr;
                float near_tn  = right_first ? tnr : tnl,          far_tn  = right_first ? tnl : tnr;
                uint32_t near_lf = __float_as_uint(right_first ? r1.z : l1.z), far_lf = __float_as_uint(right_first ? l1.z : r1.z);
                uint32_t near_ca = __float_as_uint(right_first ? r1.w : l1.w), far_ca = __float_as_uint(right_first ? l1.w : r1.w);
                if (far_hit && sp < BPT_STACK_DEPTH) {
                    stk_lf[sp] = far_lf; stk_ca[sp] = far_ca; stk_tn[sp] = far_tn; ++sp;
                }
                if (near_hit && near_tn < t) { cur_lf = near_lf; cur_ca = near_ca; }
                else state = S_POP;
            } else if (level == 1) {
                // BLAS leaf: contiguous triangles in leaf order (intersection.cpp:285-307)
                if (STATS) { c_leaves += 1; ctr.tris += count; }
                const DTriangle* tri = sc.triangles + cur_tri_base + cur_lf;
                for (uint32_t k = 0; k < count; ++k) {
                    float4 a = __ldg(&tri[k].a_idx), e1 = __ldg(&tri[k].e1), e2 = __ldg(&tri[k].e2);
                    if (triangle_test(ray, v3(a), v3(e1), v3(e2), t, hit_v, hit_w)) {
                        hit_tri = cur_tri_base + cur_lf + k;
                        hit_prim = cur_prim;     // == "hit_any" of the enclosing intersect_mesh call
                        if (OCCLUSION) { out.t = t; out.prim = hit_prim; out.tri = hit_tri; out.v = hit_v; out.w = hit_w; return; }
                    }
                }
                state = S_POP;
            } else {
                leaf_i = cur_lf; leaf_end = cur_lf + count;
                state = S_ITEMS;
            }
        }

        if (state == S_ITEMS) {
            // one TLAS leaf item per iteration (intersection.cpp:461-500)
            if (leaf_i >= leaf_end) {
                state = S_POP;
            } else {
                uint32_t prim_index = __ldg(&sc.tlas_indices[leaf_i++]);
                if (prim_index != ignored) {
                    const DPrimitive* prim = sc.primitives + prim_index;
                    float4 m0 = __ldg(&prim->inv[0]), m1 = __ldg(&prim->inv[1]), m2 = __ldg(&prim->inv[2]);
                    float4 m[3] = {m0, m1, m2};
                    RayT oray;
                    make_ray(oray, xform(m, wray.o, 1.0f), xform(m, wray.d, 0.0f));     // transform_ray :403-409
                    if (STATS) ctr.instances += 1;
                    uint32_t type = __ldg(&prim->type);
                    if (type == BPT_PRIM_SPHERE) {
                        if (sphere_test(oray, __ldg(&prim->sphere_r), t)) {
                            hit_prim = prim_index; hit_tri = 0xFFFFFFFFu;
                            if (OCCLUSION) { out.t = t; out.prim = hit_prim; out.tri = hit_tri; out.v = 0; out.w = 0; return; }
                        }
                    } else if (type == BPT_PRIM_BOX) {
                        if (box_test(oray, __ldg(&prim->box_r[0]), __ldg(&prim->box_r[1]), __ldg(&prim->box_r[2]), t)) {
                            hit_prim = prim_index; hit_tri = 0xFFFFFFFFu;
                            if (OCCLUSION) { out.t = t; out.prim = hit_prim; out.tri = hit_tri; out.v = 0; out.w = 0; return; }
                        }
                    } else if (type == BPT_PRIM_MESH) {
                        const DMesh* mesh = sc.meshes + __ldg(&prim->mesh);
                        const DNodeHalf* bn = sc.blas_nodes + __ldg(&mesh->node_base);
                        if (STATS) { ctr.mesh_calls += 1; c_pops = 1; c_inner = 0; c_leaves = 0; }
                        float4 q0 = __ldg(&bn[0].q0), q1 = __ldg(&bn[0].q1);
                        float tn;
                        if (slab_test(oray, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tn) && (tn < t)) {
                            level = 1; ray = oray; nodes = bn; blas_sp = sp;
                            cur_prim = prim_index; cur_tri_base = __ldg(&mesh->tri_base);
                            cur_lf = __float_as_uint(q1.z); cur_ca = __float_as_uint(q1.w);
                            state = S_NODE;
                        } else if (STATS) {
                            ctr.blas_pops += 1;      // root popped, box missed: the call ends with 1 traversal
                        }
                    }
                }
            }
        }

        if (state == S_POP) {
            if (level == 1 && sp == blas_sp) {
                // intersect_mesh returns: back to the TLAS leaf's item loop in world space
                if (STATS) { ctr.blas_pops += c_pops; ctr.blas_inner += c_inner; ctr.blas_leaves += c_leaves; }
                level = 0; ray = wray; nodes = sc.tlas_nodes;
                state = S_ITEMS;
            } else if (level == 0 && leaf_i < leaf_end) {
                state = S_ITEMS;
            } else if (sp == 0) {
                state = S_DONE;
            } else {
                --sp;
                if (stk_tn[sp] < t) { cur_lf = stk_lf[sp]; cur_ca = stk_ca[sp]; state = S_NODE; }
            }
        }
    }

    out.t = t; out.prim = hit_prim; out.tri = hit_tri; out.v = hit_v; out.w = hit_w;
}

} // namespace bpt
