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
    uint32_t neg;          // bit k = d_is_negative[k]  (intersection.h:13-24); bit 3 = "no NaN/inf can arise in a slab test"
};

// A ray is "tame" when every |d_k| is in [1e-20, 1e20] and every |o_k| < 1e15: then, for node boxes inside 1e15
// (checked at upload, DScene::tame_bounds), inv*(o-p) and |inv|*r are finite and below 1e36, so no slab-test
// intermediate is NaN or inf and the reference's ternary min/max (my_math.h:77-85) agree with FMNMX on every
// comparison outcome (they can differ only in NaN handling and in the sign of a zero, which no comparison sees).
#define BPT_RAY_TAME 8u

BPT_D void make_ray(RayT& r, V3 o, V3 d) {
    r.o = o; r.d = d;
    r.inv = 1.0f / d;
    r.neg = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
    float dlo = fminf(fminf(fabsf(d.x), fabsf(d.y)), fabsf(d.z)), dhi = fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fabsf(d.z));
    float ohi = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
    if (dlo >= 1e-20f && dhi <= 1e20f && ohi < 1e15f) r.neg |= BPT_RAY_TAME;     // false for NaN inputs
}

// same test for tame rays (see make_ray): FMNMX instead of compare+select
BPT_D bool slab_test_tame(const RayT& r, float px, float py, float pz, float rx, float ry, float rz, float& tn) {
    float nx = r.inv.x*(r.o.x - px), ny = r.inv.y*(r.o.y - py), nz = r.inv.z*(r.o.z - pz);
    float kx = fabsf(r.inv.x)*rx,    ky = fabsf(r.inv.y)*ry,    kz = fabsf(r.inv.z)*rz;
    float t1x = -nx - kx, t1y = -ny - ky, t1z = -nz - kz;
    float t2x = -nx + kx, t2y = -ny + ky, t2z = -nz + kz;
    tn = fmaxf(fmaxf(t1x, t1y), t1z);
    float tf = fminf(fminf(t2x, t2y), t2z);
    return (tn < tf) && (tf > 0.0f);
}

// Rays with an exactly-zero direction component (inv_d = +-inf) hit a degenerate case of the reference's slab test:
// inf*0 / inf-inf produce NaNs that the ternary min/max silently drop together with a neighbouring slab, so the test
// passes for almost every node along the ray's projection -- the reference spends 5.9 M node pops (0.37 s on a CPU
// core) on ONE such primary ray of BASELINE config 3.  Every triangle such a ray can actually hit lies (up to float
// rounding) in a box that contains the ray's constant coordinate on that axis, so nodes that miss it by more than a
// generous margin are rejected here on top of the reference test.  The visited nodes become a subsequence of the
// reference's (same order, same tn), every accepted hit of the reference is still found, hence identical hit records;
// only the visit counters of these rays drop below the reference's.
BPT_D bool parallel_axes_may_contain(const RayT& r, float px, float py, float pz, float rx, float ry, float rz) {
    const float inf = __int_as_float(0x7f800000);
    bool ok = true;
    if (fabsf(r.inv.x) == inf) ok = ok && (fabsf(r.o.x - px) <= rx + 1e-3f*(fabsf(r.o.x) + fabsf(px) + fabsf(rx)) + 1e-30f);
    if (fabsf(r.inv.y) == inf) ok = ok && (fabsf(r.o.y - py) <= ry + 1e-3f*(fabsf(r.o.y) + fabsf(py) + fabsf(ry)) + 1e-30f);
    if (fabsf(r.inv.z) == inf) ok = ok && (fabsf(r.o.z - pz) <= rz + 1e-3f*(fabsf(r.o.z) + fabsf(pz) + fabsf(rz)) + 1e-30f);
    return ok;
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

#ifndef BPT_STACK_TOP_IN_REGS
#define BPT_STACK_TOP_IN_REGS 0
#endif
#define BPT_STACK_DEPTH 64     // the reference's node_stack[64] (intersection.cpp:261, :445), far children only here

// Per-lane traversal state.  begin() does what precedes the reference's node loop (planes + TLAS root pop); the
// node / triangle / instance steps are driven by persistent_trace below, which is the ONLY traversal loop in the
// library (render passes and the bpt_trace diagnostic both run it).
struct TraversalStack {
    float4 e[BPT_STACK_DEPTH];     // {left_first, count|axis<<16, entry distance, -}: one STL.128 / LDL.128 per push / pop
};

// MODE: 0 = closest hit (intersect_scene), 1 = occlusion (intersect_shadow_ray), 2 = per ray (Src::load says which)
enum { TRACE_MODE_CLOSEST = 0, TRACE_MODE_OCCLUSION = 1, TRACE_MODE_MIXED = 2 };

template <int MODE, bool STATS>
struct Traversal {
    enum { S_NODE, S_ITEMS, S_POP, S_DONE };

    RayT ray;                            // ray in the current space (world in the TLAS, object inside a BLAS)
    V3 wo, wd, winv;                     // world-space ray, restored when intersect_mesh "returns"
    uint32_t wneg;
    float t;
    uint32_t hit_prim, hit_tri;
    float hit_v, hit_w;
    uint32_t cur_lf, cur_ca;
    uint32_t leaf_i, leaf_end;           // TLAS leaf items still to test
    uint32_t cur_prim, cur_tri_base, ignored;
    bool occ;                            // TRACE_MODE_MIXED: this ray is a shadow ray
    const DNodeHalf* nodes;
    int sp, blas_sp, state, level;       // level: 0 = TLAS, 1 = inside a mesh BLAS
    uint32_t c_pops, c_inner, c_leaves;  // per intersect_mesh call; dropped on an occlusion early-out like g_stats
    // the stack itself lives outside the struct (TraversalStack) so these scalars stay in registers

    BPT_D bool done() const { return state == S_DONE; }

    BPT_D void begin(const DScene& sc, V3 o, V3 d, float max_t, uint32_t ignored_prim, TraceCounters& ctr) {
        wo = o; wd = d;
        make_ray(ray, o, d);
        winv = ray.inv; wneg = ray.neg;
        t = max_t;
        hit_prim = BPT_HIT_MISS; hit_tri = 0xFFFFFFFFu; hit_v = 0.0f; hit_w = 0.0f;
        ignored = ignored_prim;
        sp = 0; blas_sp = 0; level = 0; leaf_i = 0; leaf_end = 0; cur_prim = 0; cur_tri_base = 0;
        c_pops = c_inner = c_leaves = 0;
        nodes = sc.tlas_nodes;
        // planes first, linearly (intersection.cpp:424-433); in occlusion mode a plane hit does not return early
        for (uint32_t i = 0; i < sc.plane_count; ++i) {
            const DPlane& pl = sc.planes[i];
            if (plane_test(ray, v3(__ldg(&pl.n[0]), __ldg(&pl.n[1]), __ldg(&pl.n[2])), __ldg(&pl.d), t)) hit_prim = BPT_HIT_PLANE | i;
        }
        // TLAS root (popped and box-tested like any node, intersection.cpp:450-454)
        float4 q0 = __ldg(&nodes[0].q0), q1 = __ldg(&nodes[0].q1);
        float tn;
        bool hit = slab_test(ray, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tn) && (tn < t);
        if (!(ray.neg & BPT_RAY_TAME)) hit = hit && parallel_axes_may_contain(ray, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
        if (STATS) ctr.tlas_pops += 1;
        cur_lf = __float_as_uint(q1.z); cur_ca = __float_as_uint(q1.w);
        state = hit ? S_NODE : S_DONE;
    }

    BPT_D void result(HitRecord& out) const {
        out.t = t; out.prim = hit_prim; out.tri = hit_tri; out.v = hit_v; out.w = hit_w;
    }
};

// Persistent-warp traversal with lane refill and warp-level phase scheduling.
//
// Measured problem with a per-thread state machine (ncu, profiles/r1_trace_v1.txt): for incoherent rays only 3-8 of 32
// lanes were active per issued instruction, because at any moment the lanes of a warp sit in different phases
// (slab-testing a sibling pair / testing a triangle / transforming the ray into an instance / idle) and the warp
// serialises over all of them every iteration, and because finished lanes idle until the slowest ray of the warp ends.
// Here each lane still owns one ray (its Traversal stays in registers) but per iteration the WARP executes only
// the phase that most of its lanes are waiting for -- ballot + popc pick it -- and "idle" is one of the phases: when
// idle lanes are the largest group (and rays remain) they fetch new rays with one warp-aggregated atomicAdd and run
// begin().  Every ray still performs exactly the reference's sequence of tests; only the interleaving between
// independent rays changes.   Src supplies load(i, o, d, max_t, ignored, occ) / store(i, hit).
// LOCAL = true: there is no shared ray queue; every lane produces its own rays -- Src supplies pending() and
// bool next(o, d, max_t, ignored, occ) instead of load(); next() may do arbitrary per-lane work (k_tail shades the
// lane's path there) -- and the call returns when no lane has anything pending.
template <int MODE, bool STATS, bool LOCAL, class Src>
BPT_D void persistent_trace(const DScene& sc, Src& src, uint32_t n, uint32_t* cursor, uint32_t refill, TraceCounters& ctr) {
    typedef Traversal<MODE, STATS> TV;
    enum { P_IDLE = 0, P_INNER = 1, P_TRI = 2, P_ITEMS = 3 };
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t lane = threadIdx.x & 31u;
    TV tv;
    TraversalStack stk;
    tv.state = TV::S_DONE;
    int phase = P_IDLE;
    uint32_t tri_k = 0;
    bool exhausted = false;
    uint32_t my_index = 0;
    const bool tame = sc.tame_bounds != 0;
#if BPT_STACK_TOP_IN_REGS
    float4 top = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    bool top_valid = false;          // entry tv.sp-1 is in `top`, entries below it are in stk
#endif
    tv.occ = false;
    auto occlusion = [&]() { return MODE == TRACE_MODE_MIXED ? tv.occ : (MODE == TRACE_MODE_OCCLUSION); };

    // after `tv.cur_*` changed: which phase does the lane wait for now
    auto classify = [&]() {
        uint32_t count = tv.cur_ca & 0xFFFFu;
        if (count == 0) phase = P_INNER;
        else if (tv.level == 1) { phase = P_TRI; tri_k = 0; if (STATS) { tv.c_leaves += 1; ctr.tris += count; } }
        else { tv.leaf_i = tv.cur_lf; tv.leaf_end = tv.cur_lf + count; phase = P_ITEMS; }
    };
    auto finish = [&]() {
        HitRecord h; tv.result(h); src.store(my_index, h);
        phase = P_IDLE;
    };
    // the reference's "pop until a node survives its box re-test" (intersection.cpp:269-277, :450-454)
    auto pop = [&]() {
        for (;;) {
            if (tv.level == 1 && tv.sp == tv.blas_sp) {
                if (STATS) { ctr.blas_pops += tv.c_pops; ctr.blas_inner += tv.c_inner; ctr.blas_leaves += tv.c_leaves; }
                tv.level = 0; tv.nodes = sc.tlas_nodes;
                tv.ray.o = tv.wo; tv.ray.d = tv.wd; tv.ray.inv = tv.winv; tv.ray.neg = tv.wneg;   // == make_ray(wo, wd) again
                phase = P_ITEMS;            // intersect_mesh returned: continue the TLAS leaf's item loop
                return;
            }
            if (tv.sp == 0) { finish(); return; }
            --tv.sp;
#if BPT_STACK_TOP_IN_REGS
            // the newest entry lives in registers: a push that is popped before the next push never touches memory
            float4 e;
            if (top_valid) { e = top; top_valid = false; }
            else e = stk.e[tv.sp];
#else
            float4 e = stk.e[tv.sp];
#endif
            if (e.z < tv.t) { tv.cur_lf = __float_as_uint(e.x); tv.cur_ca = __float_as_uint(e.y); classify(); return; }
        }
    };

    // The scheduling loop.  Its trip count is bounded on purpose: with a plain `for (;;)` whose only exit is the vote
    // below, nvcc 12.9 rotates the loop and peels its first iteration, and the resulting k_trace_merged hung on
    // B200 (deterministically, BASELINE config 4 at >= 16 spp: launch of bounce 3 never returned; every
    // instrumented build ran through, and so did the same PTX assembled with ptxas -O0; -O1 and above hang).  A second,
    // never-taken exit at the loop head keeps the loop in its source shape; tests/test_gpu_golden_and_fullsize.py::test_no_hang_* pins the behaviour.
#ifdef BPT_DBG_PLAIN_LOOP            /* reproduces the hang described above; for investigating it only */
    for (;;) {
#else
    for (unsigned long long trips = 0; trips != ~0ull; ++trips) {
#endif
        // phase populations of the warp, one byte each, from a single warp-wide add (REDUX.SUM)
        // (an idle lane that cannot get another ray votes for nothing)
        bool can_fetch;
        if constexpr (LOCAL) can_fetch = src.pending(); else can_fetch = !exhausted;
        uint32_t counts = __reduce_add_sync(FULL, (phase == P_IDLE && !can_fetch) ? 0u : (1u << (8*phase)));
        if (counts == 0u) break;
        uint32_t n_inner = (counts >> 8) & 0xFFu, n_tri = (counts >> 16) & 0xFFu, n_items = counts >> 24;
        uint32_t n_idle = counts & 0xFFu;

        int run = P_INNER; uint32_t best = n_inner;
        if (n_tri > best)   { best = n_tri;   run = P_TRI; }
        if (n_items > best) { best = n_items; run = P_ITEMS; }
        if (n_idle >= refill || n_idle > best) { run = P_IDLE; }

        if (run == P_IDLE) {
          if constexpr (LOCAL) {
            if (phase == P_IDLE && can_fetch) {
                V3 o, d; float max_t; uint32_t ign;
                bool is_occ = false;
                if (src.next(o, d, max_t, ign, is_occ)) {       // false: the lane's path ended without another ray
                    tv.occ = is_occ;
                    tv.begin(sc, o, d, max_t, ign, ctr);
#if BPT_STACK_TOP_IN_REGS
                    top_valid = false;
#endif
                    if (tv.done()) finish(); else classify();
                }
            }
          } else {
            uint32_t idle_mask = __ballot_sync(FULL, phase == P_IDLE);
            uint32_t nid = __popc(idle_mask);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(cursor, nid);
            base = __shfl_sync(FULL, base, 0);
            if (phase == P_IDLE) {
                uint32_t idx = base + __popc(idle_mask & ((1u << lane) - 1u));
                if (idx < n) {
                    V3 o, d; float max_t; uint32_t ign;
                    bool is_occ = false;
                    src.load(idx, o, d, max_t, ign, is_occ);
                    tv.occ = is_occ;
                    my_index = idx;
                    tv.begin(sc, o, d, max_t, ign, ctr);
#if BPT_STACK_TOP_IN_REGS
                    top_valid = false;
#endif
                    if (tv.done()) finish(); else classify();
                }
            }
            if (base + nid >= n) exhausted = true;
          }
        } else if (run == P_INNER) {
          // stay in this phase while at least 3/4 of the lanes that started it still want it (1 vote instead of 4)
          uint32_t keep = max(best - (best >> 2), 1u);
          do {
            if (phase == P_INNER) {
                // the parent's split axis and the ray's sign say which child is near (intersection.cpp:303-318):
                // fetch them in that order so nothing has to be swapped afterwards
                uint32_t right_first = (tv.ray.neg >> (tv.cur_ca >> 16)) & 1u;
                const DNodeHalf* pn = tv.nodes + (tv.cur_lf + right_first);
                const DNodeHalf* pf = tv.nodes + (tv.cur_lf + (right_first ^ 1u));
                float4 n0 = __ldg(&pn->q0), n1 = __ldg(&pn->q1);
                float4 f0 = __ldg(&pf->q0), f1 = __ldg(&pf->q1);
                float near_tn, far_tn;
                bool near_hit, far_hit;
                if (tame && (tv.ray.neg & BPT_RAY_TAME)) {
                    near_hit = slab_test_tame(tv.ray, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, near_tn);
                    far_hit  = slab_test_tame(tv.ray, f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, far_tn);
                } else {
                    near_hit = slab_test(tv.ray, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, near_tn);
                    far_hit  = slab_test(tv.ray, f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, far_tn);
                    near_hit = near_hit && parallel_axes_may_contain(tv.ray, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y);
                    far_hit  = far_hit  && parallel_axes_may_contain(tv.ray, f0.x, f0.y, f0.z, f0.w, f1.x, f1.y);
                }
                if (STATS) { if (tv.level) { tv.c_pops += 2; tv.c_inner += 1; } else ctr.tlas_pops += 2; }
                uint32_t near_lf = __float_as_uint(n1.z), near_ca = __float_as_uint(n1.w);
                if (far_hit && tv.sp < BPT_STACK_DEPTH) {
#if BPT_STACK_TOP_IN_REGS
                    if (top_valid) stk.e[tv.sp - 1] = top;              // the previous newest entry moves to memory
                    top = make_float4(f1.z, f1.w, far_tn, 0.0f); top_valid = true; ++tv.sp;
#else
                    stk.e[tv.sp] = make_float4(f1.z, f1.w, far_tn, 0.0f); ++tv.sp;
#endif
                }
                if (near_hit && near_tn < tv.t) { tv.cur_lf = near_lf; tv.cur_ca = near_ca; classify(); }
                else pop();
            }
          } while (__popc(__ballot_sync(FULL, phase == P_INNER)) >= keep);
        } else if (run == P_TRI) {
          uint32_t keep = max(best - (best >> 2), 1u);
          do {
            if (phase == P_TRI) {
                uint32_t slot = tv.cur_tri_base + tv.cur_lf + tri_k;
                const DTriangle* tri = sc.triangles + slot;
                float4 a = __ldg(&tri->a_idx), e1 = __ldg(&tri->e1), e2 = __ldg(&tri->e2);
                bool stop = false;
                if (triangle_test(tv.ray, v3(a), v3(e1), v3(e2), tv.t, tv.hit_v, tv.hit_w)) {
                    tv.hit_tri = slot;
                    tv.hit_prim = tv.cur_prim;
                    if (occlusion()) { finish(); stop = true; }
                }
                if (!stop && ++tri_k >= (tv.cur_ca & 0xFFFFu)) pop();
            }
          } while (__popc(__ballot_sync(FULL, phase == P_TRI)) >= keep);
        } else {
            if (phase == P_ITEMS) {
                if (tv.leaf_i >= tv.leaf_end) {
                    pop();
                } else {
                    uint32_t prim_index = __ldg(&sc.tlas_indices[tv.leaf_i++]);
                    if (prim_index != tv.ignored) {
                        const DPrimitive* prim = sc.primitives + prim_index;
                        float4 m[3] = {__ldg(&prim->inv[0]), __ldg(&prim->inv[1]), __ldg(&prim->inv[2])};
                        RayT oray;
                        oray.o = xform(m, tv.wo, 1.0f); oray.d = xform(m, tv.wd, 0.0f);     // transform_ray :403-409
                        if (STATS) ctr.instances += 1;
                        uint32_t type = __ldg(&prim->type);
                        if (type != BPT_PRIM_SPHERE) make_ray(oray, oray.o, oray.d);        // the sphere test never reads inv_d
                        if (type == BPT_PRIM_SPHERE) {
                            if (sphere_test(oray, __ldg(&prim->sphere_r), tv.t)) {
                                tv.hit_prim = prim_index; tv.hit_tri = 0xFFFFFFFFu;
                                if (occlusion()) finish();
                            }
                        } else if (type == BPT_PRIM_BOX) {
                            if (box_test(oray, __ldg(&prim->box_r[0]), __ldg(&prim->box_r[1]), __ldg(&prim->box_r[2]), tv.t)) {
                                tv.hit_prim = prim_index; tv.hit_tri = 0xFFFFFFFFu;
                                if (occlusion()) finish();
                            }
                        } else if (type == BPT_PRIM_MESH) {
                            const DMesh* mesh = sc.meshes + __ldg(&prim->mesh);
                            const DNodeHalf* bn = sc.blas_nodes + __ldg(&mesh->node_base);
                            if (STATS) { ctr.mesh_calls += 1; tv.c_pops = 1; tv.c_inner = 0; tv.c_leaves = 0; }
                            float4 q0 = __ldg(&bn[0].q0), q1 = __ldg(&bn[0].q1);
                            float tn;
                            bool root_hit = slab_test(oray, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tn) && (tn < tv.t);
                            if (!(oray.neg & BPT_RAY_TAME)) root_hit = root_hit && parallel_axes_may_contain(oray, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
                            if (root_hit) {
                                tv.level = 1; tv.ray = oray; tv.nodes = bn; tv.blas_sp = tv.sp;
                                tv.cur_prim = prim_index; tv.cur_tri_base = __ldg(&mesh->tri_base);
                                tv.cur_lf = __float_as_uint(q1.z); tv.cur_ca = __float_as_uint(q1.w);
                                classify();
                            } else if (STATS) {
                                ctr.blas_pops += 1;
                            }
                        }
                    }
                }
            }
        }
    }
}

} // namespace bpt
