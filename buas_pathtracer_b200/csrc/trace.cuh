// Two-level BVH traversal + primitive tests, equivalent hit-for-hit to the reference's
// intersect_scene_internal / intersect_mesh (Raytracer/intersection.cpp:12-182, :243-401, :411-520).
//
// Re-designed for SIMT rather than transcribed:
//   * one flattened state machine walks TLAS and BLAS with a single per-thread stack (the reference nests
//     intersect_mesh's loop inside the TLAS leaf loop, which would serialise a warp);
//   * nodes are read as 64-byte sibling pairs grouped into two-level records (wide_bvh.h): entering a record root the
//     lane fetches the node's child pair AND the near child's pair with eight 128-bit loads issued together, slab-tests
//     all four boxes and replays the binary visit order in registers -- one dependent memory round trip per two levels;
//   * the traversal stack holds 8-byte entries {child ref, entry distance}; its newest 8 entries live in shared memory
//     (a window that follows the stack top), older ones spill to local memory only when the window overflows.
// Equivalence argument (tie semantics): the reference pushes far then near and
// re-tests each box when popped, against the t of that moment.  (tn < tf && tf > 0) does not depend on t, so
// it is evaluated once at the parent; `tn < t` is evaluated for the near child immediately (nothing happens
// between the reference's push and pop of it) and for the far child when it is popped, with the stored tn.  The second
// level of a record step is exactly the step the reference performs next (the near child was just accepted and t has
// not changed), executed from registers.  Leaves are therefore visited in the reference's order with the reference's
// culling, so equal-t ties (`t <= out_t` accepts, intersection.cpp:174) resolve to the same triangle.
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

#define BPT_STACK_DEPTH 64        // the reference's node_stack[64] (intersection.cpp:261, :445); far children only here
#ifndef BPT_SSTACK
#define BPT_SSTACK 8              // the first BPT_SSTACK stack entries of a thread live in shared memory, deeper ones in local memory
#endif
#ifndef BPT_TWO_LEVEL
#define BPT_TWO_LEVEL 0           // 1: record roots fetch the near child's pair together with their own and take two levels per step
#endif                            //    (measured on B200, C2: traversal 51.2 ms against 44.9 ms with one level per step -- see DESIGN.md)
#ifndef BPT_KEEP_EIGHTHS
#define BPT_KEEP_EIGHTHS 2        // the node phase keeps running while at least this many eighths of the lanes that started it still want it
#endif                            // (C2 traversal ms at 0..6 eighths: 45.5 / 42.3 / 41.2 / 41.5 / 42.1 / 43.0 / 43.7; round 1 used 6)
#ifndef BPT_KEEP_TRI_EIGHTHS
#define BPT_KEEP_TRI_EIGHTHS 0    // the triangle phase runs until no lane wants it (C2 traversal ms at 0 / 1 / 2 / 4 / 6 eighths: 41.1 / 41.3 / 41.4 / 42.1 / 43.1)
#endif
#ifndef BPT_TRI_BIAS
#define BPT_TRI_BIAS 4            // phase selection weighs the lanes waiting for triangles by this factor: a leaf is at most two short steps and its lanes
                                  // come back to the node phase, whose long runs then start fuller (C2 traversal ms at 1 / 2 / 3 / 4 / 8 / 32: 40.5 / 39.3 /
                                  // 38.9 / 38.8 / 38.7 / 38.7; C3 best at 3-4).  The same weight on TLAS items (BPT_ITEMS_BIAS) loses.
#endif
#ifndef BPT_ITEMS_BIAS
#define BPT_ITEMS_BIAS 1
#endif
#ifndef BPT_KEEP_ITEMS_EIGHTHS
#define BPT_KEEP_ITEMS_EIGHTHS 2  // the TLAS-item / return phase keeps running like the other phases when the TLAS has inner nodes (C3 / C4:
#endif                            // -3 % traversal time); with a TLAS that is a single leaf (C2) it runs one step per vote (the loop costs +0.8 % there)
#ifndef BPT_COLD_LOCAL
#define BPT_COLD_LOCAL 0          // 1: the per-ray state outside the inner loop lives in local memory instead of shared memory
#endif
#define BPT_TRACE_THREADS 128     // block size of every kernel that runs persistent_trace
#ifndef BPT_TRIP_LIMIT
#define BPT_TRIP_LIMIT (1u << 28) // scheduling-loop trips per warp; ~200x the largest legitimate count (a 128 Mi-slot batch)
#endif

// MODE: 0 = closest hit (intersect_scene), 1 = occlusion (intersect_shadow_ray), 2 = per ray (Src::load says which)
enum { TRACE_MODE_CLOSEST = 0, TRACE_MODE_OCCLUSION = 1, TRACE_MODE_MIXED = 2 };

// Per-CTA shared memory of a traversal kernel; [..][thread] so that a warp's accesses are conflict-free whatever each
// lane's stack height is.  Besides the short stack it holds the per-ray state that the inner loop does not touch: that
// state would otherwise occupy registers the two-level node step needs for its loads.
struct TraceShared {
    uint2    stack[BPT_SSTACK][BPT_TRACE_THREADS];   // {child ref, entry distance} of stack positions 0 .. BPT_SSTACK-1
#if !BPT_COLD_LOCAL
    float    wray[6][BPT_TRACE_THREADS];             // the world-space ray (o, d) while the lane is inside a mesh BLAS
    uint32_t items_first[BPT_TRACE_THREADS];         // rest of the TLAS leaf's item loop, resumed when intersect_mesh "returns"
    uint32_t items_count[BPT_TRACE_THREADS];
    uint32_t ignored[BPT_TRACE_THREADS];             // ignored_primitive_index of the lane's ray
    uint32_t index[BPT_TRACE_THREADS];               // which ray of the source the lane is tracing
    uint32_t hit_prim[BPT_TRACE_THREADS];            // the hit record so far (t itself stays in a register: every pop reads it)
    uint32_t hit_tri[BPT_TRACE_THREADS];
    float    hit_v[BPT_TRACE_THREADS], hit_w[BPT_TRACE_THREADS];
    uint32_t cur_prim[BPT_TRACE_THREADS];            // the instance whose BLAS the lane is in, and that mesh's first triangle
    uint32_t tri_base[BPT_TRACE_THREADS];
#endif
};

// The same per-ray state as a per-thread record in LOCAL memory (-DBPT_COLD_LOCAL=1): touched a few times per ray, it then
// occupies L1 lines only while in use instead of 8 KB of shared memory per CTA taken from the L1 carve-out for good.
struct TraceCold {
    float    wray[6];
    uint32_t items_first, items_count, ignored, index, hit_prim, hit_tri;
    float    hit_v, hit_w;
    uint32_t cur_prim, tri_base;
};

// Persistent-warp traversal with lane refill and warp-level phase scheduling.
//
// Measured problem with a per-thread state machine (ncu, profiles/r1_trace_v1.txt): for incoherent rays only 3-8 of 32
// lanes were active per issued instruction, because at any moment the lanes of a warp sit in different phases
// (slab-testing node pairs / testing triangles / transforming the ray into an instance / idle) and the warp
// serialises over all of them every iteration, and because finished lanes idle until the slowest ray of the warp ends.
// Here each lane still owns one ray (its traversal state stays in registers / shared memory) but per iteration the WARP
// executes only the phase that most of its lanes are waiting for -- one REDUX picks it -- and "idle" is one of the phases:
// when idle lanes are the largest group (and rays remain) they fetch new rays with one warp-aggregated atomicAdd and run
// begin().  Every ray still performs exactly the reference's sequence of tests; only the interleaving between
// independent rays changes.   Src supplies load(i, o, d, max_t, ignored, occ) / store(i, hit).
// LOCAL = true: there is no shared ray queue; every lane produces its own rays -- Src supplies pending() and
// bool next(o, d, max_t, ignored, occ) instead of load(); next() may do arbitrary per-lane work (k_tail shades the
// lane's path there) -- and the call returns when no lane has anything pending.
// This is the ONLY traversal loop in the library (render passes, the recursive integrators and the bpt_trace diagnostic
// all run it).  Must be called by all BPT_TRACE_THREADS threads of the block, whole warps converged.
template <int MODE, bool STATS, bool LOCAL, class Src>
BPT_D void persistent_trace(const DScene& sc, Src& src, uint32_t n, uint32_t* cursor, uint32_t refill, TraceCounters& ctr) {
    // P_RET: the lane's stack is down to its floor -- intersect_mesh returns (back to the TLAS leaf's item loop) or the
    // ray is finished.  It votes with, and is handled inside, the P_ITEMS step so that the node / triangle steps carry
    // nothing but their own work (their divergent tails were 25 % of the instructions at 8 of 32 lanes, ncu r2_trace_v7).
    enum { P_IDLE = 0, P_INNER = 1, P_TRI = 2, P_ITEMS = 3, P_RET = 4 };
    __shared__ TraceShared sh;
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
#if BPT_COLD_LOCAL
    TraceCold cold_storage;
    TraceCold* cold = &cold_storage;
    asm volatile("" : "+l"(cold));              // opaque to the optimiser: the record stays in local memory, off the registers
    #define COLD(f) (cold->f)
    #define COLDW(k) (cold->wray[k])
#else
    #define COLD(f) (sh.f[tid])
    #define COLDW(k) (sh.wray[k][tid])
#endif

    RayT ray;                                   // the ray in the current space (world in the TLAS, object inside a BLAS)
    ray.o = v3(0.0f); ray.d = v3(0.0f); ray.inv = v3(0.0f); ray.neg = 0u;
    float t = 0.0f;
    uint32_t cur_a = 0u, cur_b = 0u;            // P_INNER: cur_a = ref of the inner node being entered;
                                                // P_TRI / P_ITEMS: cur_a = next triangle / TLAS item of the leaf, cur_b = how many are left
    uint32_t pair_base = 0u;
    int sp = 0, blas_sp = -1;                   // stack height; height at which the current BLAS was entered (-1: in the TLAS)
    bool occ = false;
    uint32_t c_pops = 0u, c_inner = 0u, c_leaves = 0u;   // per intersect_mesh call; dropped on an occlusion early-out like g_stats
    uint2 lstack[BPT_STACK_DEPTH - BPT_SSTACK]; // stack positions BPT_SSTACK .. 63 (local memory)
    int phase = P_IDLE;
    bool exhausted = false;
    const bool tame = sc.tame_bounds != 0;
    const bool tlas_flat = (__float_as_uint(sc.tlas_root_q1.z) & BPT_WREF_LEAF) != 0u;     // the whole TLAS is one leaf
    auto occlusion = [&]() { return MODE == TRACE_MODE_MIXED ? occ : (MODE == TRACE_MODE_OCCLUSION); };

    auto push = [&](uint32_t ref, float tn) {
        uint2 e = make_uint2(ref, __float_as_uint(tn));
        if (sp < BPT_SSTACK) sh.stack[sp][tid] = e;
        else if (sp < BPT_STACK_DEPTH) lstack[sp - BPT_SSTACK] = e;
        else { *sc.error_flag = BPT_DEVERR_STACK_OVERFLOW; return; }     // the reference overruns node_stack[64] here
        ++sp;
    };
    // a node has been accepted (its box test passed against the current t): which phase does the lane wait for now
    auto enter = [&](uint32_t ref) {
        if (!(ref & BPT_WREF_LEAF)) { cur_a = ref; phase = P_INNER; return; }
        uint32_t count = (ref >> 28) & 7u, first = ref & BPT_WREF_INDEX_MASK;
        if (count == 0u) {                      // forced leaf with more than 7 items (bvh.cpp:254, :278): range is in the side table
            uint32_t bb = blas_sp >= 0 ? __ldg(&sc.meshes[__ldg(&sc.primitives[COLD(cur_prim)].mesh)].big_base) : 0u;
            uint2 bl = __ldg(&sc.big_leaves[bb + first]);
            first = bl.x; count = bl.y;
        }
        cur_a = first; cur_b = count;
        if (blas_sp >= 0) { phase = P_TRI; if (STATS) { c_leaves += 1u; ctr.tris += count; } }
        else phase = P_ITEMS;
    };
    auto finish = [&]() {
        HitRecord h; h.t = t; h.prim = COLD(hit_prim); h.tri = COLD(hit_tri); h.v = COLD(hit_v); h.w = COLD(hit_w);
        src.store(COLD(index), h);
        phase = P_IDLE;
    };
    // the reference's "pop until a node survives its box re-test" (intersection.cpp:269-277, :450-454); when the stack is
    // down to the floor of the current traversal the lane goes to P_RET
    auto pop_enter = [&]() {
        const int floor = blas_sp > 0 ? blas_sp : 0;
        while (sp > floor) {
            --sp;
            uint2 e;
            if (sp < BPT_SSTACK) e = sh.stack[sp][tid]; else e = lstack[sp - BPT_SSTACK];
            if (__uint_as_float(e.y) < t) { enter(e.x); return; }
        }
        phase = P_RET;
    };
    // ray_intersect_bounding_volume minus its `tn < t` clause, for any ray
    auto slab_any = [&](const RayT& r, const float4& q0, const float4& q1, float& tn) {
        bool h = slab_test(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tn);
        if (!(r.neg & BPT_RAY_TAME)) h = h && parallel_axes_may_contain(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
        return h;
    };
    // what precedes the reference's node loop: planes, then the TLAS root is popped and box-tested like any node
    auto begin = [&](V3 o, V3 d, float max_t, uint32_t ignored_prim) {
        make_ray(ray, o, d);
        if (!tame) ray.neg &= ~BPT_RAY_TAME;    // node boxes beyond 1e15: every ray takes the exact compare+select slab test
        t = max_t;
        uint32_t hp = BPT_HIT_MISS;
        COLD(ignored) = ignored_prim;
        sp = 0; blas_sp = -1; pair_base = 0u;
        // planes first, linearly (intersection.cpp:424-433); in occlusion mode a plane hit does not return early
        for (uint32_t i = 0; i < sc.plane_count; ++i) {
            const DPlane& pl = sc.planes[i];
            if (plane_test(ray, v3(__ldg(&pl.n[0]), __ldg(&pl.n[1]), __ldg(&pl.n[2])), __ldg(&pl.d), t)) hp = BPT_HIT_PLANE | i;
        }
        COLD(hit_prim) = hp; COLD(hit_tri) = 0xFFFFFFFFu; COLD(hit_v) = 0.0f; COLD(hit_w) = 0.0f;
        float tn;
        bool hit = slab_any(ray, sc.tlas_root_q0, sc.tlas_root_q1, tn) && (tn < t);      // intersection.cpp:450-454
        if (STATS) ctr.tlas_pops += 1u;
        if (hit) enter(__float_as_uint(sc.tlas_root_q1.z)); else finish();
    };

    // The scheduling loop.  Its trip count is bounded: with a plain `for (;;)` whose only exit is the vote below, nvcc
    // 12.9 rotated the loop and peeled its first iteration, and that build of k_trace_merged hung on B200
    // (deterministically, BASELINE config 4 at >= 16 spp; every instrumented build ran through, and so did the same PTX
    // assembled with ptxas -O0; -O1 and above hung).  An exit at the loop head keeps the loop in its source shape --
    // tests/test_gpu_golden_and_fullsize.py::test_no_hang_* pins that -- and the bound is a real one: a warp that
    // exceeds it raises BPT_DEVERR_TRIP_LIMIT and leaves, so a scheduling bug ends as an error code, not a hung GPU.
    uint32_t trips = 0;
#ifdef BPT_DBG_PLAIN_LOOP            /* the shape that hung; for investigating it only */
    for (;;) {
#else
    for (; trips != BPT_TRIP_LIMIT; ++trips) {
#endif
        // phase populations of the warp, one byte each, from a single warp-wide add (REDUX.SUM)
        // (an idle lane that cannot get another ray votes for nothing; P_RET votes with P_ITEMS)
        bool can_fetch;
        if constexpr (LOCAL) can_fetch = src.pending(); else can_fetch = !exhausted;
        uint32_t counts = __reduce_add_sync(FULL, (phase == P_IDLE && !can_fetch) ? 0u : (1u << (8*min(phase, (int)P_ITEMS))));
        if (counts == 0u) break;
        uint32_t n_inner = (counts >> 8) & 0xFFu, n_tri = (counts >> 16) & 0xFFu, n_items = counts >> 24;
        uint32_t n_idle = counts & 0xFFu;

        int run = P_INNER; uint32_t best = n_inner, score = n_inner;
        if (n_tri*BPT_TRI_BIAS > score)     { best = n_tri;   score = n_tri*BPT_TRI_BIAS;     run = P_TRI; }
        if (n_items*BPT_ITEMS_BIAS > score) { best = n_items; score = n_items*BPT_ITEMS_BIAS; run = P_ITEMS; }
        if (n_idle >= refill || n_idle > score) { run = P_IDLE; }

        if (run == P_IDLE) {
          if constexpr (LOCAL) {
            if (phase == P_IDLE && can_fetch) {
                V3 o, d; float max_t; uint32_t ign;
                bool is_occ = false;
                if (src.next(o, d, max_t, ign, is_occ)) {       // false: the lane's path ended without another ray
                    occ = is_occ;
                    begin(o, d, max_t, ign);
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
                    occ = is_occ;
                    COLD(index) = idx;
                    begin(o, d, max_t, ign);
                }
            }
            if (base + nid >= n) exhausted = true;
          }
        } else if (run == P_INNER) {
          // stay in this phase while at least BPT_KEEP_EIGHTHS/8 of the lanes that started it still want it (one ballot per
          // step instead of the full vote; measured on C2: 6/8 -> 43.7 ms of traversal, 2/8 -> 41.2)
          uint32_t keep = max((best*BPT_KEEP_EIGHTHS + 7u) >> 3, 1u);
          do {
            if (phase == P_INNER) {
                // the node's split axis and the ray's sign say which child is near (intersection.cpp:303-318): known
                // before anything is loaded, so the children are fetched in near/far order
                const uint32_t ref = cur_a;
                const uint32_t k = (ray.neg >> (ref >> 29)) & 1u;                   // bit 31 is clear: ref >> 29 is the axis
                const DPair* pp = sc.pairs + (pair_base + (ref & BPT_WREF_INDEX_MASK));
                const DChild* pn = &pp->c[k];
                const DChild* pf = &pp->c[k ^ 1u];
                float4 n0 = __ldg(&pn->q0), n1 = __ldg(&pn->q1);
                float4 f0 = __ldg(&pf->q0), f1 = __ldg(&pf->q1);
#if BPT_TWO_LEVEL
                // at a record root the near child's own pair sits right behind the node's pair: fetch it along
                const bool two = (ref & BPT_WREF_RECORD_ROOT) != 0u && (ray.neg & BPT_RAY_TAME);
                float4 a0, a1, b0, b1;
                if (two) {
                    const DPair* pg = pp + 1 + k;
                    a0 = __ldg(&pg->c[0].q0); a1 = __ldg(&pg->c[0].q1);
                    b0 = __ldg(&pg->c[1].q0); b1 = __ldg(&pg->c[1].q1);
                }
#endif
                float near_tn, far_tn;
                bool near_hit, far_hit;
                if (ray.neg & BPT_RAY_TAME) {
                    near_hit = slab_test_tame(ray, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, near_tn);
                    far_hit  = slab_test_tame(ray, f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, far_tn);
                } else {
                    // rays with a zero / denormal / huge direction component: the exact compare+select slab test
                    near_hit = slab_any(ray, n0, n1, near_tn);
                    far_hit  = slab_any(ray, f0, f1, far_tn);
                }
                if (STATS) { if (blas_sp >= 0) { c_pops += 2u; c_inner += 1u; } else ctr.tlas_pops += 2u; }
                // The reference pushes far then near and pops: near is re-tested against the unchanged t; when it fails, far
                // is popped next and re-tested against the same t.  So: near accepted -> far (if its box is hit) waits on
                // the stack; near rejected -> far is entered directly if it would pass, and is never pushed otherwise.
                const uint32_t nref = __float_as_uint(n1.z), fref = __float_as_uint(f1.z);
                const bool go_near = near_hit && near_tn < t;
                const bool go_far = !go_near && far_hit && far_tn < t;
                if (go_near && far_hit) push(fref, far_tn);
#if BPT_TWO_LEVEL
                if (go_near && two && !(nref & BPT_WREF_LEAF)) {
                    // the step the reference performs next, from registers: the near child was accepted, t is unchanged
                    float a_tn, b_tn;
                    bool a_hit = slab_test_tame(ray, a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a_tn);
                    bool b_hit = slab_test_tame(ray, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b_tn);
                    if (STATS) { if (blas_sp >= 0) { c_pops += 2u; c_inner += 1u; } else ctr.tlas_pops += 2u; }
                    const bool right_first = ((ray.neg >> (nref >> 29)) & 1u) != 0u;
                    bool     h_near = right_first ? b_hit : a_hit,   h_far = right_first ? a_hit : b_hit;
                    float    t_near = right_first ? b_tn : a_tn,     t_far = right_first ? a_tn : b_tn;
                    uint32_t r_near = __float_as_uint(right_first ? b1.z : a1.z), r_far = __float_as_uint(right_first ? a1.z : b1.z);
                    const bool go2_near = h_near && t_near < t;
                    const bool go2_far = !go2_near && h_far && t_far < t;
                    if (go2_near && h_far) push(r_far, t_far);
                    if (go2_near || go2_far) enter(go2_near ? r_near : r_far); else pop_enter();
                } else
#endif
                if (go_near || go_far) enter(go_near ? nref : fref);
                else pop_enter();
            }
          } while (__popc(__ballot_sync(FULL, phase == P_INNER)) >= keep);
        } else if (run == P_TRI) {
          uint32_t keep = max((best*BPT_KEEP_TRI_EIGHTHS + 7u) >> 3, 1u);
          do {
            if (phase == P_TRI) {
                // up to two triangles of the leaf per step, fetched together, tested in the reference's order
                if (cur_b == 0u) {
                    pop_enter();
                } else {
                    const uint32_t slot = COLD(tri_base) + cur_a;
                    const DTriangle* tri = sc.triangles + slot;
                    const bool second = cur_b > 1u;
                    float4 a = __ldg(&tri->a_idx), e1 = __ldg(&tri->e1), e2 = __ldg(&tri->e2);
                    float4 a2 = a, e12 = e1, e22 = e2;
                    if (second) { a2 = __ldg(&tri[1].a_idx); e12 = __ldg(&tri[1].e1); e22 = __ldg(&tri[1].e2); }
                    bool stop = false;
                    float v, w;
                    if (triangle_test(ray, v3(a), v3(e1), v3(e2), t, v, w)) {
                        COLD(hit_tri) = slot; COLD(hit_prim) = COLD(cur_prim); COLD(hit_v) = v; COLD(hit_w) = w;
                        if (occlusion()) { finish(); stop = true; }
                    }
                    if (!stop && second) {
                        if (triangle_test(ray, v3(a2), v3(e12), v3(e22), t, v, w)) {
                            COLD(hit_tri) = slot + 1u; COLD(hit_prim) = COLD(cur_prim); COLD(hit_v) = v; COLD(hit_w) = w;
                            if (occlusion()) { finish(); stop = true; }
                        }
                    }
                    if (!stop) {
                        uint32_t adv = second ? 2u : 1u;
                        cur_a += adv; cur_b -= adv;
                        if (cur_b == 0u) pop_enter();
                    }
                }
            }
          } while (__popc(__ballot_sync(FULL, phase == P_TRI)) >= keep);
        } else {
          uint32_t keep = tlas_flat ? 33u : max((best*BPT_KEEP_ITEMS_EIGHTHS + 7u) >> 3, 1u);
          do {
            if (phase == P_RET) {
                if (blas_sp >= 0) {
                    // intersect_mesh returns: back to the world-space ray and the TLAS leaf's item loop
                    if (STATS) { ctr.blas_pops += c_pops; ctr.blas_inner += c_inner; ctr.blas_leaves += c_leaves; }
                    blas_sp = -1; pair_base = 0u;
                    // the world ray again: same values from the same operations.  (Keeping 1/d and the sign bits in shared
                    // memory instead of re-dividing was measured: 44.3 against 43.6 ms of traversal on C2 -- the 2 KB more of
                    // shared memory per CTA cost more L1 than the three divisions cost issue slots.)
                    make_ray(ray, v3(COLDW(0), COLDW(1), COLDW(2)),
                                  v3(COLDW(3), COLDW(4), COLDW(5)));
                    if (!tame) ray.neg &= ~BPT_RAY_TAME;
                    cur_a = COLD(items_first); cur_b = COLD(items_count);
                    phase = P_ITEMS;
                } else {
                    finish();                   // the TLAS stack is empty: intersect_scene_internal returns
                }
            }
            if (phase == P_ITEMS) {
                if (cur_b != 0u) {
                    uint32_t prim_index = __ldg(&sc.tlas_indices[cur_a]);
                    ++cur_a; --cur_b;
                    if (prim_index != COLD(ignored)) {
                        const DPrimitive* prim = sc.primitives + prim_index;
                        float4 m[3] = {__ldg(&prim->inv[0]), __ldg(&prim->inv[1]), __ldg(&prim->inv[2])};
                        RayT oray;
                        oray.o = xform(m, ray.o, 1.0f); oray.d = xform(m, ray.d, 0.0f);     // transform_ray :403-409
                        oray.inv = v3(0.0f); oray.neg = 0u;
                        if (STATS) ctr.instances += 1u;
                        uint32_t type = __ldg(&prim->type);
                        if (type != BPT_PRIM_SPHERE) {                                      // the sphere test never reads inv_d
                            make_ray(oray, oray.o, oray.d);
                            if (!tame) oray.neg &= ~BPT_RAY_TAME;
                        }
                        if (type == BPT_PRIM_SPHERE) {
                            if (sphere_test(oray, __ldg(&prim->sphere_r), t)) {
                                COLD(hit_prim) = prim_index; COLD(hit_tri) = 0xFFFFFFFFu;
                                if (occlusion()) finish();
                            }
                        } else if (type == BPT_PRIM_BOX) {
                            if (box_test(oray, __ldg(&prim->box_r[0]), __ldg(&prim->box_r[1]), __ldg(&prim->box_r[2]), t)) {
                                COLD(hit_prim) = prim_index; COLD(hit_tri) = 0xFFFFFFFFu;
                                if (occlusion()) finish();
                            }
                        } else if (type == BPT_PRIM_MESH) {
                            const DMesh* mesh = sc.meshes + __ldg(&prim->mesh);
                            if (STATS) { ctr.mesh_calls += 1u; c_pops = 1u; c_inner = 0u; c_leaves = 0u; }
                            float4 q0 = __ldg(&mesh->root_q0), q1 = __ldg(&mesh->root_q1);
                            float tn;
                            bool root_hit = slab_any(oray, q0, q1, tn) && (tn < t);          // intersect_mesh pops its root first (:269-277)
                            if (root_hit) {
                                // intersect_mesh is "called": park the world-level state, switch to the object-space ray
                                COLDW(0) = ray.o.x; COLDW(1) = ray.o.y; COLDW(2) = ray.o.z;
                                COLDW(3) = ray.d.x; COLDW(4) = ray.d.y; COLDW(5) = ray.d.z;
                                COLD(items_first) = cur_a; COLD(items_count) = cur_b;
                                ray = oray; blas_sp = sp;
                                pair_base = __ldg(&mesh->pair_base);
                                COLD(tri_base) = __ldg(&mesh->tri_base);
                                COLD(cur_prim) = prim_index;
                                enter(__float_as_uint(q1.z));
                            } else if (STATS) {
                                ctr.blas_pops += 1u;
                            }
                        }
                    }
                }
                // The leaf's items are used up (and the lane neither entered a BLAS nor finished on an occlusion hit): the
                // reference's item loop ends and its node loop pops the next TLAS node.  An empty TLAS stack ends the ray
                // right here, in the same step, instead of parking the lane for two more scheduling rounds.
                if (phase == P_ITEMS && cur_b == 0u) {
                    pop_enter();
                    if (phase == P_RET && blas_sp < 0) finish();
                }
            }
          } while (__popc(__ballot_sync(FULL, phase >= P_ITEMS)) >= keep);
        }
    }
#ifndef BPT_DBG_PLAIN_LOOP
    if (trips == BPT_TRIP_LIMIT) *sc.error_flag = BPT_DEVERR_TRIP_LIMIT;
#endif
    #undef COLD
    #undef COLDW
}

} // namespace bpt
