// Device construction of the reference's mesh BVH: the Wald-2007 16-bin SAH build of
// Raytracer/bvh.cpp:138-287, reproduced NODE FOR NODE -- same node array (numbering included), same item order inside the
// leaves -- as csrc/bvh_build.cpp (the host builder, itself memcmp-equal to the reference).
//
// The reference is a depth-first recursion over one array that it partitions in place; nothing in it is parallel as
// written.  What makes a parallel build with identical output possible:
//   * a node's result depends only on the SET and ORDER of its entries, so all nodes of one tree level can be processed
//     at once (level-synchronous, breadth first); the depth-first NUMBERING is restored afterwards from subtree sizes;
//   * bounds and bins are min/max/count reductions (exact in any order; see the note on signed zeros below); the SAH
//     sweep over 16 bins is tiny and stays sequential per node, with the host builder's float expressions verbatim;
//   * the Hoare partition (bvh.cpp:26-51) has a closed form: with A = positions where the up-scan stops (!(e < p)),
//     ascending, and B = positions where the down-scan stops (!(e > p)), descending, it swaps A[k] <-> B[k] for
//     k < m = #{k : A[k] < B[k]} and returns min(A[m], B[m-1]) (A[0] when m = 0).  Ranks come from two prefix sums.
// One thread per entry for the reductions / partition, one thread per node for the decisions; up to ~40 levels of a
// dozen small launches each.  Signed zeros: the reference's ternary min/max keep the LAST of two equal values, an atomic
// min keeps -0 over +0; a bound can differ in the sign of a zero only when both signs occur among the tied extremes.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace bpt {
namespace gbvh {

constexpr uint32_t kInvalid = 0xFFFFFFFFu;
constexpr int kBins = 16;
constexpr uint32_t kMaxLeaf = 4;          // bvh.h:23
constexpr float kEpsilonB = 0.001f;       // common.h:35

struct GEntry { float p[3]; uint32_t index; float r[3]; uint32_t pad; };      // BVHSortEntry, 32 B

struct GNode {                            // breadth-first node record
    uint32_t first, count;
    uint32_t child;                       // breadth-first id of the left child (right = child + 1); kInvalid = leaf
    uint32_t axis;
    float bv_p[3], bv_r[3];
    uint32_t inner_count;                 // inner nodes in this subtree
    uint32_t rank;                        // pre-order rank among inner nodes (= order of the reference's node allocation)
    uint32_t final_index;                 // index in the reference's node array
    uint32_t pad;
};

struct GAcc {                             // per node of the current level
    uint32_t bv_lo[3], bv_hi[3], cr_lo[3], cr_hi[3];          // order-preserving encodings of floats
    uint32_t bin_count[kBins];
    uint32_t bin_lo[kBins][3], bin_hi[kBins][3];
    float k0, k1, split_p;
    uint32_t axis;
    uint32_t state;                       // 0 = leaf, 1 = wants a split (binning), 2 = splits
    uint32_t swaps;                       // m
    uint32_t split_rel;
    uint32_t child_local;                 // index of the left child within the next level
};

__device__ __forceinline__ uint32_t enc(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float dec(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }
__device__ __forceinline__ float fmin_t(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float fmax_t(float a, float b) { return a > b ? a : b; }

__device__ __forceinline__ uint32_t float_to_u32_x86(float v) {           // see bvh_build.cpp
    if (!(v == v)) return 0;
    if (v >= 9223372036854775808.0f || v <= -9223372036854775808.0f) return 0;
    return (uint32_t)(long long)v;
}

struct Box3 { float lo[3], hi[3]; };
__device__ __forceinline__ void box_invert(Box3& b) { for (int k = 0; k < 3; ++k) { b.lo[k] = FLT_MAX; b.hi[k] = -FLT_MAX; } }
__device__ __forceinline__ void box_union(Box3& a, const Box3& b) {
    for (int k = 0; k < 3; ++k) { a.lo[k] = fmin_t(a.lo[k], b.lo[k]); a.hi[k] = fmax_t(a.hi[k], b.hi[k]); }
}
__device__ __forceinline__ float surface_area(const Box3& b) {
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return 2.0f*(dx*dy + dx*dz + dy*dz);
}
__device__ __forceinline__ uint32_t largest_axis(const Box3& b) {
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    uint32_t axis = 0; float m = dx;
    if (m < dy) { m = dy; axis = 1; }
    if (m < dz) { m = dz; axis = 2; }
    return axis;
}

// create_bvh_for_mesh's entry setup (bvh.cpp:342-391): one BVHSortEntry per triangle
__global__ void k_make_entries(const float* __restrict__ positions, uint32_t n, GEntry* __restrict__ out, uint32_t* __restrict__ seg) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* tri = positions + (size_t)i*9;
    GEntry e;
    e.index = i; e.pad = 0;
    for (int k = 0; k < 3; ++k) {
        float a = tri[k], b = tri[3 + k], c = tri[6 + k];
        float lo = fmin_t(a, fmin_t(b, c));
        float hi = fmax_t(a, fmax_t(b, c));
        e.p[k] = 0.5f*(lo + hi);
        e.r[k] = 0.5f*(hi - lo);
    }
    out[i] = e;
    seg[i] = 0;
}

__global__ void k_init_root(GNode* nodes, uint32_t n) {
    GNode r = {};
    r.first = 0; r.count = n; r.child = kInvalid;
    nodes[0] = r;
}

__global__ void k_init_acc(GAcc* acc, uint32_t count) {
    uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= count) return;
    GAcc& a = acc[j];
    for (int k = 0; k < 3; ++k) {
        a.bv_lo[k] = a.cr_lo[k] = enc(FLT_MAX);
        a.bv_hi[k] = a.cr_hi[k] = enc(-FLT_MAX);
    }
    for (int b = 0; b < kBins; ++b) {
        a.bin_count[b] = 0;
        for (int k = 0; k < 3; ++k) { a.bin_lo[b][k] = enc(FLT_MAX); a.bin_hi[b][k] = enc(-FLT_MAX); }
    }
    a.state = 0; a.swaps = 0; a.split_rel = 0; a.child_local = kInvalid; a.axis = 0; a.k0 = a.k1 = a.split_p = 0.0f;
}

// compute_bounding_volume (bvh.cpp:6-17) for every node of the level at once.  Warps that lie inside one node reduce
// with REDUX first; blocks that lie inside one node combine their warps in shared memory and issue 12 global atomics.
__global__ void k_bounds(const GEntry* __restrict__ e, const uint32_t* __restrict__ seg, uint32_t n, GAcc* acc) {
    __shared__ uint32_t s_v[12];
    __shared__ uint32_t s_node;
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    uint32_t s = i < n ? seg[i] : kInvalid;
    if (threadIdx.x == 0) s_node = s;
    if (threadIdx.x < 12) s_v[threadIdx.x] = (threadIdx.x % 6) < 3 ? enc(FLT_MAX) : enc(-FLT_MAX);   // lo,hi,lo,hi triples
    __syncthreads();
    const uint32_t node = s_node;
    const int block_uniform = __syncthreads_and(i >= n || s == node) && node != kInvalid;
    uint32_t v[12];
    for (int k = 0; k < 12; ++k) v[k] = (k % 6) < 3 ? enc(FLT_MAX) : enc(-FLT_MAX);
    if (s != kInvalid) {
        GEntry q = e[i];
        for (int k = 0; k < 3; ++k) {
            v[k] = enc(q.p[k] - q.r[k]); v[3 + k] = enc(q.p[k] + q.r[k]);
            v[6 + k] = enc(q.p[k]);      v[9 + k] = enc(q.p[k]);
        }
    }
    if (block_uniform) {
        for (int k = 0; k < 12; ++k) {
            uint32_t r = (k % 6) < 3 ? __reduce_min_sync(0xFFFFFFFFu, v[k]) : __reduce_max_sync(0xFFFFFFFFu, v[k]);
            if ((threadIdx.x & 31u) == 0) { if ((k % 6) < 3) atomicMin(&s_v[k], r); else atomicMax(&s_v[k], r); }
        }
        __syncthreads();
        if (threadIdx.x < 12) {
            uint32_t k = threadIdx.x;
            uint32_t* dst = k < 3 ? &acc[node].bv_lo[k] : k < 6 ? &acc[node].bv_hi[k - 3] : k < 9 ? &acc[node].cr_lo[k - 6] : &acc[node].cr_hi[k - 9];
            if ((k % 6) < 3) atomicMin(dst, s_v[k]); else atomicMax(dst, s_v[k]);
        }
        return;
    }
    uint32_t s0 = __shfl_sync(0xFFFFFFFFu, s, 0);
    if (__all_sync(0xFFFFFFFFu, s == s0) && s0 != kInvalid) {
        for (int k = 0; k < 3; ++k) {
            uint32_t a = __reduce_min_sync(0xFFFFFFFFu, v[k]), b = __reduce_max_sync(0xFFFFFFFFu, v[3 + k]);
            uint32_t c = __reduce_min_sync(0xFFFFFFFFu, v[6 + k]), d = __reduce_max_sync(0xFFFFFFFFu, v[9 + k]);
            if ((threadIdx.x & 31u) == 0) {
                atomicMin(&acc[s0].bv_lo[k], a); atomicMax(&acc[s0].bv_hi[k], b);
                atomicMin(&acc[s0].cr_lo[k], c); atomicMax(&acc[s0].cr_hi[k], d);
            }
        }
    } else if (s != kInvalid) {
        for (int k = 0; k < 3; ++k) {
            atomicMin(&acc[s].bv_lo[k], v[k]);     atomicMax(&acc[s].bv_hi[k], v[3 + k]);
            atomicMin(&acc[s].cr_lo[k], v[6 + k]); atomicMax(&acc[s].cr_hi[k], v[9 + k]);
        }
    }
}

// bv_p / bv_r, the leaf test, and the binning parameters (bvh.cpp:233-236, :141-155)
__global__ void k_decide(GNode* nodes, GAcc* acc, uint32_t count, int midpoint) {
    uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= count) return;
    GNode& nd = nodes[j];
    GAcc& a = acc[j];
    Box3 bv, cr;
    for (int k = 0; k < 3; ++k) { bv.lo[k] = dec(a.bv_lo[k]); bv.hi[k] = dec(a.bv_hi[k]); cr.lo[k] = dec(a.cr_lo[k]); cr.hi[k] = dec(a.cr_hi[k]); }
    for (int k = 0; k < 3; ++k) {
        nd.bv_p[k] = 0.5f*(bv.lo[k] + bv.hi[k]);
        nd.bv_r[k] = 0.5f*(bv.hi[k] - bv.lo[k]);
    }
    nd.child = kInvalid; nd.axis = 0;
    if (nd.count <= kMaxLeaf) { a.state = 0; return; }
    if (midpoint) {                     // partition_objects_midpoint_split (bvh.cpp:53-62): no binning, no SAH test
        uint32_t ax = largest_axis(bv);
        a.axis = ax;
        a.split_p = 0.5f*(bv.lo[ax] + bv.hi[ax]);
        a.state = 2;
        return;
    }
    uint32_t axis = largest_axis(cr);
    a.axis = axis;
    a.k0 = cr.lo[axis];
    a.k1 = ((float)kBins*(1.0f - kEpsilonB)) / (cr.hi[axis] - cr.lo[axis]);
    a.state = 1;
}

// Binning (bvh.cpp:157-166).  Near the root every thread of a block feeds the same node's 16 bins: those blocks
// accumulate in shared memory and issue one set of global atomics per block (7 per entry became 0.4 per entry; the
// plain version spent 68 % of a 1.3 M-triangle build in this kernel).  Blocks that straddle nodes use global atomics.
__global__ void k_bin(const GEntry* __restrict__ e, const uint32_t* __restrict__ seg, uint32_t n, GAcc* acc) {
    __shared__ uint32_t s_cnt[kBins], s_lo[kBins][3], s_hi[kBins][3];
    __shared__ uint32_t s_node;
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    uint32_t s = i < n ? seg[i] : kInvalid;
    if (threadIdx.x == 0) s_node = s;
    if (threadIdx.x < kBins) {
        s_cnt[threadIdx.x] = 0;
        for (int k = 0; k < 3; ++k) { s_lo[threadIdx.x][k] = enc(FLT_MAX); s_hi[threadIdx.x][k] = enc(-FLT_MAX); }
    }
    __syncthreads();
    const uint32_t node = s_node;
    const int uniform = __syncthreads_and(i >= n || s == node) && node != kInvalid;
    bool active = s != kInvalid && acc[s].state == 1;
    uint32_t b = 0;
    GEntry q;
    if (active) {
        q = e[i];
        const GAcc& a = acc[s];
        b = float_to_u32_x86(a.k1*(q.p[a.axis] - a.k0));
        if (b >= (uint32_t)kBins) b = kBins - 1;
    }
    if (uniform) {
        if (active) {
            atomicAdd(&s_cnt[b], 1u);
            for (int k = 0; k < 3; ++k) {
                atomicMin(&s_lo[b][k], enc(q.p[k] - q.r[k]));
                atomicMax(&s_hi[b][k], enc(q.p[k] + q.r[k]));
            }
        }
        __syncthreads();
        GAcc& a = acc[node];
        if (threadIdx.x < kBins && s_cnt[threadIdx.x] != 0) {
            uint32_t t = threadIdx.x;
            atomicAdd(&a.bin_count[t], s_cnt[t]);
            for (int k = 0; k < 3; ++k) { atomicMin(&a.bin_lo[t][k], s_lo[t][k]); atomicMax(&a.bin_hi[t][k], s_hi[t][k]); }
        }
    } else if (active) {
        GAcc& a = acc[s];
        atomicAdd(&a.bin_count[b], 1u);
        for (int k = 0; k < 3; ++k) {
            atomicMin(&a.bin_lo[b][k], enc(q.p[k] - q.r[k]));
            atomicMax(&a.bin_hi[b][k], enc(q.p[k] + q.r[k]));
        }
    }
}

// the sweep of partition_objects_sah_binned (bvh.cpp:170-205), one thread per node, float expressions as in bvh_build.cpp
__global__ void k_sah(GNode* nodes, GAcc* acc, uint32_t count) {
    uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= count) return;
    GAcc& a = acc[j];
    if (a.state != 1) return;
    const GNode& nd = nodes[j];
    Box3 bv;
    for (int k = 0; k < 3; ++k) { bv.lo[k] = dec(a.bv_lo[k]); bv.hi[k] = dec(a.bv_hi[k]); }
    float parent_sah = (float)nd.count*surface_area(bv);
    float best_sah = parent_sah;
    float split_p = 0.0f;

    uint32_t left_count[kBins];
    Box3 left_box[kBins];
    {
        uint32_t rc = 0; Box3 run; box_invert(run);
        for (int b = 0; b < kBins; ++b) { left_count[b] = 0; box_invert(left_box[b]); }
        for (int b = 0; b < kBins - 1; ++b) {
            Box3 bb; for (int k = 0; k < 3; ++k) { bb.lo[k] = dec(a.bin_lo[b][k]); bb.hi[k] = dec(a.bin_hi[b][k]); }
            rc += a.bin_count[b];
            box_union(run, bb);
            left_count[b] = rc; left_box[b] = run;
        }
    }
    {
        uint32_t rc = 0; Box3 run; box_invert(run);
        for (int b = kBins - 1; b >= 1; --b) {
            Box3 bb; for (int k = 0; k < 3; ++k) { bb.lo[k] = dec(a.bin_lo[b][k]); bb.hi[k] = dec(a.bin_hi[b][k]); }
            rc += a.bin_count[b];
            box_union(run, bb);
            float l_sah = (float)left_count[b]*surface_area(left_box[b]);
            float r_sah = (float)rc*surface_area(run);
            float sah = l_sah + r_sah;
            if ((sah > 0.0f) && (sah < best_sah)) {
                best_sah = sah;
                split_p = a.k0 + ((float)b / a.k1);
            }
        }
    }
    if (!(best_sah < parent_sah)) { a.state = 0; return; }
    a.split_p = split_p;
    a.state = 2;
}

// where the two scans of the Hoare partition stop (bvh.cpp:30-40); also resets the permutation to the identity
__global__ void k_flags(const GEntry* __restrict__ e, const uint32_t* __restrict__ seg, uint32_t n, const GAcc* __restrict__ acc,
                        uint2* __restrict__ flags, uint32_t* __restrict__ perm) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    perm[i] = i;
    uint2 f = make_uint2(0u, 0u);
    uint32_t s = seg[i];
    if (s != kInvalid && acc[s].state == 2) {
        float v = e[i].p[acc[s].axis], p = acc[s].split_p;
        f.x = (v < p) ? 0u : 1u;
        f.y = (v > p) ? 0u : 1u;
    }
    flags[i] = f;
}

// ---- exclusive prefix sums of the two flag columns: scan[i] = sum of flags[0..i), scan[n] = total ----------------------
constexpr int kScanBlock = 256, kScanItems = 8, kScanTile = kScanBlock*kScanItems;

__global__ void k_scan_tiles(const uint2* __restrict__ flags, uint32_t n, uint2* __restrict__ scan, uint2* __restrict__ tile_sums) {
    __shared__ uint2 warp_sums[kScanBlock/32];
    uint32_t base = blockIdx.x*kScanTile + threadIdx.x*kScanItems;
    uint2 v[kScanItems];
    uint2 sum = make_uint2(0u, 0u);
    for (int k = 0; k < kScanItems; ++k) {
        uint32_t i = base + k;
        uint2 f = i < n ? flags[i] : make_uint2(0u, 0u);
        v[k] = sum;
        sum.x += f.x; sum.y += f.y;
    }
    // exclusive scan of the per-thread sums across the block
    uint2 incl = sum;
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t ax = __shfl_up_sync(0xFFFFFFFFu, incl.x, d), ay = __shfl_up_sync(0xFFFFFFFFu, incl.y, d);
        if ((threadIdx.x & 31) >= d) { incl.x += ax; incl.y += ay; }
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint2 warp_off = make_uint2(0u, 0u);
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) { warp_off.x += warp_sums[w].x; warp_off.y += warp_sums[w].y; }
    uint2 excl = make_uint2(incl.x - sum.x + warp_off.x, incl.y - sum.y + warp_off.y);
    for (int k = 0; k < kScanItems; ++k) {
        uint32_t i = base + k;
        if (i < n) scan[i] = make_uint2(v[k].x + excl.x, v[k].y + excl.y);
    }
    if (threadIdx.x == kScanBlock - 1) tile_sums[blockIdx.x] = make_uint2(excl.x + sum.x, excl.y + sum.y);
}

// one block scans the tile sums in place (exclusive); writes the grand total to scan[n]
__global__ void k_scan_tile_sums(uint2* tile_sums, uint32_t tiles, uint2* scan, uint32_t n) {
    __shared__ uint2 carry;
    __shared__ uint2 warp_sums[32];
    if (threadIdx.x == 0) carry = make_uint2(0u, 0u);
    __syncthreads();
    for (uint32_t base = 0; base < tiles; base += blockDim.x) {
        uint32_t i = base + threadIdx.x;
        uint2 v = i < tiles ? tile_sums[i] : make_uint2(0u, 0u);
        uint2 incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t ax = __shfl_up_sync(0xFFFFFFFFu, incl.x, d), ay = __shfl_up_sync(0xFFFFFFFFu, incl.y, d);
            if ((threadIdx.x & 31) >= d) { incl.x += ax; incl.y += ay; }
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint2 off = carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) { off.x += warp_sums[w].x; off.y += warp_sums[w].y; }
        if (i < tiles) tile_sums[i] = make_uint2(incl.x - v.x + off.x, incl.y - v.y + off.y);
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = make_uint2(incl.x + off.x, incl.y + off.y);
        __syncthreads();
    }
    if (threadIdx.x == 0) scan[n] = carry;
}

__global__ void k_scan_add(uint2* __restrict__ scan, uint32_t n, const uint2* __restrict__ tile_sums) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint2 o = tile_sums[i / kScanTile];
    uint2 s = scan[i];
    scan[i] = make_uint2(s.x + o.x, s.y + o.y);
}

// A[k] (ascending) and B[k] (descending) of every splitting node, stored at posA/posB[first + k]
__global__ void k_scatter(const uint32_t* __restrict__ seg, uint32_t n, const GNode* __restrict__ nodes, const GAcc* __restrict__ acc,
                          const uint2* __restrict__ flags, const uint2* __restrict__ scan,
                          uint32_t* __restrict__ posA, uint32_t* __restrict__ posB) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = seg[i];
    if (s == kInvalid || acc[s].state != 2) return;
    uint32_t first = nodes[s].first, cnt = nodes[s].count;
    uint2 f = flags[i], sc = scan[i], s0 = scan[first], s1 = scan[first + cnt];
    if (f.x) posA[first + (sc.x - s0.x)] = i;
    if (f.y) posB[first + ((s1.y - s0.y) - 1u - (sc.y - s0.y))] = i;
}

// the swaps: k-th pair is exchanged iff A[k] < B[k]
__global__ void k_pair(const uint32_t* __restrict__ seg, uint32_t n, const GNode* __restrict__ nodes, GAcc* acc,
                       const uint2* __restrict__ scan, const uint32_t* __restrict__ posA, const uint32_t* __restrict__ posB,
                       uint32_t* __restrict__ perm) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = seg[i];
    if (s == kInvalid || acc[s].state != 2) return;
    uint32_t first = nodes[s].first, cnt = nodes[s].count;
    uint2 s0 = scan[first], s1 = scan[first + cnt];
    uint32_t nA = s1.x - s0.x, nB = s1.y - s0.y;
    uint32_t k = i - first;
    if (k >= nA || k >= nB) return;
    uint32_t a = posA[i], b = posB[i];
    if (a < b) {
        perm[a] = b; perm[b] = a;
        atomicMax(&acc[s].swaps, k + 1u);
    }
}

// split index, the "make a leaf after all" rule (bvh.cpp:254), and allocation of the two children in the next level
__global__ void k_split(GNode* nodes, GAcc* acc, uint32_t count, const uint2* __restrict__ scan,
                        const uint32_t* __restrict__ posA, const uint32_t* __restrict__ posB,
                        GNode* next_nodes, uint32_t next_base, uint32_t* next_count) {
    uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= count) return;
    GAcc& a = acc[j];
    GNode& nd = nodes[j];
    if (a.state != 2) return;
    uint32_t first = nd.first, cnt = nd.count;
    uint2 s0 = scan[first], s1 = scan[first + cnt];
    uint32_t nA = s1.x - s0.x;
    uint32_t m = a.swaps;
    uint32_t split;
    if (m == 0) split = nA > 0 ? posA[first] - first : cnt;
    else {
        uint32_t viaB = posB[first + m - 1] - first;
        uint32_t viaA = nA > m ? posA[first + m] - first : kInvalid;
        split = viaA < viaB ? viaA : viaB;
    }
    if (split == 0 || split > cnt - 1) { a.state = 0; return; }
    a.split_rel = split;
    uint32_t c = atomicAdd(next_count, 2u);
    a.child_local = c;
    nd.child = next_base + c;
    nd.axis = a.axis;
    GNode l = {}, r = {};
    l.first = first;         l.count = split;       l.child = kInvalid;
    r.first = first + split; r.count = cnt - split; r.child = kInvalid;
    next_nodes[c] = l; next_nodes[c + 1] = r;
}

// apply the permutation into the other entry buffer and hand every entry to its node of the next level
__global__ void k_apply(const GEntry* __restrict__ in, GEntry* __restrict__ out, const uint32_t* __restrict__ perm,
                        const uint32_t* __restrict__ seg, uint32_t* __restrict__ seg_out, uint32_t n,
                        const GNode* __restrict__ nodes, const GAcc* __restrict__ acc) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = in[perm[i]];
    uint32_t s = seg[i], ns = kInvalid;
    if (s != kInvalid && acc[s].state == 2) ns = acc[s].child_local + ((i - nodes[s].first) < acc[s].split_rel ? 0u : 1u);
    seg_out[i] = ns;
}

// ---- depth-first numbering (bvh.cpp:259-272: children are allocated when the parent splits, left subtree first) -------
__global__ void k_inner_count(GNode* all, uint32_t base, uint32_t count) {
    uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= count) return;
    GNode& nd = all[base + j];
    nd.inner_count = nd.child == kInvalid ? 0u : 1u + all[nd.child].inner_count + all[nd.child + 1].inner_count;
}

__global__ void k_rank(GNode* all, uint32_t base, uint32_t count) {
    uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= count) return;
    GNode& nd = all[base + j];
    if (base == 0) { nd.rank = 0; nd.final_index = 0; }
    if (nd.child == kInvalid) return;
    GNode& l = all[nd.child];
    GNode& r = all[nd.child + 1];
    l.final_index = 2u + 2u*nd.rank;  l.rank = nd.rank + 1u;
    r.final_index = l.final_index + 1u; r.rank = nd.rank + 1u + l.inner_count;
}

struct OutNode { float bv_p[3]; float bv_r[3]; uint32_t left_first; uint16_t count; uint16_t split_axis; };   // == bpt_bvh_node

__global__ void k_emit(const GNode* __restrict__ all, uint32_t total, OutNode* __restrict__ out) {
    uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= total) return;
    const GNode& nd = all[j];
    OutNode o;
    for (int k = 0; k < 3; ++k) { o.bv_p[k] = nd.bv_p[k]; o.bv_r[k] = nd.bv_r[k]; }
    if (nd.child == kInvalid) { o.left_first = nd.first; o.count = (uint16_t)nd.count; o.split_axis = 0; }
    else { o.left_first = 2u + 2u*nd.rank; o.count = 0; o.split_axis = (uint16_t)nd.axis; }
    out[nd.final_index] = o;
}

__global__ void k_indices(const GEntry* __restrict__ e, uint32_t n, uint32_t* __restrict__ out) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) out[i] = e[i].index;
}

} // namespace gbvh
} // namespace bpt
