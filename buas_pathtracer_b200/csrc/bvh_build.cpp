// Host BVH construction: Wald-2007 16-bin SAH, output bit-identical to the reference's
// create_bvh / create_bvh_for_mesh (Raytracer/bvh.cpp:138-213, :222-287, :328-426) and
// create_scene_bvh (Raytracer/scene.cpp:173-242).
//
// Written from scratch as an explicit-stack builder (the reference recurses).  What has to match for the
// node arrays to memcmp-equal the reference's:
//   * float op order of every bound / centroid / SAH expression (no FMA contraction; built with
//     -ffp-contract=off), ternary min/max (my_math.h:77-85);
//   * the binning quirk that bin `i` is counted on BOTH sides of candidate split i (bvh.cpp:195-197 reads
//     l_splits[bin_index], not bin_index-1) and that candidate 15 always evaluates to NaN;
//   * the Hoare partition's exact swap sequence (bvh.cpp:26-51), which fixes the item order inside leaves;
//   * node numbering: root = 0, slot 1 skipped, children allocated pairwise when the parent is split and the
//     whole left subtree numbered before the right one (bvh.cpp:259-272, :302-303).
#include "host_scene.h"

#include <float.h>
#include <math.h>
#include <string.h>

namespace bpt {

namespace {

const float kEpsilon = 0.001f;          // common.h:35
const uint32_t kMaxLeaf = 4;            // bvh.h:23
const int kBins = 16;                   // bvh.cpp:147

inline float fmin_t(float a, float b) { return a < b ? a : b; }   // my_math.h:77-85
inline float fmax_t(float a, float b) { return a > b ? a : b; }

struct Box3 {
    float lo[3], hi[3];
    void invert() { for (int k = 0; k < 3; ++k) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; } }   // my_math.h:1099-1105
};

inline void box_union(Box3& a, const Box3& b) {                   // union_of(a, b) -> a
    for (int k = 0; k < 3; ++k) { a.lo[k] = fmin_t(a.lo[k], b.lo[k]); a.hi[k] = fmax_t(a.hi[k], b.hi[k]); }
}

inline float surface_area(const Box3& b) {                         // my_math.h:1131-1138
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return 2.0f*(dx*dy + dx*dz + dy*dz);
}

inline uint32_t largest_axis(const Box3& b) {                      // my_math.h:1113-1129
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    uint32_t axis = 0; float m = dx;
    if (m < dy) { m = dy; axis = 1; }
    if (m < dz) { m = dz; axis = 2; }
    return axis;
}

// (u32)float as x86-64 does it (cvttss2si to 64 bit, keep the low word): NaN -> 0.
inline uint32_t float_to_u32_x86(float v) {
    if (!(v == v)) return 0;
    if (v >= 9223372036854775808.0f || v <= -9223372036854775808.0f) return 0;
    return (uint32_t)(int64_t)v;
}

struct BinAcc {
    uint32_t count;
    Box3 bounds;
};

uint32_t hoare_partition(SortEntry* e, uint32_t count, float split_p, uint32_t axis);

// Returns the split index (0 = "no split") and writes the axis.  bvh.cpp:138-213 + :26-51.
uint32_t partition_sah_binned(SortEntry* e, uint32_t count, const Box3& bv, const Box3& cr, uint32_t* out_axis) {
    float parent_sah = (float)count*surface_area(bv);
    float best_sah = parent_sah;
    float split_p = 0.0f;
    uint32_t axis = largest_axis(cr);

    BinAcc bins[kBins];
    for (int b = 0; b < kBins; ++b) { bins[b].count = 0; bins[b].bounds.invert(); }

    float k0 = cr.lo[axis];
    float k1 = ((float)kBins*(1.0f - kEpsilon)) / (cr.hi[axis] - cr.lo[axis]);

    for (uint32_t i = 0; i < count; ++i) {
        const SortEntry& s = e[i];
        uint32_t b = float_to_u32_x86(k1*(s.p[axis] - k0));
        if (b >= (uint32_t)kBins) b = kBins - 1;   // unreachable for finite input (reference would index out of bounds)
        BinAcc& bin = bins[b];
        bin.count += 1;
        for (int k = 0; k < 3; ++k) {
            bin.bounds.lo[k] = fmin_t(bin.bounds.lo[k], s.p[k] - s.r[k]);
            bin.bounds.hi[k] = fmax_t(bin.bounds.hi[k], s.p[k] + s.r[k]);
        }
    }

    // prefix over bins 0..14; entry 15 stays {0, inverted}
    BinAcc left[kBins];
    for (int b = 0; b < kBins; ++b) { left[b].count = 0; left[b].bounds.invert(); }
    {
        BinAcc run; run.count = 0; run.bounds.invert();
        for (int b = 0; b < kBins - 1; ++b) {
            run.count += bins[b].count;
            box_union(run.bounds, bins[b].bounds);
            left[b] = run;
        }
    }

    // suffix sweep 15..1, evaluating candidate b against left[b] (sic)
    {
        BinAcc run; run.count = 0; run.bounds.invert();
        for (int b = kBins - 1; b >= 1; --b) {
            run.count += bins[b].count;
            box_union(run.bounds, bins[b].bounds);

            float l_sah = (float)left[b].count*surface_area(left[b].bounds);
            float r_sah = (float)run.count*surface_area(run.bounds);
            float sah = l_sah + r_sah;
            if ((sah > 0.0f) && (sah < best_sah)) {
                best_sah = sah;
                split_p = k0 + ((float)b / k1);
            }
        }
    }

    *out_axis = 0;
    if (!(best_sah < parent_sah)) return 0;
    *out_axis = axis;

    return hoare_partition(e, count, split_p, axis);
}

// The Hoare partition of bvh.cpp:26-51 around (split_p, axis).  The reference's scans are unguarded; running past either
// end of the range can only end in "split_index == 0" or "split_index > count-1", both of which mean "make a leaf"
// (bvh.cpp:254) and neither of which swaps anything, so the guarded scans are equivalent.
uint32_t hoare_partition(SortEntry* e, uint32_t count, float split_p, uint32_t axis) {
    int64_t i = -1, j = (int64_t)count;
    for (;;) {
        do { ++i; } while (i < (int64_t)count && e[i].p[axis] < split_p);
        do { --j; } while (j >= 0 && e[j].p[axis] > split_p);
        if (i >= j) break;
        SortEntry tmp = e[i]; e[i] = e[j]; e[j] = tmp;
    }
    if (i >= (int64_t)count) return count;   // caller turns this into a leaf
    return (uint32_t)i;
}

// partition_objects_midpoint_split (bvh.cpp:53-62): largest axis of the BOUNDS, pivot at their centre
uint32_t partition_midpoint(SortEntry* e, uint32_t count, const Box3& bv, uint32_t* out_axis) {
    uint32_t axis = largest_axis(bv);
    float pivot = 0.5f*(bv.lo[axis] + bv.hi[axis]);
    *out_axis = axis;
    return hoare_partition(e, count, pivot, axis);
}

// partition_objects_sah + evaluate_sah (bvh.cpp:64-136): every centroid is a candidate plane, O(n^2)
uint32_t partition_sah_full(SortEntry* e, uint32_t count, const Box3& bv, const Box3& cr, uint32_t* out_axis) {
    float parent_sah = (float)count*surface_area(bv);
    float best_sah = parent_sah;
    float best_split_p = 0.0f;
    uint32_t best_axis = 0;
    uint32_t axis = largest_axis(cr);
    for (uint32_t c = 0; c < count; ++c) {
        float split_p = e[c].p[axis];
        uint32_t l_count = 0, r_count = 0;
        Box3 l, r;
        l.invert(); r.invert();
        for (uint32_t i = 0; i < count; ++i) {
            Box3& side = e[i].p[axis] <= split_p ? l : r;
            if (e[i].p[axis] <= split_p) ++l_count; else ++r_count;
            for (int k = 0; k < 3; ++k) {
                side.lo[k] = fmin_t(side.lo[k], e[i].p[k] - e[i].r[k]);
                side.hi[k] = fmax_t(side.hi[k], e[i].p[k] + e[i].r[k]);
            }
        }
        float l_sah = surface_area(l)*(float)l_count;
        float r_sah = surface_area(r)*(float)r_count;
        float sah = l_sah + r_sah;
        if (best_sah > sah) { best_sah = sah; best_split_p = split_p; best_axis = axis; }
    }
    *out_axis = 0;
    if (!(best_sah < parent_sah)) return 0;
    *out_axis = best_axis;
    return hoare_partition(e, count, best_split_p, best_axis);
}

struct BuildTask {
    uint32_t node, first, count;
};

} // namespace

void build_bvh_sah_binned(std::vector<SortEntry>& entries, HostBVH* out) { build_bvh(entries, out, BPT_BVH_SAH_BINNED); }

void build_bvh(std::vector<SortEntry>& entries, HostBVH* out, int method) {
    uint32_t n = (uint32_t)entries.size();
    // construct_bvh_internal (bvh.cpp:289-326): 2N zeroed slots, root = 0, slot 1 skipped
    std::vector<bpt_bvh_node> nodes((size_t)2*n + 2);
    memset(nodes.data(), 0, nodes.size()*sizeof(bpt_bvh_node));
    uint32_t node_count = 2;

    std::vector<BuildTask> stack;
    stack.reserve(128);
    stack.push_back({0, 0, n});

    while (!stack.empty()) {
        BuildTask t = stack.back();
        stack.pop_back();
        SortEntry* e = entries.data() + t.first;
        bpt_bvh_node* node = &nodes[t.node];

        // compute_bounding_volume (bvh.cpp:6-17)
        Box3 bv, cr;
        bv.invert(); cr.invert();
        for (uint32_t i = 0; i < t.count; ++i) {
            for (int k = 0; k < 3; ++k) {
                bv.lo[k] = fmin_t(bv.lo[k], e[i].p[k] - e[i].r[k]);
                bv.hi[k] = fmax_t(bv.hi[k], e[i].p[k] + e[i].r[k]);
                cr.lo[k] = fmin_t(cr.lo[k], e[i].p[k]);
                cr.hi[k] = fmax_t(cr.hi[k], e[i].p[k]);
            }
        }
        for (int k = 0; k < 3; ++k) {
            node->bv_p[k] = 0.5f*(bv.lo[k] + bv.hi[k]);
            node->bv_r[k] = 0.5f*(bv.hi[k] - bv.lo[k]);
        }

        bool make_leaf = t.count <= kMaxLeaf;
        if (!make_leaf) {
            uint32_t axis = 0;
            uint32_t split = method == BPT_BVH_MIDPOINT_SPLIT ? partition_midpoint(e, t.count, bv, &axis)
                           : method == BPT_BVH_SAH_FULL      ? partition_sah_full(e, t.count, bv, cr, &axis)
                                                             : partition_sah_binned(e, t.count, bv, cr, &axis);
            if (split == 0 || split > t.count - 1) {
                make_leaf = true;
            } else {
                node->split_axis = (uint16_t)axis;
                uint32_t left = node_count; node_count += 2;
                node->left_first = left;
                // right is pushed first so the entire left subtree is numbered before it (DFS)
                stack.push_back({left + 1, t.first + split, t.count - split});
                stack.push_back({left,     t.first,         split});
            }
        }
        if (make_leaf) {
            node->left_first = t.first;
            node->count = (uint16_t)t.count;
        }
    }

    out->nodes.assign(nodes.begin(), nodes.begin() + node_count);
    float ext = 0.0f;
    for (uint32_t i = 0; i < node_count; ++i)
        for (int k = 0; k < 3; ++k) {
            float e = fabsf(nodes[i].bv_p[k]) + fabsf(nodes[i].bv_r[k]);
            if (!(e <= ext)) ext = e;          // NaN sticks
        }
    out->max_abs_extent = ext;
    out->indices.resize(n);
    for (uint32_t i = 0; i < n; ++i) out->indices[i] = entries[i].index;
    build_wide_bvh(out, n);              // the device layout of the same tree (cannot fail for a tree built here)
}

void build_mesh_bvh(HostMesh* mesh) { build_mesh_bvh(mesh, BPT_BVH_SAH_BINNED); }

void build_mesh_bvh(HostMesh* mesh, int method) {
    // create_bvh_for_mesh (bvh.cpp:342-391)
    uint32_t n = mesh->triangle_count;
    std::vector<SortEntry> entries(n);
    const float* tri = mesh->positions.data();
    for (uint32_t i = 0; i < n; ++i, tri += 9) {
        SortEntry& s = entries[i];
        s.index = i;
        for (int k = 0; k < 3; ++k) {
            float a = tri[k], b = tri[3 + k], c = tri[6 + k];
            float lo = fmin_t(a, fmin_t(b, c));
            float hi = fmax_t(a, fmax_t(b, c));
            s.p[k] = 0.5f*(lo + hi);
            s.r[k] = 0.5f*(hi - lo);
        }
    }
    build_bvh(entries, &mesh->bvh, method);
    mesh->leaf_triangles.resize((size_t)n*9);
    for (uint32_t i = 0; i < n; ++i) {
        memcpy(&mesh->leaf_triangles[(size_t)i*9], &mesh->positions[(size_t)mesh->bvh.indices[i]*9], 9*sizeof(float));
    }
}

void build_scene_bvh(bpt_scene* scene) {
    // create_scene_bvh (scene.cpp:173-242)
    std::vector<SortEntry> entries;
    for (uint32_t pi = 1; pi < scene->primitives.size(); ++pi) {
        const HostPrimitive& prim = scene->primitives[pi];
        float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        switch (prim.type) {
            case BPT_PRIM_SPHERE:
                for (int k = 0; k < 3; ++k) { lo[k] = -prim.sphere_r; hi[k] = prim.sphere_r; }
                break;
            case BPT_PRIM_BOX:
                for (int k = 0; k < 3; ++k) { lo[k] = -prim.box_r[k]; hi[k] = prim.box_r[k]; }
                break;
            case BPT_PRIM_MESH: {
                const bpt_bvh_node& root = scene->meshes[prim.mesh].bvh.nodes[0];
                for (int k = 0; k < 3; ++k) { lo[k] = root.bv_p[k] - root.bv_r[k]; hi[k] = root.bv_p[k] + root.bv_r[k]; }
            } break;
            default:
                continue;   // planes never land here; unknown types are skipped like the reference's warning path
        }
        const bpt_m4x4& m = (prim.transform >= 0 ? scene->transforms[prim.transform] : identity_transform()).forward;
        Box3 b; b.invert();
        // corner order of scene.cpp:226-233 (irrelevant for min/max, kept for clarity)
        static const int corner[8][3] = {{0,0,0},{1,0,0},{0,1,0},{0,0,1},{1,1,0},{1,0,1},{0,1,1},{1,1,1}};
        for (int c = 0; c < 8; ++c) {
            float x = corner[c][0] ? hi[0] : lo[0];
            float y = corner[c][1] ? hi[1] : lo[1];
            float z = corner[c][2] ? hi[2] : lo[2];
            float q[3];
            for (int r = 0; r < 3; ++r) q[r] = x*m.e[r][0] + y*m.e[r][1] + z*m.e[r][2] + 1.0f*m.e[r][3];   // my_math.h:947-954
            for (int k = 0; k < 3; ++k) { b.lo[k] = fmin_t(b.lo[k], q[k]); b.hi[k] = fmax_t(b.hi[k], q[k]); }
        }
        SortEntry s;
        s.index = pi;
        for (int k = 0; k < 3; ++k) { s.p[k] = 0.5f*(b.lo[k] + b.hi[k]); s.r[k] = 0.5f*(b.hi[k] - b.lo[k]); }
        entries.push_back(s);
    }
    build_bvh_sah_binned(entries, &scene->tlas);
    scene->has_tlas = true;
}

} // namespace bpt
