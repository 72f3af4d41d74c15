// Host-side scene model of the B200 path-tracing core.
//
// Mirrors the *interface* of the reference's Scene (Raytracer/scene.h:91-149) -- same builder calls, same
// index conventions (material 0 / primitive 0 are null entries, planes live in their own array, lights are
// PrimitiveIDs) -- but is an index-based, pointer-free model that can be flattened to the device in one pass.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/bpt.h"
#include "wide_bvh.h"

namespace bpt {

// Page-locked host memory for the arrays bpt_upload_scene() copies to the device (so the H2D copies run at PCIe speed
// and asynchronously).  Falls back to malloc when no CUDA driver is present (host-only use, e.g. the CPU test suite).
void* pinned_alloc(size_t bytes);
void  pinned_free(void* p);

template <class T>
struct PinnedAllocator {
    typedef T value_type;
    PinnedAllocator() {}
    template <class U> PinnedAllocator(const PinnedAllocator<U>&) {}
    T* allocate(size_t n) { return (T*)pinned_alloc(n*sizeof(T)); }
    void deallocate(T* p, size_t) { pinned_free(p); }
    template <class U> bool operator==(const PinnedAllocator<U>&) const { return true; }
    template <class U> bool operator!=(const PinnedAllocator<U>&) const { return false; }
};
template <class T> using PinnedVec = std::vector<T, PinnedAllocator<T>>;

// The device layout of a BVH (wide_bvh.h): built once on the host next to the reference-format node array, uploaded as is.
struct WideBVH {
    PinnedVec<WPair> pairs;              // records of three sibling pairs, depth-first
    std::vector<WBigLeaf> big_leaves;    // leaves holding more than BPT_WREF_INLINE_COUNT_MAX items
    WChild root{};                       // the tree root as a child record (it is box-tested like any node, intersection.cpp:269-277)
    uint32_t depth = 0;                  // deepest leaf below the root = most far children one ray can have pending
    int  mode = 0;
    bool valid = false;
};

struct HostBVH {
    PinnedVec<bpt_bvh_node> nodes;       // bit-identical to BVH::nodes[0..node_count) (bvh.h:39-45)
    PinnedVec<uint32_t>     indices;     // BVH::indices
    float max_abs_extent = 0.0f;         // max over nodes/axes of |bv_p| + |bv_r| (NaN/inf propagate): FMNMX slab-test precondition
    WideBVH wide;                        // the same tree as the device reads it
};

struct HostMesh {
    uint32_t triangle_count = 0;
    bool     has_normals = false;
    std::vector<float> positions;        // 9 floats per triangle, caller order (Mesh::triangles)
    PinnedVec<float> normals;            // 9 floats per triangle, caller order (get_normals(mesh))
    HostBVH  bvh;                        // MeshBVH (bvh.h:52-58), BVHStorage_Scalar
    PinnedVec<float> leaf_triangles;     // MeshBVH::triangles: positions re-ordered into leaf order
};

struct HostPrimitive {                   // Primitive (primitives.h:92-106) without pointers
    uint32_t type = BPT_PRIM_NONE;
    uint32_t material = 0;
    int32_t  transform = -1;             // index into Scene::transforms, -1 = the shared identity (scene.cpp:76)
    float    plane_n[3] = {0, 0, 0};
    float    plane_d = 0;
    float    sphere_r = 0;
    float    box_r[3] = {0, 0, 0};
    uint32_t mesh = 0;
};

} // namespace bpt

struct bpt_scene {
    std::vector<bpt_material>       materials;
    std::vector<uint32_t>           lights;
    std::vector<bpt::HostPrimitive> planes;
    std::vector<bpt::HostPrimitive> primitives;
    std::vector<bpt_m4x4inv>        transforms;
    std::vector<bpt::HostMesh>      meshes;
    bpt::HostBVH                    tlas;
    bool                            has_tlas = false;

    float top_sky_color[3] = {0, 0, 0};
    float bot_sky_color[3] = {0, 0, 0};
    float ambient_light[3] = {0, 0, 0};
    uint32_t skydome_w = 0, skydome_h = 0;
    bpt::PinnedVec<float> skydome;       // w*h*3

    bpt_camera   new_camera{};           // Scene::new_camera (what the user edits)
    bpt_settings new_settings{};         // Scene::new_settings
    bpt_filter_cache filter{};           // g_filter_cache (Raytracer.h:42), per scene here
};

namespace bpt {

void set_error(const char* fmt, ...);

struct SortEntry {                       // BVHSortEntry (bvh.h:25-29)
    uint32_t index;
    float p[3];
    float r[3];
};

// bvh_build.cpp
void build_bvh_sah_binned(std::vector<SortEntry>& entries, HostBVH* out);
void build_mesh_bvh(HostMesh* mesh);
void build_bvh(std::vector<SortEntry>& entries, HostBVH* out, int method);      // BPT_BVH_* (bvh.h:7-11)
void build_mesh_bvh(HostMesh* mesh, int method);
void build_scene_bvh(bpt_scene* scene);

// wide_bvh.cpp: re-layout into two-level pair records (validates the node array; BPT_ERR_ARG when it is malformed)
int build_wide_bvh(const bpt_bvh_node* nodes, uint32_t node_count, uint32_t item_count, WideBVH* out, int mode);
int build_wide_bvh(HostBVH* bvh, uint32_t item_count);

// host_scene.cpp
void recompute_camera(bpt_camera* camera);
const bpt_m4x4inv& identity_transform();

} // namespace bpt
