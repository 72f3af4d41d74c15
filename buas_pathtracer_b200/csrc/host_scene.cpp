// Host scene model + the host half of the C ABI (include/bpt.h section 1).
// Behaviour follows Raytracer/scene.cpp:9-254 and the camera / filter helpers of Raytracer/raytracer.cpp
// (:26-59, :164-185, :1424-1453); all of it is re-implemented against an index-based model (host_scene.h).
#include "host_scene.h"

#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <mutex>
#include <set>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace bpt {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list args;
    va_start(args, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, args);
    va_end(args);
}

namespace {
std::mutex g_pin_mutex;
std::set<void*> g_malloced;     // blocks that came from malloc because cudaHostAlloc was unavailable
}

void* pinned_alloc(size_t bytes) {
    if (bytes == 0) bytes = 1;
    void* p = nullptr;
    if (bytes >= (1u << 16) && cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess && p) return p;
    cudaGetLastError();             // clear "no driver" etc.; small blocks are not worth pinning
    p = malloc(bytes);
    if (!p) throw std::bad_alloc();
    std::lock_guard<std::mutex> lock(g_pin_mutex);
    g_malloced.insert(p);
    return p;
}

void pinned_free(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lock(g_pin_mutex);
        auto it = g_malloced.find(p);
        if (it != g_malloced.end()) { g_malloced.erase(it); free(p); return; }
    }
    cudaFreeHost(p);
}

const bpt_m4x4inv& identity_transform() {
    static const bpt_m4x4inv id = {
        {{{1,0,0,0},{0,1,0,0},{0,0,1,0},{0,0,0,1}}},
        {{{1,0,0,0},{0,1,0,0},{0,0,1,0},{0,0,0,1}}},
    };
    return id;
}

// ---- the handful of MathLib functions the host path needs, op-for-op (my_math.h:453-500) -----------------------
namespace {

struct F3 { float x, y, z; };
inline F3 f3(const float* v) { return {v[0], v[1], v[2]}; }
inline void st(float* d, F3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
inline float dot3(F3 a, F3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }
inline F3 cross3(F3 a, F3 b) { return {a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x}; }
inline F3 noz3(F3 a) {
    F3 r = {0, 0, 0};
    float lsq = dot3(a, a);
    if ((lsq > 0.0001f) && (lsq < INFINITY)) {
        float l = sqrtf(lsq);
        r = {a.x / l, a.y / l, a.z / l};
    }
    return r;
}
inline F3 normalize3(F3 a) {
    float rcp = 1.0f / sqrtf(dot3(a, a));
    return {a.x*rcp, a.y*rcp, a.z*rcp};
}

const float kPi = 3.14159265359f;      // my_math.h:15

// ---- reconstruction kernels (reconstruction_filters.cpp:9-95) ----------------------------------------------------
float sinc_pi(float x) { return sinf(kPi*x) / (kPi*x); }
float lanczos(float x, float radius) {
    x = fabsf(x);
    if (x < 0.0001f) return 1.0f;
    if (x <= radius) return sinc_pi(x)*sinc_pi(x / radius);
    return 0.0f;
}
float gaussian(float x, float alpha, float radius) {
    float radius_exp = (float)exp(-alpha*radius*radius);   // the reference calls double exp() on a float argument here
    float v = expf(-alpha*x*x) - radius_exp;
    return 0.0f > v ? 0.0f : v;
}
float mitchell_netravali(float x) {
    static const float B = 1.0f / 3.0f;
    static const float C = 1.0f / 3.0f;
    x = fabsf(x);
    if (x > 1.0f) {
        return ((-B - 6*C)*x*x*x + (6*B + 30*C)*x*x + (-12*B - 48*C)*x + (8*B + 24*C))*(1.0f / 6.0f);
    }
    return ((12 - 9*B - 6*C)*x*x*x + (-18 + 12*B + 6*C)*x*x + (6 - 2*B))*(1.0f / 6.0f);
}

struct FilterOption { const char* name; int kind; uint32_t radius; };
// g_filters order (reconstruction_filters.cpp:97-106)
const FilterOption kFilters[] = {
    {"Box", 0, 0}, {"Gaussian 3", 1, 3}, {"Gaussian 12", 2, 12}, {"Mitchell Netravali", 3, 2},
    {"Lanczos 3", 4, 3}, {"Lanczos 4", 5, 4}, {"Lanczos 6", 6, 6}, {"Lanczos 12", 7, 12},
};

float eval_filter(int kind, float x) {
    switch (kind) {
        case 1: return gaussian(x, 3.0f, 3.0f);
        case 2: return gaussian(x, 0.03f, 12.0f);
        case 3: return mitchell_netravali(x);
        case 4: return lanczos(x, 3.0f);
        case 5: return lanczos(x, 4.0f);
        case 6: return lanczos(x, 6.0f);
        case 7: return lanczos(x, 12.0f);
    }
    return 0.0f;
}

void load_filter(bpt_filter_cache* cache, int index) {
    // load_reconstruction_kernel (raytracer.cpp:164-185): 256 samples over [0, radius], upper half stays zero
    memset(cache, 0, sizeof(*cache));
    const FilterOption& f = kFilters[index];
    if (f.kind != 0) {
        cache->kernel_size = f.radius;
        cache->cache_size = 256;
        for (uint32_t i = 0; i < 256; ++i) {
            float x = ((float)f.radius*(float)i) / (float)(256 - 1);
            cache->cache[i] = eval_filter(f.kind, x);
        }
    }
}

const char* kIntegrators[] = {"Advanced Pathtracer", "Whitted", "Ground Truth Recursive",
                              "Ground Truth Iterative", "Normals", "Distances"};

bpt::HostPrimitive* add_primitive(bpt_scene* s, uint32_t type, uint32_t material, const bpt_m4x4inv* xf, uint32_t* out_id) {
    // add_primitive (scene.cpp:69-105)
    std::vector<bpt::HostPrimitive>& buf = (type == BPT_PRIM_PLANE) ? s->planes : s->primitives;
    uint32_t id = (uint32_t)buf.size();
    buf.emplace_back();
    bpt::HostPrimitive* p = &buf.back();
    p->type = type;
    p->material = material;
    if (xf) {
        s->transforms.push_back(*xf);
        p->transform = (int32_t)s->transforms.size() - 1;
    }
    if (material < s->materials.size() && (s->materials[material].flags & BPT_MATERIAL_EMISSIVE)) {
        s->lights.push_back(id);
    }
    s->has_tlas = false;
    *out_id = id;
    return p;
}

} // namespace

void recompute_camera(bpt_camera* c) {
    // recompute_camera (raytracer.cpp:50-59); note vfov is NOT halved
    float film_w = c->aspect_ratio, film_h = 1.0f;
    c->half_film_w = 0.5f*film_w;
    c->half_film_h = 0.5f*film_h;
    c->film_distance = film_h / tanf(c->vfov);
}

} // namespace bpt

using namespace bpt;

extern "C" {

const char* bpt_last_error(void) { return g_error; }
const char* bpt_version(void) { return "buas-pathtracer_b200 0.1 (sm_100a)"; }

bpt_scene* bpt_scene_create(void) {
    bpt_scene* s = new bpt_scene();
    // init_scene (raytracer.cpp:1424-1453)
    bpt_material null_material; memset(&null_material, 0, sizeof(null_material));
    s->materials.push_back(null_material);
    s->primitives.emplace_back();
    load_filter(&s->filter, 3);
    bpt_settings& st = s->new_settings;
    memset(&st, 0, sizeof(st));
    st.next_event_estimation = 1;
    st.importance_sample_lights = 1;
    st.importance_sample_diffuse = 1;
    st.use_mis = 1;
    st.russian_roulette = 1;
    st.sampling_strategy = BPT_SAMPLING_STRATIFIED;
    st.use_path_guide = 0;
    st.caustics = 1;
    st.lens_distortion = 1.0f;
    st.f_factor = 0.0f;
    st.diaphragm_edges = 6.0f;
    st.phi_shutter_max = 0.5f;
    st.vignette_strength = 0.25f;
    st.samples_per_pixel = 1;
    st.max_bounce_count = 12;
    st.integrator = BPT_INTEGRATOR_ADVANCED;
    memset(&s->new_camera, 0, sizeof(s->new_camera));
    return s;
}

void bpt_scene_destroy(bpt_scene* s) { delete s; }

uint32_t bpt_add_material(bpt_scene* s, const bpt_material* m) {
    uint32_t id = (uint32_t)s->materials.size();
    bpt_material mat = *m;
    if (mat.emission_color[0]*1.0f + mat.emission_color[1]*1.0f + mat.emission_color[2]*1.0f > 0.0f) {
        mat.flags |= BPT_MATERIAL_EMISSIVE;
    }
    s->materials.push_back(mat);
    return id;
}

uint32_t bpt_add_diffuse_material(bpt_scene* s, const float c[3], float ior, float roughness, int32_t checkers, const float cc[3]) {
    bpt_material m; memset(&m, 0, sizeof(m));
    if (checkers) m.flags |= BPT_MATERIAL_CHECKERS;
    memcpy(m.checker_color, cc, 12);
    memcpy(m.albedo, c, 12);
    m.ior = ior;
    m.roughness = roughness;
    s->materials.push_back(m);
    return (uint32_t)s->materials.size() - 1;
}

uint32_t bpt_add_translucent_material(bpt_scene* s, const float absorb[3], float ior, float roughness) {
    bpt_material m; memset(&m, 0, sizeof(m));
    m.is_participating_medium = 1;
    memcpy(m.absorb, absorb, 12);
    m.ior = ior;
    m.roughness = roughness;
    s->materials.push_back(m);
    return (uint32_t)s->materials.size() - 1;
}

uint32_t bpt_add_emissive_material(bpt_scene* s, const float e[3]) {
    bpt_material m; memset(&m, 0, sizeof(m));
    m.flags |= BPT_MATERIAL_EMISSIVE;
    memcpy(m.emission_color, e, 12);
    s->materials.push_back(m);
    return (uint32_t)s->materials.size() - 1;
}

uint32_t bpt_add_plane(bpt_scene* s, uint32_t mat, const float n[3], float d) {
    uint32_t id;
    HostPrimitive* p = add_primitive(s, BPT_PRIM_PLANE, mat, nullptr, &id);
    st(p->plane_n, noz3(f3(n)));
    p->plane_d = d;
    return id;
}

uint32_t bpt_add_sphere(bpt_scene* s, uint32_t mat, float r, const bpt_m4x4inv* xf) {
    uint32_t id;
    HostPrimitive* p = add_primitive(s, BPT_PRIM_SPHERE, mat, xf, &id);
    p->sphere_r = r;
    return id;
}

uint32_t bpt_add_box(bpt_scene* s, uint32_t mat, const float r[3], const bpt_m4x4inv* xf) {
    uint32_t id;
    HostPrimitive* p = add_primitive(s, BPT_PRIM_BOX, mat, xf, &id);
    memcpy(p->box_r, r, 12);
    return id;
}

uint32_t bpt_create_mesh(bpt_scene* s, uint32_t triangle_count, const float* positions, const float* normals) {
    if (!positions || triangle_count == 0) { set_error("bpt_create_mesh: empty mesh"); return 0xFFFFFFFFu; }
    s->meshes.emplace_back();
    HostMesh& m = s->meshes.back();
    m.triangle_count = triangle_count;
    m.positions.assign(positions, positions + (size_t)triangle_count*9);
    if (normals) {
        m.has_normals = true;
        m.normals.assign(normals, normals + (size_t)triangle_count*9);
    }
    build_mesh_bvh(&m);
    return (uint32_t)s->meshes.size() - 1;
}

uint32_t bpt_create_mesh_ex(bpt_scene* s, uint32_t triangle_count, const float* positions, const float* normals, int32_t method) {
    if (method < BPT_BVH_MIDPOINT_SPLIT || method > BPT_BVH_SAH_FULL) { set_error("bpt_create_mesh_ex: unknown method %d", method); return 0xFFFFFFFFu; }
    if (!positions || triangle_count == 0) { set_error("bpt_create_mesh_ex: empty mesh"); return 0xFFFFFFFFu; }
    s->meshes.emplace_back();
    HostMesh& m = s->meshes.back();
    m.triangle_count = triangle_count;
    m.positions.assign(positions, positions + (size_t)triangle_count*9);
    if (normals) {
        m.has_normals = true;
        m.normals.assign(normals, normals + (size_t)triangle_count*9);
    }
    build_mesh_bvh(&m, method);
    return (uint32_t)s->meshes.size() - 1;
}

uint32_t bpt_create_mesh_with_bvh(bpt_scene* s, uint32_t triangle_count, const float* positions, const float* normals,
                                  const bpt_bvh_node* nodes, uint32_t node_count, const uint32_t* indices) {
    if (!positions || triangle_count == 0 || !nodes || node_count == 0 || !indices) { set_error("bpt_create_mesh_with_bvh: null argument"); return 0xFFFFFFFFu; }
    for (uint32_t i = 0; i < triangle_count; ++i)
        if (indices[i] >= triangle_count) { set_error("bpt_create_mesh_with_bvh: index %u out of range", indices[i]); return 0xFFFFFFFFu; }
    s->meshes.emplace_back();
    HostMesh& m = s->meshes.back();
    m.triangle_count = triangle_count;
    m.positions.assign(positions, positions + (size_t)triangle_count*9);
    if (normals) {
        m.has_normals = true;
        m.normals.assign(normals, normals + (size_t)triangle_count*9);
    }
    m.bvh.nodes.assign(nodes, nodes + node_count);
    m.bvh.indices.assign(indices, indices + triangle_count);
    float ext = 0.0f;
    for (uint32_t i = 0; i < node_count; ++i)
        for (int k = 0; k < 3; ++k) {
            float e = fabsf(nodes[i].bv_p[k]) + fabsf(nodes[i].bv_r[k]);
            if (!(e <= ext)) ext = e;
        }
    m.bvh.max_abs_extent = ext;
    // the re-layout walks the whole tree and rejects malformed node arrays (children past the array, cycles, leaf ranges
    // past the triangles) that would otherwise make the device traversal read out of bounds or never end
    if (build_wide_bvh(&m.bvh, triangle_count) != BPT_OK) { s->meshes.pop_back(); return 0xFFFFFFFFu; }
    m.leaf_triangles.resize((size_t)triangle_count*9);
    for (uint32_t i = 0; i < triangle_count; ++i)
        memcpy(&m.leaf_triangles[(size_t)i*9], &m.positions[(size_t)indices[i]*9], 9*sizeof(float));
    return (uint32_t)s->meshes.size() - 1;
}

uint32_t bpt_add_mesh(bpt_scene* s, uint32_t mat, uint32_t mesh, const bpt_m4x4inv* xf) {
    if (mesh >= s->meshes.size()) { set_error("bpt_add_mesh: unknown mesh %u", mesh); return 0xFFFFFFFFu; }
    uint32_t id;
    HostPrimitive* p = add_primitive(s, BPT_PRIM_MESH, mat, xf, &id);
    p->mesh = mesh;
    return id;
}

int bpt_set_sky(bpt_scene* s, const float top[3], const float bot[3]) {
    memcpy(s->top_sky_color, top, 12);
    memcpy(s->bot_sky_color, bot, 12);
    return BPT_OK;
}

int bpt_set_ambient_light(bpt_scene* s, const float rgb[3]) {
    memcpy(s->ambient_light, rgb, 12);
    return BPT_OK;
}

int bpt_set_skydome(bpt_scene* s, uint32_t w, uint32_t h, const float* pixels) {
    if (!pixels) { s->skydome.clear(); s->skydome_w = s->skydome_h = 0; return BPT_OK; }
    s->skydome_w = w; s->skydome_h = h;
    s->skydome.assign(pixels, pixels + (size_t)w*h*3);
    return BPT_OK;
}

int bpt_get_camera(const bpt_scene* s, bpt_camera* out) { *out = s->new_camera; return BPT_OK; }
int bpt_set_camera(bpt_scene* s, const bpt_camera* c) { s->new_camera = *c; return BPT_OK; }

int bpt_aim_camera(bpt_scene* s, const float d[3]) {
    // aim_camera (raytracer.cpp:26-40)
    bpt_camera* c = &s->new_camera;
    F3 z = noz3(f3(d));
    F3 x = noz3(cross3({0, 1, 0}, z));
    F3 y = noz3(cross3(z, x));
    st(c->z, z); st(c->x, x); st(c->y, y);
    recompute_camera(c);
    return BPT_OK;
}

int bpt_aim_camera_at(bpt_scene* s, const float at[3]) {
    // aim_camera_at (raytracer.cpp:42-48)
    bpt_camera* c = &s->new_camera;
    F3 v = {at[0] - c->p[0], at[1] - c->p[1], at[2] - c->p[2]};
    F3 d = normalize3(v);
    float neg[3] = {-d.x, -d.y, -d.z};
    bpt_aim_camera(s, neg);
    c->focus_distance = sqrtf(dot3(v, v));
    return BPT_OK;
}

int bpt_get_settings(const bpt_scene* s, bpt_settings* out) { *out = s->new_settings; return BPT_OK; }
int bpt_set_settings(bpt_scene* s, const bpt_settings* in) {
    s->new_settings = *in;
    if (s->new_settings.integrator < 0 || s->new_settings.integrator > BPT_INTEGRATOR_DISTANCES) s->new_settings.integrator = 0;
    return BPT_OK;
}

int bpt_find_integrator(const char* name) {
    for (int i = 0; i < 6; ++i) if (0 == strcmp(name, kIntegrators[i])) return i;
    return 0;   // the reference returns the default when not found (integrators.cpp:836)
}

int bpt_load_reconstruction_kernel(bpt_scene* s, const char* filter_name) {
    int index = 0;   // Box when not found (reconstruction_filters.cpp:112)
    for (int i = 0; i < 8; ++i) if (0 == strcmp(filter_name, kFilters[i].name)) { index = i; break; }
    load_filter(&s->filter, index);
    return index;
}

int bpt_get_filter_cache(const bpt_scene* s, bpt_filter_cache* out) { *out = s->filter; return BPT_OK; }
int bpt_set_filter_cache(bpt_scene* s, const bpt_filter_cache* in) {
    if (in->kernel_size > 12 || in->cache_size > 256) { set_error("bpt_set_filter_cache: radius > 12 or LUT > 256"); return BPT_ERR_ARG; }
    s->filter = *in;
    return BPT_OK;
}

int bpt_create_scene_bvh(bpt_scene* s) {
    build_scene_bvh(s);
    return BPT_OK;
}

int bpt_get_scene_bvh(const bpt_scene* s, const bpt_bvh_node** nodes, uint32_t* node_count, const uint32_t** indices, uint32_t* index_count) {
    if (!s->has_tlas) { set_error("bpt_get_scene_bvh: call bpt_create_scene_bvh first"); return BPT_ERR_STATE; }
    *nodes = s->tlas.nodes.data(); *node_count = (uint32_t)s->tlas.nodes.size();
    *indices = s->tlas.indices.data(); *index_count = (uint32_t)s->tlas.indices.size();
    return BPT_OK;
}

int bpt_get_mesh_bvh(const bpt_scene* s, uint32_t mesh, const bpt_bvh_node** nodes, uint32_t* node_count,
                     const uint32_t** indices, uint32_t* index_count, const float** tris) {
    if (mesh >= s->meshes.size()) { set_error("bpt_get_mesh_bvh: unknown mesh %u", mesh); return BPT_ERR_ARG; }
    const HostMesh& m = s->meshes[mesh];
    *nodes = m.bvh.nodes.data(); *node_count = (uint32_t)m.bvh.nodes.size();
    *indices = m.bvh.indices.data(); *index_count = (uint32_t)m.bvh.indices.size();
    if (tris) *tris = m.leaf_triangles.data();
    return BPT_OK;
}

int bpt_get_wide_bvh(const bpt_scene* s, int32_t mesh, const bpt_wide_child** pairs, uint32_t* pair_count,
                     bpt_wide_child* root, const uint32_t** big_leaves, uint32_t* big_leaf_count, uint32_t* depth) {
    static_assert(sizeof(bpt_wide_child) == sizeof(WChild) && sizeof(WBigLeaf) == 8, "wide BVH view layout");
    if (!s || !pairs || !pair_count || !root || !big_leaves || !big_leaf_count || !depth) { set_error("bpt_get_wide_bvh: null argument"); return BPT_ERR_ARG; }
    const HostBVH* bvh = nullptr;
    if (mesh < 0) {
        if (!s->has_tlas) { set_error("bpt_get_wide_bvh: call bpt_create_scene_bvh first"); return BPT_ERR_STATE; }
        bvh = &s->tlas;
    } else {
        if ((size_t)mesh >= s->meshes.size()) { set_error("bpt_get_wide_bvh: unknown mesh %d", mesh); return BPT_ERR_ARG; }
        bvh = &s->meshes[(size_t)mesh].bvh;
    }
    if (!bvh->wide.valid) { set_error("bpt_get_wide_bvh: no device layout was built"); return BPT_ERR_STATE; }
    *pairs = (const bpt_wide_child*)bvh->wide.pairs.data(); *pair_count = (uint32_t)bvh->wide.pairs.size();
    memcpy(root, &bvh->wide.root, sizeof(WChild));
    *big_leaves = (const uint32_t*)bvh->wide.big_leaves.data(); *big_leaf_count = (uint32_t)bvh->wide.big_leaves.size();
    *depth = bvh->wide.depth;
    return BPT_OK;
}

int bpt_get_counts(const bpt_scene* s, uint32_t* materials, uint32_t* primitives, uint32_t* planes, uint32_t* lights, uint32_t* meshes) {
    if (materials)  *materials  = (uint32_t)s->materials.size();
    if (primitives) *primitives = (uint32_t)s->primitives.size();
    if (planes)     *planes     = (uint32_t)s->planes.size();
    if (lights)     *lights     = (uint32_t)s->lights.size();
    if (meshes)     *meshes     = (uint32_t)s->meshes.size();
    return BPT_OK;
}

int bpt_write_bitmap(const char* file_name, const uint32_t* pixels, uint32_t w, uint32_t h) {
    // write_bitmap (assets.cpp:671-724): BITMAPINFOHEADER, 32 bpp, negative height = top-down, 4096 px/m
    if (!file_name || !pixels || w == 0 || h == 0) { set_error("bpt_write_bitmap: bad arguments"); return BPT_ERR_ARG; }
    uint32_t pixel_size = 4u*w*h;
    unsigned char hdr[54];
    memset(hdr, 0, sizeof(hdr));
    auto put16 = [&](int at, uint16_t v) { memcpy(hdr + at, &v, 2); };
    auto put32 = [&](int at, uint32_t v) { memcpy(hdr + at, &v, 4); };
    put16(0, 0x4D42); put32(2, 54 + pixel_size); put32(10, 54); put32(14, 40);
    put32(18, w); put32(22, (uint32_t)(-(int32_t)h)); put16(26, 1); put16(28, 32);
    put32(30, 0); put32(34, pixel_size); put32(38, 4096); put32(42, 4096);
    FILE* f = fopen(file_name, "wb");
    if (!f) { set_error("bpt_write_bitmap: cannot open %s", file_name); return BPT_ERR_ARG; }
    bool ok = fwrite(hdr, 1, 54, f) == 54 && fwrite(pixels, 1, pixel_size, f) == pixel_size;
    fclose(f);
    if (!ok) { set_error("bpt_write_bitmap: short write to %s", file_name); return BPT_ERR_ARG; }
    return BPT_OK;
}

// ---- procedural inputs for the BASELINE.json configs --------------------------------------------------------------

// bpt_make_displaced_icosphere / bpt_make_procedural_skydome: procedural_inputs.cpp

} // extern "C"
