// Device half of the C ABI (include/bpt.h section 2): context, scene flattening/upload, the per-pass wavefront
// schedule, diagnostics.  There is no CPU fallback anywhere in this file: without a usable CUDA device every entry
// point fails with BPT_ERR_CUDA.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "host_scene.h"
#include "kernels.cuh"

using namespace bpt;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return BPT_ERR_CUDA; } } while (0)

namespace {

enum Stage { ST_RAYGEN, ST_TRACE, ST_SHADE, ST_SHADOW, ST_SPLAT, ST_COUNT };

struct TimedSpan { cudaEvent_t a, b; int stage; };

} // namespace

struct bpt_ctx {
    int device = 0;
    int sm_count = 148;
    uint32_t refill = 8;                  // persistent warps fetch new rays once this many lanes are idle (or idle is the largest group)
    int trace_ctas_per_sm = 8;            // resident CTAs of the persistent traversal kernels (occupancy query)
    cudaStream_t stream = nullptr;

    DScene sc{};
    bool scene_ready = false;
    bool tables_ready = false;
    std::vector<void*> scene_allocs;      // freed on re-upload / destroy
    uint32_t* tri_original = nullptr;     // DTriangle slot -> original triangle index (MeshBVH::indices)

    uint8_t *d_perm = nullptr, *d_sobol = nullptr, *d_scramble = nullptr, *d_rank = nullptr;
    float* d_filter = nullptr;

    float4* film = nullptr;
    bool film_owned = false;
    uint32_t film_w = 0, film_h = 0;

    // Two independent batch pipelines (stream + path state + queues): consecutive batches alternate between them so
    // the tail of one batch's persistent kernels overlaps the next batch's head.
    struct Pipe {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        uint32_t max_slots = 0;
        DPathState st{};
        DQueues q{};
        std::vector<void*> allocs;
        bpt_sample_record* d_records = nullptr;
        uint64_t d_record_capacity = 0;
    } pipes[2];
    int n_pipes = 2;
    int32_t* d_row_map = nullptr;
    uint32_t row_map_capacity = 0;

    DStats* d_stats = nullptr;
    bool stats_enabled = false;

    bpt_sample_record* host_records = nullptr;
    uint64_t host_record_capacity = 0;
    bool detailed_timing = false;
    std::vector<TimedSpan> spans;
    size_t spans_used = 0;
    cudaEvent_t pass_begin = nullptr, pass_end = nullptr;
    bool pass_recorded = false;
    uint32_t launches = 0, trace_launches = 0;
    uint64_t total_launches = 0;
    uint64_t h2d_bytes = 0, d2h_bytes = 0;
    uint64_t samples = 0;
};

namespace {

template <typename T>
int upload(bpt_ctx* ctx, const T* host, size_t count, const T** out, std::vector<void*>* owner) {
    *out = nullptr;
    size_t bytes = std::max<size_t>(count*sizeof(T), 256);
    void* d = nullptr;
    CK(cudaMalloc(&d, bytes));
    owner->push_back(d);
    if (count) CK(cudaMemcpyAsync(d, host, count*sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += count*sizeof(T);
    *out = (const T*)d;
    return BPT_OK;
}

void free_all(std::vector<void*>* v) {
    for (void* p : *v) cudaFree(p);
    v->clear();
}

int ensure_state(bpt_ctx* ctx, bpt_ctx::Pipe* pp, uint32_t slots) {
    (void)ctx;
    if (slots <= pp->max_slots) return BPT_OK;
    free_all(&pp->allocs);
    pp->max_slots = 0;
    auto alloc = [&](void** p, size_t bytes) -> int {
        CK(cudaMalloc(p, bytes));
        pp->allocs.push_back(*p);
        return BPT_OK;
    };
    size_t n = slots;
    int rc = 0;
    rc |= alloc((void**)&pp->st.ray_o, n*16);
    rc |= alloc((void**)&pp->st.ray_d, n*16);
    rc |= alloc((void**)&pp->st.hit, n*16);
    rc |= alloc((void**)&pp->st.hit_w, n*4);
    rc |= alloc((void**)&pp->st.throughput, n*16);
    rc |= alloc((void**)&pp->st.radiance, n*16);
    rc |= alloc((void**)&pp->st.rng, n*16);
    rc |= alloc((void**)&pp->st.prev_n, n*16);
    rc |= alloc((void**)&pp->st.jitter, n*8);
    rc |= alloc((void**)&pp->st.mstack_at, n);
    rc |= alloc((void**)&pp->st.mstack, n*2*BPT_MATERIAL_STACK_DEPTH);
    rc |= alloc((void**)&pp->st.primary_d, n*16);
    rc |= alloc((void**)&pp->st.primary_o, n*16);
    rc |= alloc((void**)&pp->q.active[0], n*4);
    rc |= alloc((void**)&pp->q.active[1], n*4);
    rc |= alloc((void**)&pp->q.shadow, n*sizeof(DShadowItem));
    rc |= alloc((void**)&pp->q.counters, 256);
    if (rc) return BPT_ERR_CUDA;
    pp->max_slots = slots;
    return BPT_OK;
}

void begin_span(bpt_ctx* ctx, int stage, cudaStream_t stream) {
    if (!ctx->detailed_timing) return;
    if (ctx->spans_used == ctx->spans.size()) {
        TimedSpan s; s.stage = stage;
        cudaEventCreate(&s.a); cudaEventCreate(&s.b);
        ctx->spans.push_back(s);
    }
    ctx->spans[ctx->spans_used].stage = stage;
    cudaEventRecord(ctx->spans[ctx->spans_used].a, stream);
}

void end_span(bpt_ctx* ctx, cudaStream_t stream) {
    if (!ctx->detailed_timing) return;
    cudaEventRecord(ctx->spans[ctx->spans_used].b, stream);
    ctx->spans_used++;
}

uint32_t grid_for(const bpt_ctx* ctx, uint64_t work, uint32_t threads, uint32_t ctas_per_sm) {
    uint64_t need = (work + threads - 1)/threads;
    uint64_t cap = (uint64_t)ctx->sm_count*ctas_per_sm;
    return (uint32_t)std::max<uint64_t>(1, std::min(need, cap));
}

void fill_rows(float4* dst, const bpt_m4x4& m) {
    for (int r = 0; r < 3; ++r) dst[r] = make_float4(m.e[r][0], m.e[r][1], m.e[r][2], m.e[r][3]);
}

void latch_settings(bpt_ctx* ctx, const bpt_scene* scene) {
    // what render_all_tiles does when a render (re)starts (raytracer.cpp:711-720)
    bpt_camera cam = scene->new_camera;
    recompute_camera(&cam);
    ctx->sc.camera = cam;
    ctx->sc.settings = scene->new_settings;
    ctx->sc.filter_radius = scene->filter.kernel_size;
    ctx->sc.filter_lut_size = scene->filter.cache_size;
    memcpy(ctx->sc.top_sky, scene->top_sky_color, 12);
    memcpy(ctx->sc.bot_sky, scene->bot_sky_color, 12);
}

} // namespace

extern "C" {

int bpt_create(int device, bpt_ctx** out_ctx) {
    if (!out_ctx) { set_error("bpt_create: null out_ctx"); return BPT_ERR_ARG; }
    *out_ctx = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("bpt_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
        return BPT_ERR_CUDA;
    }
    if (device < 0 || device >= count) { set_error("bpt_create: device %d out of range (0..%d)", device, count - 1); return BPT_ERR_ARG; }
    CK(cudaSetDevice(device));
    bpt_ctx* ctx = new bpt_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (auto& pp : ctx->pipes) {
        CK(cudaStreamCreateWithFlags(&pp.stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&pp.done, cudaEventDisableTiming));
    }
    CK(cudaMalloc((void**)&ctx->d_stats, sizeof(DStats)));
    CK(cudaMemset(ctx->d_stats, 0, sizeof(DStats)));
    CK(cudaMalloc((void**)&ctx->d_filter, 512*sizeof(float)));
    CK(cudaEventCreate(&ctx->pass_begin));
    CK(cudaEventCreate(&ctx->pass_end));
    {
        int a = 0, b = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_trace_closest<false>, 128, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_trace_shadow<false>, 128, 0);
        int m = std::min(a, b);
        if (m > 0) ctx->trace_ctas_per_sm = m;
    }
    if (const char* e = getenv("BPT_REFILL")) { int v = atoi(e); if (v >= 1 && v <= 33) ctx->refill = (uint32_t)v; }
    if (const char* e = getenv("BPT_PIPES")) { int v = atoi(e); if (v >= 1 && v <= 2) ctx->n_pipes = v; }
    if (const char* e = getenv("BPT_TRACE_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 32) ctx->trace_ctas_per_sm = v; }
    const char* dt = getenv("BPT_DETAILED_TIMING");
    ctx->detailed_timing = dt && atoi(dt) != 0;
    *out_ctx = ctx;
    return BPT_OK;
}

void bpt_destroy(bpt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_all(&ctx->scene_allocs);
    for (auto& pp : ctx->pipes) {
        cudaStreamSynchronize(pp.stream);
        free_all(&pp.allocs);
        cudaFree(pp.d_records);
        cudaEventDestroy(pp.done);
        cudaStreamDestroy(pp.stream);
    }
    cudaFree(ctx->d_row_map);
    if (ctx->film_owned && ctx->film) cudaFree(ctx->film);
    cudaFree(ctx->d_stats); cudaFree(ctx->d_filter);
    cudaFree(ctx->d_perm); cudaFree(ctx->d_sobol); cudaFree(ctx->d_scramble); cudaFree(ctx->d_rank);
    for (auto& s : ctx->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    cudaEventDestroy(ctx->pass_begin); cudaEventDestroy(ctx->pass_end);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int bpt_set_sampler_tables(bpt_ctx* ctx, const uint8_t* perm, const uint8_t* sobol, const uint8_t* scramble, const uint8_t* rank) {
    if (!ctx || !perm || !sobol || !scramble || !rank) { set_error("bpt_set_sampler_tables: null argument"); return BPT_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    if (!ctx->d_perm) {
        CK(cudaMalloc((void**)&ctx->d_perm, 256*64));
        CK(cudaMalloc((void**)&ctx->d_sobol, 256*256));
        CK(cudaMalloc((void**)&ctx->d_scramble, 128*128*8));
        CK(cudaMalloc((void**)&ctx->d_rank, 128*128*8));
    }
    CK(cudaMemcpyAsync(ctx->d_perm, perm, 256*64, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_sobol, sobol, 256*256, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_scramble, scramble, 128*128*8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_rank, rank, 128*128*8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->sc.strata_perm = ctx->d_perm; ctx->sc.bn_sobol = ctx->d_sobol;
    ctx->sc.bn_scramble = ctx->d_scramble; ctx->sc.bn_rank = ctx->d_rank;
    ctx->tables_ready = true;
    return BPT_OK;
}

int bpt_update_settings(bpt_ctx* ctx, const bpt_scene* scene) {
    if (!ctx || !scene) { set_error("bpt_update_settings: null argument"); return BPT_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    latch_settings(ctx, scene);
    CK(cudaMemcpyAsync(ctx->d_filter, scene->filter.cache, 512*sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += 512*sizeof(float) + sizeof(bpt_camera) + sizeof(bpt_settings);
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->sc.filter_lut = ctx->d_filter;
    return BPT_OK;
}

int bpt_upload_scene(bpt_ctx* ctx, const bpt_scene* scene) {
    if (!ctx || !scene) { set_error("bpt_upload_scene: null argument"); return BPT_ERR_ARG; }
    if (!scene->has_tlas) { set_error("bpt_upload_scene: call bpt_create_scene_bvh first"); return BPT_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    free_all(&ctx->scene_allocs);
    ctx->scene_ready = false;
    DScene& sc = ctx->sc;
    std::vector<void*>* own = &ctx->scene_allocs;

    // materials (+ the integrator's local "air", integrators.cpp:597-599)
    std::vector<DMaterial> mats(scene->materials.size() + 1);
    memset(mats.data(), 0, mats.size()*sizeof(DMaterial));
    for (size_t i = 0; i < scene->materials.size(); ++i) memcpy(&mats[i], &scene->materials[i], sizeof(bpt_material));
    mats.back().ior = 1.0f;
    mats.back().is_participating_medium = 1;
    if (mats.size() > 0xFFFF) { set_error("bpt_upload_scene: more than 65534 materials"); return BPT_ERR_UNSUPPORTED; }

    // meshes: concatenate BLAS node arrays and leaf-ordered triangles
    std::vector<DMesh> meshes(scene->meshes.size());
    std::vector<bpt_bvh_node> blas_nodes;
    std::vector<DTriangle> tris;
    std::vector<float4> normals;
    std::vector<uint32_t> tri_original;
    bool any_normals = false;
    for (const HostMesh& m : scene->meshes) any_normals |= m.has_normals;
    for (size_t mi = 0; mi < scene->meshes.size(); ++mi) {
        const HostMesh& m = scene->meshes[mi];
        DMesh& dm = meshes[mi];
        dm.node_base = (uint32_t)blas_nodes.size();
        dm.tri_base = (uint32_t)tris.size();
        dm.triangle_count = m.triangle_count;
        dm.has_normals = m.has_normals ? 1u : 0u;
        blas_nodes.insert(blas_nodes.end(), m.bvh.nodes.begin(), m.bvh.nodes.end());
        if (blas_nodes.size() & 1) blas_nodes.emplace_back();     // keep sibling pairs 64-byte aligned
        size_t base = tris.size();
        tris.resize(base + m.triangle_count);
        if (any_normals) normals.resize((base + m.triangle_count)*3, make_float4(0, 0, 0, 0));
        tri_original.insert(tri_original.end(), m.bvh.indices.begin(), m.bvh.indices.end());
        for (uint32_t i = 0; i < m.triangle_count; ++i) {
            const float* p = &m.leaf_triangles[(size_t)i*9];
            DTriangle& t = tris[base + i];
            uint32_t orig = m.bvh.indices[i];
            float orig_bits; memcpy(&orig_bits, &orig, 4);
            t.a_idx = make_float4(p[0], p[1], p[2], orig_bits);
            t.e1 = make_float4(p[3] - p[0], p[4] - p[1], p[5] - p[2], 0.0f);     // edge1 = b - a (intersection.cpp:145)
            t.e2 = make_float4(p[6] - p[0], p[7] - p[1], p[8] - p[2], 0.0f);     // edge2 = c - a (:146)
            if (m.has_normals) {
                const float* nn = &m.normals[(size_t)orig*9];
                for (int k = 0; k < 3; ++k) normals[(base + i)*3 + k] = make_float4(nn[k*3], nn[k*3 + 1], nn[k*3 + 2], 0.0f);
            }
        }
    }

    std::vector<DPrimitive> prims(scene->primitives.size());
    memset(prims.data(), 0, prims.size()*sizeof(DPrimitive));
    for (size_t i = 0; i < scene->primitives.size(); ++i) {
        const HostPrimitive& hp = scene->primitives[i];
        DPrimitive& dp = prims[i];
        const bpt_m4x4inv& xf = hp.transform >= 0 ? scene->transforms[hp.transform] : identity_transform();
        fill_rows(dp.inv, xf.inverse);
        fill_rows(dp.fwd, xf.forward);
        dp.type = hp.type; dp.material = hp.material; dp.mesh = hp.mesh;
        dp.sphere_r = hp.sphere_r;
        memcpy(dp.box_r, hp.box_r, 12);
    }
    std::vector<DPlane> planes(scene->planes.size());
    memset(planes.data(), 0, planes.size()*sizeof(DPlane));
    for (size_t i = 0; i < scene->planes.size(); ++i) {
        memcpy(planes[i].n, scene->planes[i].plane_n, 12);
        planes[i].d = scene->planes[i].plane_d;
        planes[i].material = scene->planes[i].material;
    }
    for (uint32_t l : scene->lights) {
        if (l >= scene->primitives.size()) { set_error("bpt_upload_scene: light id %u is not a primitive (emissive plane?)", l); return BPT_ERR_UNSUPPORTED; }
    }

    // FMNMX slab test precondition (trace.cuh make_ray): all node boxes finite and inside 1e15
    bool tame = true;
    auto check_nodes = [&](const std::vector<bpt_bvh_node>& nodes) {
        for (const bpt_bvh_node& nd : nodes)
            for (int k = 0; k < 3; ++k) {
                float ext = fabsf(nd.bv_p[k]) + fabsf(nd.bv_r[k]);
                if (!(ext < 1e15f)) {
                    // the empty-scene root has bv_r = -inf by construction (it can never be hit); anything else disables the fast path
                    tame = false;
                }
            }
    };
    check_nodes(scene->tlas.nodes);
    check_nodes(blas_nodes);
    sc.tame_bounds = tame ? 1u : 0u;

    int rc = 0;
    const bpt_bvh_node* d_tlas = nullptr; const bpt_bvh_node* d_blas = nullptr;
    rc |= upload(ctx, scene->tlas.nodes.data(), scene->tlas.nodes.size(), &d_tlas, own);
    rc |= upload(ctx, scene->tlas.indices.data(), scene->tlas.indices.size(), &sc.tlas_indices, own);
    rc |= upload(ctx, blas_nodes.data(), blas_nodes.size(), &d_blas, own);
    rc |= upload(ctx, tris.data(), tris.size(), &sc.triangles, own);
    rc |= upload(ctx, meshes.data(), meshes.size(), &sc.meshes, own);
    rc |= upload(ctx, prims.data(), prims.size(), &sc.primitives, own);
    rc |= upload(ctx, planes.data(), planes.size(), &sc.planes, own);
    rc |= upload(ctx, mats.data(), mats.size(), &sc.materials, own);
    rc |= upload(ctx, scene->lights.data(), scene->lights.size(), &sc.lights, own);
    const uint32_t* d_orig = nullptr;
    rc |= upload(ctx, tri_original.data(), tri_original.size(), &d_orig, own);
    sc.normals = nullptr;
    if (any_normals) rc |= upload(ctx, normals.data(), normals.size(), &sc.normals, own);
    sc.skydome = nullptr; sc.skydome_w = sc.skydome_h = 0;
    if (!scene->skydome.empty()) {
        std::vector<float4> sky((size_t)scene->skydome_w*scene->skydome_h);
        for (size_t i = 0; i < sky.size(); ++i) sky[i] = make_float4(scene->skydome[i*3], scene->skydome[i*3 + 1], scene->skydome[i*3 + 2], 0.0f);
        rc |= upload(ctx, sky.data(), sky.size(), &sc.skydome, own);
        sc.skydome_w = scene->skydome_w; sc.skydome_h = scene->skydome_h;
    }
    if (rc) return BPT_ERR_CUDA;
    sc.tlas_nodes = (const DNodeHalf*)d_tlas;
    sc.blas_nodes = (const DNodeHalf*)d_blas;
    ctx->tri_original = (uint32_t*)d_orig;
    sc.plane_count = (uint32_t)planes.size();
    sc.primitive_count = (uint32_t)prims.size();
    sc.material_count = (uint32_t)scene->materials.size();
    sc.air_material = sc.material_count;
    sc.light_count = (uint32_t)scene->lights.size();

    rc = bpt_update_settings(ctx, scene);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->scene_ready = true;
    return BPT_OK;
}

int bpt_film_resize(bpt_ctx* ctx, uint32_t w, uint32_t h) {
    if (!ctx || w == 0 || h == 0) { set_error("bpt_film_resize: bad arguments"); return BPT_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->film_owned && ctx->film) cudaFree(ctx->film);
    ctx->film = nullptr;
    CK(cudaMalloc((void**)&ctx->film, (size_t)w*h*sizeof(float4)));
    ctx->film_owned = true;
    ctx->film_w = w; ctx->film_h = h;
    return bpt_film_clear(ctx);
}

int bpt_film_clear(bpt_ctx* ctx) {
    if (!ctx || !ctx->film) { set_error("bpt_film_clear: no film"); return BPT_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->film, 0, (size_t)ctx->film_w*ctx->film_h*sizeof(float4), ctx->stream));
    return BPT_OK;
}

int bpt_film_use_external(bpt_ctx* ctx, void* device_ptr, uint32_t w, uint32_t h) {
    if (!ctx || !device_ptr || w == 0 || h == 0) { set_error("bpt_film_use_external: bad arguments"); return BPT_ERR_ARG; }
    if (((uintptr_t)device_ptr & 15) != 0) { set_error("bpt_film_use_external: pointer must be 16-byte aligned"); return BPT_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->film_owned && ctx->film) cudaFree(ctx->film);
    ctx->film = (float4*)device_ptr;
    ctx->film_owned = false;
    ctx->film_w = w; ctx->film_h = h;
    return BPT_OK;
}

int bpt_film_device_ptr(bpt_ctx* ctx, void** out) {
    if (!ctx || !out) { set_error("bpt_film_device_ptr: null argument"); return BPT_ERR_ARG; }
    *out = ctx->film;
    return ctx->film ? BPT_OK : BPT_ERR_STATE;
}

int bpt_download_film(bpt_ctx* ctx, float* out) {
    if (!ctx || !out) { set_error("bpt_download_film: null argument"); return BPT_ERR_ARG; }
    if (!ctx->film) { set_error("bpt_download_film: no film"); return BPT_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out, ctx->film, (size_t)ctx->film_w*ctx->film_h*sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->d2h_bytes += (uint64_t)ctx->film_w*ctx->film_h*sizeof(float4);
    return BPT_OK;
}

int bpt_sync(bpt_ctx* ctx) {
    if (!ctx) { set_error("bpt_sync: null ctx"); return BPT_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return BPT_OK;
}

int bpt_stats_enable(bpt_ctx* ctx, int enable) {
    if (!ctx) return BPT_ERR_ARG;
    ctx->stats_enabled = enable != 0;
    return BPT_OK;
}

int bpt_get_stats(bpt_ctx* ctx, bpt_stats* out, int reset) {
    if (!ctx || !out) { set_error("bpt_get_stats: null argument"); return BPT_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    DStats h;
    CK(cudaMemcpy(&h, ctx->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
    out->rays = h.v[0]; out->shadow_rays = h.v[1]; out->tlas_node_pops = h.v[2]; out->instances_visited = h.v[3];
    out->mesh_intersection_count = h.v[4]; out->mesh_bvh_traversals = h.v[5]; out->mesh_node_traversals = h.v[6];
    out->mesh_leaf_traversals = h.v[7]; out->triangles_tested = h.v[8]; out->samples = ctx->samples;
    out->shadow_tlas_node_pops = h.v[10]; out->shadow_instances_visited = h.v[11];
    out->shadow_mesh_intersection_count = h.v[12]; out->shadow_mesh_bvh_traversals = h.v[13];
    out->shadow_mesh_node_traversals = h.v[14]; out->shadow_mesh_leaf_traversals = h.v[15];
    out->shadow_triangles_tested = h.v[16];
    if (reset) { CK(cudaMemset(ctx->d_stats, 0, sizeof(DStats))); ctx->samples = 0; }
    return BPT_OK;
}

int bpt_set_sample_records(bpt_ctx* ctx, bpt_sample_record* host_records, uint64_t capacity) {
    if (!ctx) return BPT_ERR_ARG;
    ctx->host_records = host_records;
    ctx->host_record_capacity = host_records ? capacity : 0;
    return BPT_OK;
}

int bpt_trace(bpt_ctx* ctx, uint32_t n, const bpt_ray* rays, int mode, uint32_t ignored, bpt_hit* out) {
    if (!ctx || (n && (!rays || !out))) { set_error("bpt_trace: null argument"); return BPT_ERR_ARG; }
    if (!ctx->scene_ready) { set_error("bpt_trace: no scene uploaded"); return BPT_ERR_STATE; }
    if (mode != BPT_TRACE_CLOSEST && mode != BPT_TRACE_OCCLUSION) { set_error("bpt_trace: bad mode"); return BPT_ERR_ARG; }
    if (n == 0) return BPT_OK;
    CK(cudaSetDevice(ctx->device));
    bpt_ray* d_rays = nullptr; bpt_hit* d_hits = nullptr;
    CK(cudaMalloc((void**)&d_rays, (size_t)n*sizeof(bpt_ray)));
    CK(cudaMalloc((void**)&d_hits, (size_t)n*sizeof(bpt_hit)));
    CK(cudaMemcpyAsync(d_rays, rays, (size_t)n*sizeof(bpt_ray), cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += (uint64_t)n*sizeof(bpt_ray); ctx->d2h_bytes += (uint64_t)n*sizeof(bpt_hit);
    uint32_t* d_cursor = nullptr;
    CK(cudaMalloc((void**)&d_cursor, 256));
    CK(cudaMemsetAsync(d_cursor, 0, 256, ctx->stream));
    uint32_t grid = grid_for(ctx, n, 128, ctx->trace_ctas_per_sm);
    bool st = ctx->stats_enabled;
    if (mode == BPT_TRACE_CLOSEST) {
        if (st) k_trace_api<false, true ><<<grid, 128, 0, ctx->stream>>>(ctx->sc, d_rays, n, ignored, d_hits, ctx->tri_original, d_cursor, ctx->refill, ctx->d_stats);
        else    k_trace_api<false, false><<<grid, 128, 0, ctx->stream>>>(ctx->sc, d_rays, n, ignored, d_hits, ctx->tri_original, d_cursor, ctx->refill, ctx->d_stats);
    } else {
        if (st) k_trace_api<true, true ><<<grid, 128, 0, ctx->stream>>>(ctx->sc, d_rays, n, ignored, d_hits, ctx->tri_original, d_cursor, ctx->refill, ctx->d_stats);
        else    k_trace_api<true, false><<<grid, 128, 0, ctx->stream>>>(ctx->sc, d_rays, n, ignored, d_hits, ctx->tri_original, d_cursor, ctx->refill, ctx->d_stats);
    }
    ctx->total_launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_hits, (size_t)n*sizeof(bpt_hit), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_rays); cudaFree(d_hits); cudaFree(d_cursor);
    if (e != cudaSuccess) { set_error("bpt_trace: %s", cudaGetErrorString(e)); return BPT_ERR_CUDA; }
    return BPT_OK;
}

static int render_rows(bpt_ctx* ctx, int32_t x0, int32_t x1, const std::vector<int32_t>& rows,
                       uint32_t frame_count, uint32_t spp, uint32_t seed_mode, uint32_t seed_salt, const char* who) {
    if (!ctx->scene_ready) { set_error("%s: no scene uploaded", who); return BPT_ERR_STATE; }
    if (!ctx->tables_ready) { set_error("%s: sampler tables not set (bpt_set_sampler_tables)", who); return BPT_ERR_STATE; }
    if (!ctx->film) { set_error("%s: no film (bpt_film_resize)", who); return BPT_ERR_STATE; }
    if (seed_mode != BPT_SEED_PER_PIXEL) { set_error("%s: unknown seed mode", who); return BPT_ERR_ARG; }
    if (x0 < 0 || x1 > (int32_t)ctx->film_w || x0 >= x1 || rows.empty() || spp == 0) { set_error("%s: bad rect/spp", who); return BPT_ERR_ARG; }
    for (int32_t y : rows) if (y < 0 || y >= (int32_t)ctx->film_h) { set_error("%s: bad rect/spp (row %d)", who, y); return BPT_ERR_ARG; }
    if (ctx->sc.settings.integrator != BPT_INTEGRATOR_ADVANCED) {
        set_error("%s: only the \"Advanced Pathtracer\" integrator runs on the device (SURVEY 8a5)", who);
        return BPT_ERR_UNSUPPORTED;
    }
    if (ctx->sc.filter_lut_size != 0 && ctx->sc.filter_radius == 0) { set_error("%s: filter LUT with radius 0", who); return BPT_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));

    DScene& sc = ctx->sc;
    sc.film_w = ctx->film_w; sc.film_h = ctx->film_h;
    uint32_t rect_w = (uint32_t)(x1 - x0), rect_h = (uint32_t)rows.size();

    bool want_records = ctx->host_records != nullptr;
    // per-stage timing and record read-back want one batch at a time; otherwise two pipelines overlap
    int n_pipes = (ctx->detailed_timing || want_records) ? 1 : ctx->n_pipes;

    // batch shape
    uint64_t cap = 64ull << 20;        // 64 Mi path slots per pipeline (~23 GB of path state each): few, large batches
    if (const char* e = getenv("BPT_MAX_SLOTS")) { uint64_t v = strtoull(e, nullptr, 10); if (v >= 1024) cap = v; }
    uint32_t S, rows_per_batch;
    uint64_t n_batches;
retry_shape:
    S = (uint32_t)std::min<uint64_t>(spp, std::max<uint64_t>(1, cap / rect_w));
    rows_per_batch = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(rect_h, cap / ((uint64_t)rect_w*S)));
    n_batches = (uint64_t)((spp + S - 1)/S) * ((rect_h + rows_per_batch - 1)/rows_per_batch);
    if (n_batches < (uint64_t)n_pipes) n_pipes = 1;
    if (n_pipes == 2 && ((rect_h + rows_per_batch - 1)/rows_per_batch) & 1) {
        // even out the last pair of batches
        uint32_t nb = (rect_h + rows_per_batch - 1)/rows_per_batch + 1;
        rows_per_batch = (rect_h + nb - 1)/nb;
    }
    uint64_t slots64 = (uint64_t)rect_w*rows_per_batch*S;
    if (slots64 > 0x7FFFFFFFull) { set_error("%s: batch too large", who); return BPT_ERR_ARG; }
    for (int p = 0; p < n_pipes; ++p) {
        int rc = ensure_state(ctx, &ctx->pipes[p], (uint32_t)slots64);
        if (rc) {
            // out of device memory for this batch size: halve the batch and retry (the film/scene stay resident)
            cudaGetLastError();
            for (auto& pp : ctx->pipes) { free_all(&pp.allocs); pp.max_slots = 0; }
            if (cap <= (1ull << 20)) return rc;
            cap >>= 1;
            goto retry_shape;
        }
    }

    uint64_t total_samples = (uint64_t)rect_w*rect_h*spp;
    if (want_records) {
        bpt_ctx::Pipe& pp = ctx->pipes[0];
        if (ctx->host_record_capacity < total_samples) { set_error("%s: record buffer too small", who); return BPT_ERR_ARG; }
        if (S != spp) { set_error("%s: records need all samples of a pixel in one batch (lower spp or raise BPT_MAX_SLOTS)", who); return BPT_ERR_UNSUPPORTED; }
        if (pp.d_record_capacity < slots64) {
            cudaFree(pp.d_records); pp.d_records = nullptr; pp.d_record_capacity = 0;
            CK(cudaMalloc((void**)&pp.d_records, slots64*sizeof(bpt_sample_record)));
            pp.d_record_capacity = slots64;
        }
    }

    if (ctx->row_map_capacity < rect_h) {
        CK(cudaStreamSynchronize(ctx->stream));
        for (auto& pp : ctx->pipes) CK(cudaStreamSynchronize(pp.stream));
        cudaFree(ctx->d_row_map); ctx->d_row_map = nullptr; ctx->row_map_capacity = 0;
        CK(cudaMalloc((void**)&ctx->d_row_map, (size_t)std::max<uint32_t>(rect_h, 4096)*sizeof(int32_t)));
        ctx->row_map_capacity = std::max<uint32_t>(rect_h, 4096);
    }
    // the previous pass may still be reading the old row map on the pipe streams: order the copy after them
    for (int p = 0; p < 2; ++p) { CK(cudaEventRecord(ctx->pipes[p].done, ctx->pipes[p].stream)); CK(cudaStreamWaitEvent(ctx->stream, ctx->pipes[p].done, 0)); }
    CK(cudaMemcpyAsync(ctx->d_row_map, rows.data(), (size_t)rect_h*sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));

    ctx->spans_used = 0;
    ctx->launches = 0; ctx->trace_launches = 0;
    CK(cudaEventRecord(ctx->pass_begin, ctx->stream));
    for (int p = 0; p < n_pipes; ++p) CK(cudaStreamWaitEvent(ctx->pipes[p].stream, ctx->pass_begin, 0));
    const bool stats = ctx->stats_enabled;
    uint32_t max_bounce = sc.settings.max_bounce_count;
    uint32_t batch_index = 0;

    for (uint32_t sa = 0; sa < spp; sa += S) {
        uint32_t Sb = std::min(S, spp - sa);
        for (uint32_t row = 0; row < rect_h; row += rows_per_batch, ++batch_index) {
            bpt_ctx::Pipe& pp = ctx->pipes[batch_index % n_pipes];
            cudaStream_t s = pp.stream;
            BatchDesc b;
            b.x0 = x0; b.row0 = row; b.row_map = ctx->d_row_map;
            b.rect_w = rect_w; b.rows = std::min(rows_per_batch, rect_h - row);
            b.sa = sa; b.S = Sb;
            b.frame_count = frame_count; b.salt = seed_salt;
            b.slots = rect_w*b.rows*Sb;
            b.want_records = want_records ? 1u : 0u;

            begin_span(ctx, ST_RAYGEN, s);
            k_raygen<<<grid_for(ctx, b.slots, 256, 8), 256, 0, s>>>(sc, pp.st, b);
            end_span(ctx, s);
            ctx->launches++;

            uint32_t* counters = pp.q.counters;
            for (uint32_t bounce = 0; bounce < max_bounce; ++bounce) {
                int in = bounce & 1, out = in ^ 1;
                // counters: [in] = active count for this bounce (bounce 0 uses the identity queue), [out] and [2] (shadow) reset
                const uint32_t* in_queue = bounce == 0 ? nullptr : pp.q.active[in];
                const uint32_t* in_count = bounce == 0 ? nullptr : counters + in;
                k_reset_counters<<<1, 32, 0, s>>>(counters, (1 << out) | (1 << 2) | (1 << 3) | (1 << 4));
                ctx->launches++;
                uint32_t work = b.slots;    // upper bound; kernels read the true count on the device

                begin_span(ctx, ST_TRACE, s);
                uint32_t tg = grid_for(ctx, work, 128, ctx->trace_ctas_per_sm);
                if (stats) k_trace_closest<true ><<<tg, 128, 0, s>>>(sc, pp.st, in_queue, in_count, b.slots, counters + 3, ctx->refill, ctx->d_stats);
                else       k_trace_closest<false><<<tg, 128, 0, s>>>(sc, pp.st, in_queue, in_count, b.slots, counters + 3, ctx->refill, ctx->d_stats);
                end_span(ctx, s);
                ctx->launches++; ctx->trace_launches++;

                begin_span(ctx, ST_SHADE, s);
                k_shade<<<grid_for(ctx, work, 128, 16), 128, 0, s>>>(sc, pp.st, b, bounce, in_queue, in_count, b.slots,
                                                                    pp.q.active[out], counters + out, pp.q.shadow, counters + 2, ctx->d_stats);
                end_span(ctx, s);
                ctx->launches++;

                begin_span(ctx, ST_SHADOW, s);
                if (stats) k_trace_shadow<true ><<<tg, 128, 0, s>>>(sc, pp.st, pp.q.shadow, counters + 2, counters + 4, ctx->refill, ctx->d_stats);
                else       k_trace_shadow<false><<<tg, 128, 0, s>>>(sc, pp.st, pp.q.shadow, counters + 2, counters + 4, ctx->refill, ctx->d_stats);
                end_span(ctx, s);
                ctx->launches++; ctx->trace_launches++;
            }

            begin_span(ctx, ST_SPLAT, s);
            uint32_t pixels = rect_w*b.rows;
            if (sc.filter_lut_size != 0 && sc.filter_radius == 2) {
                k_splat<2><<<grid_for(ctx, pixels, 128, 16), 128, 0, s>>>(sc, pp.st, b, ctx->film);
            } else {
                k_splat_generic<<<grid_for(ctx, b.slots, 128, 16), 128, 0, s>>>(sc, pp.st, b, ctx->film);
            }
            end_span(ctx, s);
            ctx->launches++;

            if (want_records) {
                k_write_records<<<grid_for(ctx, b.slots, 256, 8), 256, 0, s>>>(pp.st, b, pp.d_records);
                ctx->launches++;
                uint64_t first = (uint64_t)row*rect_w*spp;    // S == spp here: records are pixel-major / sample-minor
                CK(cudaMemcpyAsync(ctx->host_records + first, pp.d_records, (size_t)b.slots*sizeof(bpt_sample_record),
                                   cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                ctx->d2h_bytes += (uint64_t)b.slots*sizeof(bpt_sample_record);
            }
        }
    }
    for (int p = 0; p < n_pipes; ++p) {
        CK(cudaEventRecord(ctx->pipes[p].done, ctx->pipes[p].stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->pipes[p].done, 0));
    }
    CK(cudaEventRecord(ctx->pass_end, ctx->stream));
    ctx->pass_recorded = true;
    ctx->total_launches += ctx->launches;
    ctx->samples += total_samples;
    CK(cudaGetLastError());
    return BPT_OK;
}

int bpt_render_pass(bpt_ctx* ctx, int32_t x0, int32_t y0, int32_t x1, int32_t y1,
                    uint32_t frame_count, uint32_t spp, uint32_t seed_mode, uint32_t seed_salt) {
    if (!ctx) { set_error("bpt_render_pass: null ctx"); return BPT_ERR_ARG; }
    if (y0 < 0 || y0 >= y1 || y1 > (int32_t)ctx->film_h) {
        if (!ctx->scene_ready) { set_error("bpt_render_pass: no scene uploaded"); return BPT_ERR_STATE; }
        if (!ctx->film) { set_error("bpt_render_pass: no film (bpt_film_resize)"); return BPT_ERR_STATE; }
        set_error("bpt_render_pass: bad rect/spp"); return BPT_ERR_ARG;
    }
    std::vector<int32_t> rows((size_t)(y1 - y0));
    for (int32_t y = y0; y < y1; ++y) rows[(size_t)(y - y0)] = y;
    return render_rows(ctx, x0, x1, rows, frame_count, spp, seed_mode, seed_salt, "bpt_render_pass");
}

int bpt_render_pass_bands(bpt_ctx* ctx, int32_t x0, int32_t x1, uint32_t n_bands, const int32_t* y0y1,
                          uint32_t frame_count, uint32_t spp, uint32_t seed_mode, uint32_t seed_salt) {
    if (!ctx || !y0y1 || n_bands == 0) { set_error("bpt_render_pass_bands: null argument"); return BPT_ERR_ARG; }
    std::vector<int32_t> rows;
    for (uint32_t i = 0; i < n_bands; ++i) {
        int32_t a = y0y1[2*i], b = y0y1[2*i + 1];
        if (a < 0 || a >= b || b > (int32_t)ctx->film_h) { set_error("bpt_render_pass_bands: bad band %u", i); return BPT_ERR_ARG; }
        for (int32_t y = a; y < b; ++y) rows.push_back(y);
    }
    return render_rows(ctx, x0, x1, rows, frame_count, spp, seed_mode, seed_salt, "bpt_render_pass_bands");
}

int bpt_get_pass_timing(bpt_ctx* ctx, bpt_pass_timing* out) {
    if (!ctx || !out) { set_error("bpt_get_pass_timing: null argument"); return BPT_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    if (!ctx->pass_recorded) { set_error("bpt_get_pass_timing: no pass rendered"); return BPT_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->pass_end));
    CK(cudaEventElapsedTime(&out->total_ms, ctx->pass_begin, ctx->pass_end));
    float acc[ST_COUNT] = {0, 0, 0, 0, 0};
    for (size_t i = 0; i < ctx->spans_used; ++i) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->spans[i].a, ctx->spans[i].b) == cudaSuccess) acc[ctx->spans[i].stage] += ms;
    }
    out->raygen_ms = acc[ST_RAYGEN]; out->trace_ms = acc[ST_TRACE]; out->shade_ms = acc[ST_SHADE];
    out->shadow_ms = acc[ST_SHADOW]; out->splat_ms = acc[ST_SPLAT];
    out->kernel_launches = ctx->launches; out->trace_launches = ctx->trace_launches;
    return BPT_OK;
}

int bpt_resolve_bgra8(bpt_ctx* ctx, const bpt_post_settings* post, const uint8_t* dither_rgb8,
                      uint32_t dither_w, uint32_t dither_h, uint32_t* out_pixels) {
    if (!ctx || !post || !out_pixels) { set_error("bpt_resolve_bgra8: null argument"); return BPT_ERR_ARG; }
    if (!ctx->film) { set_error("bpt_resolve_bgra8: no film"); return BPT_ERR_STATE; }
    if (dither_rgb8 && (dither_w == 0 || dither_h == 0 || (dither_w & (dither_w - 1)) || (dither_h & (dither_h - 1)))) {
        set_error("bpt_resolve_bgra8: dither tile must have power-of-two width and height"); return BPT_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    size_t n = (size_t)ctx->film_w*ctx->film_h;
    uint32_t* d_out = nullptr; uint8_t* d_dither = nullptr;
    CK(cudaMalloc((void**)&d_out, n*sizeof(uint32_t)));
    if (dither_rgb8) {
        size_t db = (size_t)dither_w*dither_h*3;
        if (cudaMalloc((void**)&d_dither, db) != cudaSuccess) { cudaFree(d_out); set_error("bpt_resolve_bgra8: out of memory"); return BPT_ERR_CUDA; }
        cudaMemcpyAsync(d_dither, dither_rgb8, db, cudaMemcpyHostToDevice, ctx->stream);
        ctx->h2d_bytes += db;
    }
    k_resolve<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(ctx->film, ctx->film_w, ctx->film_h, *post, d_dither, dither_w, dither_h, d_out);
    ctx->total_launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_pixels, d_out, n*sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_out); cudaFree(d_dither);
    if (e != cudaSuccess) { set_error("bpt_resolve_bgra8: %s", cudaGetErrorString(e)); return BPT_ERR_CUDA; }
    ctx->d2h_bytes += n*sizeof(uint32_t);
    return BPT_OK;
}

int bpt_set_detailed_timing(bpt_ctx* ctx, int enable) {
    if (!ctx) return BPT_ERR_ARG;
    ctx->detailed_timing = enable != 0;
    return BPT_OK;
}

int bpt_get_transfer_bytes(bpt_ctx* ctx, uint64_t* h2d, uint64_t* d2h, int reset) {
    if (!ctx) return BPT_ERR_ARG;
    if (h2d) *h2d = ctx->h2d_bytes;
    if (d2h) *d2h = ctx->d2h_bytes;
    if (reset) { ctx->h2d_bytes = 0; ctx->d2h_bytes = 0; }
    return BPT_OK;
}

} // extern "C"
