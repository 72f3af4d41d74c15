// Device half of the C ABI (include/bpt.h section 2): context, scene flattening/upload, the per-pass wavefront
// schedule, diagnostics.  There is no CPU fallback anywhere in this file: without a usable CUDA device every entry
// point fails with BPT_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "host_scene.h"
#include "kernels.cuh"
#include "bvh_device.cuh"

using namespace bpt;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return BPT_ERR_CUDA; } } while (0)

namespace {

enum SceneSlot { SL_MATERIALS, SL_PRIMITIVES, SL_PLANES, SL_MESHES, SL_LIGHTS, SL_PAIRS, SL_TLAS_INDICES, SL_BIG_LEAVES,
                 SL_TRIANGLES, SL_TRI_ORIGINAL, SL_RAW_TRIANGLES, SL_NORMALS, SL_RAW_NORMALS, SL_SKYDOME, SL_RAW_SKYDOME,
                 SL_TRACE_RAYS, SL_TRACE_HITS, SL_TRACE_CURSOR, SL_RESOLVE_OUT, SL_RESOLVE_DITHER, SL_COUNT };

#define BPT_MAX_PIPES 4
enum Stage { ST_RAYGEN, ST_TRACE, ST_SHADE, ST_SHADOW, ST_SPLAT, ST_COUNT };

struct TimedSpan { cudaEvent_t a, b; int stage; };

} // namespace

struct bpt_ctx {
    int device = 0;
    int sm_count = 148;
    uint32_t refill = 16;                 // persistent warps fetch new rays once this many lanes are idle (or idle is the largest group); round 1: 12-24 +1.5 % over 8; round 2: 8-33 within 0.3 %
    int trace_ctas_per_sm = 8;            // resident CTAs of the persistent traversal kernels (occupancy query)
    cudaStream_t stream = nullptr;

    DScene sc{};                          // what the next pass renders: the active scene set + latched camera / settings
    bool scene_ready = false;
    bool tables_ready = false;
    struct Slot { void* p = nullptr; size_t cap = 0; };
    // The uploaded scene is double-buffered: bpt_upload_scene_async fills the set that is not being rendered, on a copy
    // stream, while passes over the active set are still running; the next pass flips to it.
    struct SceneSet {
        Slot slots[SL_COUNT];             // grow-only device buffers
        char* staging = nullptr;          // pinned host block for the small flattened tables
        size_t staging_capacity = 0;
        float* d_filter = nullptr;        // the reconstruction-filter LUT latched with this upload
        uint32_t* tri_original = nullptr; // DTriangle slot -> original triangle index (MeshBVH::indices)
        uint32_t stack_bound = 0;
        cudaEvent_t last_use = nullptr;   // behind the last pass / trace enqueued that reads this set
        cudaEvent_t uploaded = nullptr;   // behind the last upload into this set (the staging block is free again)
        bool used = false, staged = false;
    } sets[2];
    int active_set = 0;
    bool pending = false;                 // an uploaded scene waits in sets[1 - active_set] for the next pass to flip to it
    DScene pending_sc{};
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    Slot misc_slots[SL_COUNT];            // bpt_trace / bpt_resolve scratch (not part of a scene)
    uint32_t* tri_original = nullptr;     // of the active set
    float4* film_front = nullptr;         // snapshot of the film a bpt_download_film_async copies out from (the reference's front buffer, raytracer.cpp:705-709)
    uint32_t film_front_w = 0, film_front_h = 0;
    cudaEvent_t front_ready = nullptr, download_done = nullptr;
    bool download_pending = false;

    uint8_t *d_perm = nullptr, *d_sobol = nullptr, *d_scramble = nullptr, *d_rank = nullptr;

    uint32_t stack_bound = 0;             // TLAS depth + deepest BLAS of the active scene: most stack entries one ray can have pending

    float4* film = nullptr;
    float4* film_reduced = nullptr;       // root of a multi-GPU job: sum of all ranks' films (bpt_reduce_film)
    uint32_t film_reduced_w = 0, film_reduced_h = 0;
    bool film_owned = false;
    uint32_t film_w = 0, film_h = 0;

    // Two independent batch pipelines (stream + path state + queues): consecutive batches alternate between them so
    // the tail of one batch's persistent kernels overlaps the next batch's head.
    struct Pipe {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        uint32_t max_slots = 0, mstack_levels = 0;
        bool has_record_arrays = false;
        DPathState st{};
        DQueues q{};
        std::vector<void*> allocs;
        bpt_sample_record* d_records = nullptr;
        uint64_t d_record_capacity = 0;
    } pipes[BPT_MAX_PIPES];
    int n_pipes = 2;
    bool pipes_forced = false;            // BPT_PIPES given: no automatic widening for small batches
    uint64_t wide_pipes_budget = 32ull << 30;   // bytes of path state that four batch streams may hold together (see render_rows)
    uint32_t min_batches = 0;             // experiment knob (BPT_MIN_BATCHES): at least this many batches per pass
    uint32_t tail_threshold = 65536;      // paths: at or below this many survivors a batch finishes inside k_tail (0 = never)
    uint32_t shade_late_threads = BPT_SHADE_THREADS;   // block size of k_shade from the second bounce on (its block-wide sort couples the warps of a block)
    uint32_t tail_refill = 8;             // k_tail: idle lanes shade / start their next ray once this many wait (or they are the largest group)
    bool ray_prefilter = true;         // k_shade settles the shadow rays that never reach a BLAS (kernels.cuh shadow_tlas_head)
    bool merge_traces = true;
    uint32_t merge_max_slots = 16u << 20; // batches larger than this trace the two populations separately             // trace bounce b's extension rays and bounce b-1's shadow rays in one launch
    int32_t* d_row_map = nullptr;
    uint32_t row_map_capacity = 0;
    std::vector<int32_t> row_map_host;    // what d_row_map holds: a pass over the same rows as the last one uploads nothing
    // Consecutive passes overlap: the pipelines of pass k+1 start while pass k's last batches are still in their kernel
    // tails.  Only what touches the film is ordered across passes -- a k_splat of pass k+1 waits for `film_free`, recorded
    // on ctx->stream behind the join of pass k and behind every reader / clear of the film enqueued since.
    cudaEvent_t film_free = nullptr;
    bool pipes_need_setup = true;         // something the pipelines read was enqueued on ctx->stream (row map, scene flip): they wait for it once
    bool last_pass_single = false;
    // Batch -> stream round-robin continues across passes: consecutive one-batch passes alternate between the batch
    // streams, so pass k+1's full-machine launches run underneath the chain of small launches (each as long as its
    // longest ray) that ends pass k.  Making a batch wait until the one before it is past its bulk (an event behind the
    // shading of bounce 0 / 1 / 2) was measured and is not needed: the streams fall out of step by themselves, one
    // persistent kernel at a time (8-rank share of C2: 7.57 ms with or without; C2 full frame 54.0 without, 54.2-54.4 with).
    uint64_t batch_serial = 0;

    DStats* d_stats = nullptr;
    bool stats_enabled = false;
    uint32_t* h_error = nullptr;          // page-locked, device-mapped word the kernels raise (BPT_DEVERR_*), read after a sync

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

int device_slot(bpt_ctx::Slot* slots, int id, size_t bytes, void** out) {
    bpt_ctx::Slot& sl = slots[id];
    bytes = std::max<size_t>(bytes, 256);
    if (sl.cap < bytes) {
        if (sl.p) cudaFree(sl.p);
        sl.p = nullptr; sl.cap = 0;
        CK(cudaMalloc(&sl.p, bytes));
        sl.cap = bytes;
    }
    *out = sl.p;
    return BPT_OK;
}

void free_all(std::vector<void*>* v) {
    for (void* p : *v) cudaFree(p);
    v->clear();
}

// `levels`: material-stack levels beyond the implicit level 0 a path of this pass can reach (one push per bounce at most,
// 63 in all: integrators.cpp:602) -- the stack planes are 2 bytes x slots each, so a 12-bounce pass gets 12 of them, not 63.
// `records`: the two arrays only the per-sample records read.
int ensure_state(bpt_ctx* ctx, bpt_ctx::Pipe* pp, uint32_t slots, uint32_t levels, bool records) {
    (void)ctx;
    if (slots <= pp->max_slots && levels <= pp->mstack_levels && (!records || pp->has_record_arrays)) return BPT_OK;
    slots = std::max(slots, pp->max_slots); levels = std::max(levels, pp->mstack_levels); records = records || pp->has_record_arrays;
    free_all(&pp->allocs);
    pp->max_slots = 0; pp->mstack_levels = 0; pp->has_record_arrays = false;
    pp->st.primary_d = nullptr; pp->st.primary_o = nullptr;
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
    rc |= alloc((void**)&pp->st.mstack, n*2*std::max<uint32_t>(levels, 1));     // level 0 ("air") is implicit
    if (records) {
        rc |= alloc((void**)&pp->st.primary_d, n*16);
        rc |= alloc((void**)&pp->st.primary_o, n*16);
    }
    rc |= alloc((void**)&pp->q.active[0], n*4);
    rc |= alloc((void**)&pp->q.active[1], n*4);
    rc |= alloc((void**)&pp->q.shadow, n*sizeof(DShadowItem));
    rc |= alloc((void**)&pp->q.counters, 256);
    if (rc) return BPT_ERR_CUDA;
    pp->max_slots = slots; pp->mstack_levels = levels; pp->has_record_arrays = records;
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

// BPT_DEBUG_SYNC=1: synchronise after every launch of a pass and say which kernel it was (finding a hung launch)
void debug_sync(const char* what, uint32_t bounce, cudaStream_t s) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("BPT_DEBUG_SYNC"); on = (e && atoi(e)) ? 1 : 0; }
    if (!on) return;
    fprintf(stderr, "launch %s bounce %u ...", what, bounce); fflush(stderr);
    cudaError_t e = cudaStreamSynchronize(s);
    fprintf(stderr, " %s\n", cudaGetErrorString(e)); fflush(stderr);
}

uint32_t grid_for(const bpt_ctx* ctx, uint64_t work, uint32_t threads, uint32_t ctas_per_sm) {
    uint64_t need = (work + threads - 1)/threads;
    uint64_t cap = (uint64_t)ctx->sm_count*ctas_per_sm;
    return (uint32_t)std::max<uint64_t>(1, std::min(need, cap));
}

void fill_rows(float4* dst, const bpt_m4x4& m) {
    for (int r = 0; r < 3; ++r) dst[r] = make_float4(m.e[r][0], m.e[r][1], m.e[r][2], m.e[r][3]);
}

void latch_settings(DScene* sc, const bpt_scene* scene) {
    // what render_all_tiles does when a render (re)starts (raytracer.cpp:711-720)
    bpt_camera cam = scene->new_camera;
    recompute_camera(&cam);
    sc->camera = cam;
    sc->settings = scene->new_settings;
    sc->filter_radius = scene->filter.kernel_size;
    sc->filter_lut_size = scene->filter.cache_size;
    memcpy(sc->top_sky, scene->top_sky_color, 12);
    memcpy(sc->bot_sky, scene->bot_sky_color, 12);
    memcpy(sc->ambient_light, scene->ambient_light, 12);
}

// the next pass (or trace) is about to be enqueued: switch to a scene that bpt_upload_scene_async left waiting
int flip_to_pending_scene(bpt_ctx* ctx) {
    if (!ctx->pending) return BPT_OK;
    int target = 1 - ctx->active_set;
    CK(cudaStreamWaitEvent(ctx->stream, ctx->sets[target].uploaded, 0));
    // the batch streams wait for the upload itself, not for ctx->stream (which carries the join of the previous pass): the
    // first batch over the new scene overlaps the last batch over the old one
    for (auto& pp : ctx->pipes) CK(cudaStreamWaitEvent(pp.stream, ctx->sets[target].uploaded, 0));
    DScene n = ctx->pending_sc;
    n.strata_perm = ctx->sc.strata_perm; n.bn_sobol = ctx->sc.bn_sobol; n.bn_scramble = ctx->sc.bn_scramble; n.bn_rank = ctx->sc.bn_rank;
    n.error_flag = ctx->sc.error_flag;
    n.film_w = ctx->sc.film_w; n.film_h = ctx->sc.film_h;
    ctx->sc = n;
    ctx->active_set = target;
    ctx->tri_original = ctx->sets[target].tri_original;
    ctx->stack_bound = ctx->sets[target].stack_bound;
    ctx->pending = false;
    ctx->scene_ready = true;
    return BPT_OK;
}

// a pass / trace over the active scene set has just been enqueued on ctx->stream
void mark_scene_use(bpt_ctx* ctx) {
    bpt_ctx::SceneSet& set = ctx->sets[ctx->active_set];
    cudaEventRecord(set.last_use, ctx->stream);
    set.used = true;
}

// what the kernels raised since the last check (the stream they ran on has been synchronised by the caller)
int device_error(bpt_ctx* ctx, const char* who) {
    uint32_t e = *(volatile uint32_t*)ctx->h_error;
    if (e == BPT_DEVERR_NONE) return BPT_OK;
    *ctx->h_error = BPT_DEVERR_NONE;
    if (e == BPT_DEVERR_STACK_OVERFLOW) {
        set_error("%s: a ray had more than %d far children pending (depth bound of this scene's BVHs: %u); the reference overruns "
                  "its node_stack[64] on such a scene (intersection.cpp:261, :445) -- results of this call are incomplete", who, BPT_STACK_DEPTH, ctx->stack_bound);
        return BPT_ERR_UNSUPPORTED;
    }
    set_error("%s: a traversal warp exceeded its scheduling trip limit (device error %u); results of this call are incomplete", who, e);
    return BPT_ERR_CUDA;
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
    struct Guard { bpt_ctx* c; ~Guard() { if (c) bpt_destroy(c); } } guard{ctx};      // a failing CK() below releases what exists so far
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
    CK(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    for (auto& set : ctx->sets) {
        CK(cudaMalloc((void**)&set.d_filter, 512*sizeof(float)));
        CK(cudaEventCreateWithFlags(&set.last_use, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&set.uploaded, cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&ctx->front_ready, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->download_done, cudaEventDisableTiming));
    CK(cudaHostAlloc((void**)&ctx->h_error, 64, cudaHostAllocMapped));
    *ctx->h_error = BPT_DEVERR_NONE;
    CK(cudaHostGetDevicePointer((void**)&ctx->sc.error_flag, ctx->h_error, 0));
    CK(cudaEventCreate(&ctx->pass_begin));
    CK(cudaEventCreateWithFlags(&ctx->film_free, cudaEventDisableTiming));
    CK(cudaEventCreate(&ctx->pass_end));
    {
        int a = 0, b = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_trace_closest<false>, 128, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_trace_shadow<false>, 128, 0);
        int m = std::min(a, b);
        if (m > 0) ctx->trace_ctas_per_sm = m;
    }
    if (const char* e = getenv("BPT_REFILL")) { int v = atoi(e); if (v >= 1 && v <= 33) ctx->refill = (uint32_t)v; }
    if (const char* e = getenv("BPT_SHADE_LATE_THREADS")) { int v = atoi(e); if (v == 32 || v == 64 || v == 128 || v == 256) ctx->shade_late_threads = (uint32_t)std::min(v, BPT_SHADE_THREADS); }
    if (const char* e = getenv("BPT_RAY_PREFILTER")) ctx->ray_prefilter = atoi(e) != 0;
    if (const char* e = getenv("BPT_MERGE_TRACES")) ctx->merge_traces = atoi(e) != 0;
    if (const char* e = getenv("BPT_MERGE_MAX_SLOTS")) { long long v = atoll(e); if (v >= 0 && v <= 0x7FFFFFFFll) ctx->merge_max_slots = (uint32_t)v; }
    if (const char* e = getenv("BPT_TAIL_REFILL")) { int v = atoi(e); if (v >= 1 && v <= 33) ctx->tail_refill = (uint32_t)v; }
    if (const char* e = getenv("BPT_TAIL_THRESHOLD")) { long v = atol(e); if (v >= 0 && v <= (1 << 22)) ctx->tail_threshold = (uint32_t)v; }
    if (const char* e = getenv("BPT_PIPES")) { int v = atoi(e); if (v >= 1 && v <= BPT_MAX_PIPES) { ctx->n_pipes = v; ctx->pipes_forced = true; } }
    if (const char* e = getenv("BPT_WIDE_PIPES_GB")) ctx->wide_pipes_budget = strtoull(e, nullptr, 10) << 30;
    if (const char* e = getenv("BPT_MIN_BATCHES")) { int v = atoi(e); if (v >= 1 && v <= 64) ctx->min_batches = (uint32_t)v; }
    if (const char* e = getenv("BPT_TRACE_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 32) ctx->trace_ctas_per_sm = v; }
    const char* dt = getenv("BPT_DETAILED_TIMING");
    ctx->detailed_timing = dt && atoi(dt) != 0;
    guard.c = nullptr;
    *out_ctx = ctx;
    return BPT_OK;
}

void bpt_destroy(bpt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->h2d_stream) cudaStreamSynchronize(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamSynchronize(ctx->d2h_stream);
    for (auto& set : ctx->sets) {
        for (auto& sl : set.slots) if (sl.p) cudaFree(sl.p);
        if (set.staging) cudaFreeHost(set.staging);
        cudaFree(set.d_filter);
        if (set.last_use) cudaEventDestroy(set.last_use);
        if (set.uploaded) cudaEventDestroy(set.uploaded);
    }
    for (auto& sl : ctx->misc_slots) if (sl.p) cudaFree(sl.p);
    cudaFree(ctx->film_front);
    if (ctx->front_ready) cudaEventDestroy(ctx->front_ready);
    if (ctx->download_done) cudaEventDestroy(ctx->download_done);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    for (auto& pp : ctx->pipes) {
        cudaStreamSynchronize(pp.stream);
        free_all(&pp.allocs);
        cudaFree(pp.d_records);
        cudaEventDestroy(pp.done);
        cudaStreamDestroy(pp.stream);
    }
    cudaFree(ctx->d_row_map);
    if (ctx->film_owned && ctx->film) cudaFree(ctx->film);
    cudaFree(ctx->film_reduced);
    cudaFree(ctx->d_stats);
    if (ctx->h_error) cudaFreeHost(ctx->h_error);
    cudaFree(ctx->d_perm); cudaFree(ctx->d_sobol); cudaFree(ctx->d_scramble); cudaFree(ctx->d_rank);
    for (auto& s : ctx->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    cudaEventDestroy(ctx->pass_begin); cudaEventDestroy(ctx->pass_end); cudaEventDestroy(ctx->film_free);
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
    { int rc = flip_to_pending_scene(ctx); if (rc) return rc; }
    latch_settings(&ctx->sc, scene);
    float* d_filter = ctx->sets[ctx->active_set].d_filter;
    CK(cudaMemcpyAsync(d_filter, scene->filter.cache, 512*sizeof(float), cudaMemcpyHostToDevice, ctx->stream));   // behind the passes that still read the old LUT
    ctx->h2d_bytes += 512*sizeof(float) + sizeof(bpt_camera) + sizeof(bpt_settings);
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->sc.filter_lut = d_filter;
    return BPT_OK;
}

int bpt_upload_scene_async(bpt_ctx* ctx, const bpt_scene* scene) {
    if (!ctx || !scene) { set_error("bpt_upload_scene: null argument"); return BPT_ERR_ARG; }
    if (!scene->has_tlas) { set_error("bpt_upload_scene: call bpt_create_scene_bvh first"); return BPT_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    // The set that is NOT being rendered receives the upload, on the copy stream: passes over the active set keep running.
    const int target = 1 - ctx->active_set;
    bpt_ctx::SceneSet& set = ctx->sets[target];
    cudaStream_t s = ctx->h2d_stream;
    if (set.staged) CK(cudaEventSynchronize(set.uploaded));            // its pinned staging block is about to be rewritten by the host
    if (set.used) CK(cudaStreamWaitEvent(s, set.last_use, 0));         // the last pass that read this set must be through
    ctx->pending = false;                                              // an earlier upload nobody rendered is replaced
    ctx->pending_sc = DScene{};
    DScene& sc = ctx->pending_sc;

    if (scene->materials.size() + 1 > 0xFFFF) { set_error("bpt_upload_scene: more than 65534 materials"); return BPT_ERR_UNSUPPORTED; }
    for (uint32_t l : scene->lights) {
        if (l >= scene->primitives.size()) { set_error("bpt_upload_scene: light id %u is not a primitive (emissive plane?)", l); return BPT_ERR_UNSUPPORTED; }
    }

    // Small tables are flattened on the host into one pinned staging block; the big arrays (BLAS nodes, leaf-ordered
    // triangles, indices, normals, skydome) already live in page-locked memory inside the host scene (PinnedVec) and go to
    // the device as they are; the 48-byte DTriangle records and float4 texels are then built by kernels on the device.
    size_t n_mats = scene->materials.size() + 1, n_prims = scene->primitives.size(), n_planes = scene->planes.size(),
           n_meshes = scene->meshes.size(), n_lights = scene->lights.size();
    size_t small_bytes = n_mats*sizeof(DMaterial) + n_prims*sizeof(DPrimitive) + n_planes*sizeof(DPlane) + n_meshes*sizeof(DMesh) + 1024;
    if (set.staging_capacity < small_bytes) {
        if (set.staging) cudaFreeHost(set.staging);
        set.staging = nullptr; set.staging_capacity = 0;
        CK(cudaHostAlloc((void**)&set.staging, small_bytes*2, cudaHostAllocDefault));
        set.staging_capacity = small_bytes*2;
    }
    char* stage = set.staging;
    auto carve = [&](size_t bytes) { char* p = stage; stage += (bytes + 255) & ~(size_t)255; return p; };
    DMaterial* mats = (DMaterial*)carve(n_mats*sizeof(DMaterial));
    DPrimitive* prims = (DPrimitive*)carve(n_prims*sizeof(DPrimitive));
    DPlane* planes = (DPlane*)carve(n_planes*sizeof(DPlane));
    DMesh* meshes = (DMesh*)carve(n_meshes*sizeof(DMesh));

    // materials (+ the integrator's local "air", integrators.cpp:597-599)
    memset(mats, 0, n_mats*sizeof(DMaterial));
    for (size_t i = 0; i + 1 < n_mats; ++i) memcpy(&mats[i], &scene->materials[i], sizeof(bpt_material));
    mats[n_mats - 1].ior = 1.0f;
    mats[n_mats - 1].is_participating_medium = 1;

    memset(prims, 0, n_prims*sizeof(DPrimitive));
    for (size_t i = 0; i < n_prims; ++i) {
        const HostPrimitive& hp = scene->primitives[i];
        DPrimitive& dp = prims[i];
        const bpt_m4x4inv& xf = hp.transform >= 0 ? scene->transforms[hp.transform] : identity_transform();
        fill_rows(dp.inv, xf.inverse);
        fill_rows(dp.fwd, xf.forward);
        dp.type = hp.type; dp.material = hp.material; dp.mesh = hp.mesh;
        dp.sphere_r = hp.sphere_r;
        memcpy(dp.box_r, hp.box_r, 12);
    }
    memset(planes, 0, n_planes*sizeof(DPlane));
    for (size_t i = 0; i < n_planes; ++i) {
        memcpy(planes[i].n, scene->planes[i].plane_n, 12);
        planes[i].d = scene->planes[i].plane_d;
        planes[i].material = scene->planes[i].material;
    }

    // mesh directory + totals; FMNMX slab-test precondition (trace.cuh make_ray): all node boxes inside 1e15.
    // The traversal stack holds at most one pending far child per level of descent: TLAS depth + deepest BLAS.
    if (!scene->tlas.wide.valid) { set_error("bpt_upload_scene: the scene BVH has no device layout (bpt_create_scene_bvh failed?)"); return BPT_ERR_STATE; }
    size_t total_pairs = scene->tlas.wide.pairs.size(), total_big = scene->tlas.wide.big_leaves.size(), total_tris = 0;
    uint32_t deepest_blas = 0;
    bool any_normals = false;
    bool tame = scene->tlas.max_abs_extent < 1e15f || scene->tlas.indices.empty();
    for (size_t mi = 0; mi < n_meshes; ++mi) {
        const HostMesh& m = scene->meshes[mi];
        if (!m.bvh.wide.valid) { set_error("bpt_upload_scene: mesh %zu has no device BVH layout", mi); return BPT_ERR_STATE; }
        memcpy(&meshes[mi].root_q0, &m.bvh.wide.root, sizeof(WChild));
        meshes[mi].pair_base = (uint32_t)total_pairs;
        meshes[mi].tri_base = (uint32_t)total_tris;
        meshes[mi].triangle_count = m.triangle_count;
        meshes[mi].has_normals = m.has_normals ? 1u : 0u;
        meshes[mi].big_base = (uint32_t)total_big;
        total_pairs += m.bvh.wide.pairs.size();
        total_big += m.bvh.wide.big_leaves.size();
        total_tris += m.triangle_count;
        any_normals |= m.has_normals;
        tame = tame && (m.bvh.max_abs_extent < 1e15f);
        deepest_blas = std::max(deepest_blas, m.bvh.wide.depth);
    }
    if (total_pairs > 0x7FFFFFFFull || total_tris > 0x7FFFFFFFull) { set_error("bpt_upload_scene: scene too large for 32-bit pair / triangle indices"); return BPT_ERR_UNSUPPORTED; }
    set.stack_bound = scene->tlas.wide.depth + deepest_blas;
    sc.tame_bounds = tame ? 1u : 0u;
    memcpy(&sc.tlas_root_q0, &scene->tlas.wide.root, sizeof(WChild));

    void* d = nullptr;
    #define SLOT(id, bytes) do { int rc_ = device_slot(set.slots, id, (bytes), &d); if (rc_) return rc_; } while (0)
    #define H2D(dst, src, bytes) do { if ((bytes) > 0) { CK(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, s)); ctx->h2d_bytes += (bytes); } } while (0)
    SLOT(SL_MATERIALS, n_mats*sizeof(DMaterial));   sc.materials = (const DMaterial*)d;   H2D(d, mats, n_mats*sizeof(DMaterial));
    SLOT(SL_PRIMITIVES, n_prims*sizeof(DPrimitive)); sc.primitives = (const DPrimitive*)d; H2D(d, prims, n_prims*sizeof(DPrimitive));
    SLOT(SL_PLANES, n_planes*sizeof(DPlane));       sc.planes = (const DPlane*)d;         H2D(d, planes, n_planes*sizeof(DPlane));
    SLOT(SL_MESHES, n_meshes*sizeof(DMesh));        sc.meshes = (const DMesh*)d;          H2D(d, meshes, n_meshes*sizeof(DMesh));
    SLOT(SL_LIGHTS, n_lights*sizeof(uint32_t));     sc.lights = (const uint32_t*)d;       H2D(d, scene->lights.data(), n_lights*sizeof(uint32_t));
    SLOT(SL_TLAS_INDICES, scene->tlas.indices.size()*sizeof(uint32_t)); sc.tlas_indices = (const uint32_t*)d;
    H2D(d, scene->tlas.indices.data(), scene->tlas.indices.size()*sizeof(uint32_t));

    // pair records: the TLAS's first (its refs are relative to pair 0), then every BLAS's at its pair_base
    SLOT(SL_PAIRS, total_pairs*sizeof(DPair));             DPair* d_pairs = (DPair*)d;              sc.pairs = d_pairs;
    SLOT(SL_BIG_LEAVES, total_big*sizeof(uint2));          uint2* d_big = (uint2*)d;                sc.big_leaves = d_big;
    SLOT(SL_TRIANGLES, total_tris*sizeof(DTriangle));      DTriangle* d_tris = (DTriangle*)d;       sc.triangles = d_tris;
    SLOT(SL_TRI_ORIGINAL, total_tris*sizeof(uint32_t));    uint32_t* d_orig = (uint32_t*)d;         set.tri_original = d_orig;
    SLOT(SL_RAW_TRIANGLES, total_tris*9*sizeof(float));    float* d_raw = (float*)d;
    float4* d_normals = nullptr; float* d_raw_normals = nullptr;
    sc.normals = nullptr;
    if (any_normals) {
        SLOT(SL_NORMALS, total_tris*3*sizeof(float4));     d_normals = (float4*)d; sc.normals = d_normals;
        SLOT(SL_RAW_NORMALS, total_tris*9*sizeof(float));  d_raw_normals = (float*)d;
    }
    H2D(d_pairs, scene->tlas.wide.pairs.data(), scene->tlas.wide.pairs.size()*sizeof(WPair));
    H2D(d_big, scene->tlas.wide.big_leaves.data(), scene->tlas.wide.big_leaves.size()*sizeof(WBigLeaf));
    for (size_t mi = 0; mi < n_meshes; ++mi) {
        const HostMesh& m = scene->meshes[mi];
        size_t pb = meshes[mi].pair_base, tb = meshes[mi].tri_base, bb = meshes[mi].big_base, nt = m.triangle_count;
        H2D(d_pairs + pb, m.bvh.wide.pairs.data(), m.bvh.wide.pairs.size()*sizeof(WPair));
        H2D(d_big + bb, m.bvh.wide.big_leaves.data(), m.bvh.wide.big_leaves.size()*sizeof(WBigLeaf));
        H2D(d_raw + tb*9, m.leaf_triangles.data(), nt*9*sizeof(float));
        H2D(d_orig + tb, m.bvh.indices.data(), nt*sizeof(uint32_t));
        if (m.has_normals) H2D(d_raw_normals + tb*9, m.normals.data(), nt*9*sizeof(float));
        k_build_triangles<<<grid_for(ctx, nt, 256, 8), 256, 0, s>>>(d_raw + tb*9, d_orig + tb, (uint32_t)nt, d_tris + tb,
                                                                    m.has_normals ? d_raw_normals + tb*9 : nullptr,
                                                                    d_normals ? d_normals + tb*3 : nullptr);
        ctx->total_launches += 1;
    }

    sc.skydome = nullptr; sc.skydome_w = sc.skydome_h = 0;
    if (!scene->skydome.empty()) {
        size_t texels = (size_t)scene->skydome_w*scene->skydome_h;
        SLOT(SL_RAW_SKYDOME, texels*3*sizeof(float)); float* d_raw_sky = (float*)d;
        SLOT(SL_SKYDOME, texels*sizeof(float4));      float4* d_sky = (float4*)d;
        H2D(d_raw_sky, scene->skydome.data(), texels*3*sizeof(float));
        k_expand_rgb<<<grid_for(ctx, texels, 256, 8), 256, 0, s>>>(d_raw_sky, (uint32_t)texels, d_sky);
        ctx->total_launches += 1;
        sc.skydome = d_sky;
        sc.skydome_w = scene->skydome_w; sc.skydome_h = scene->skydome_h;
    }
    #undef SLOT
    #undef H2D
    sc.plane_count = (uint32_t)n_planes;
    sc.primitive_count = (uint32_t)n_prims;
    sc.material_count = (uint32_t)scene->materials.size();
    sc.air_material = sc.material_count;
    sc.light_count = (uint32_t)n_lights;

    // camera / settings / filter latch with the upload (render_all_tiles :711-720)
    latch_settings(&sc, scene);
    CK(cudaMemcpyAsync(set.d_filter, scene->filter.cache, 512*sizeof(float), cudaMemcpyHostToDevice, s));
    ctx->h2d_bytes += 512*sizeof(float) + sizeof(bpt_camera) + sizeof(bpt_settings);
    sc.filter_lut = set.d_filter;
    CK(cudaEventRecord(set.uploaded, s));
    CK(cudaGetLastError());
    set.staged = true;
    ctx->pending = true;
    return BPT_OK;
}

int bpt_upload_scene(bpt_ctx* ctx, const bpt_scene* scene) {
    int rc = bpt_upload_scene_async(ctx, scene);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->h2d_stream));      // the synchronous form: on return the scene is on the device and active
    CK(cudaGetLastError());
    return flip_to_pending_scene(ctx);
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
    return device_error(ctx, "bpt_download_film");
}

// Snapshot + asynchronous read-back.  The film is copied device-to-device into a front buffer behind everything enqueued so
// far (the reference's "copy back buffer to front buffer", raytracer.cpp:705-709), and the front buffer goes to the host on
// a copy stream: the next pass can clear / accumulate into the film while the previous image is still on its way out.
int bpt_download_film_async(bpt_ctx* ctx, float* out, int reduced) {
    if (!ctx || !out) { set_error("bpt_download_film_async: null argument"); return BPT_ERR_ARG; }
    const float4* src = reduced ? ctx->film_reduced : ctx->film;
    if (!src) { set_error("bpt_download_film_async: no %sfilm", reduced ? "reduced " : ""); return BPT_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    size_t bytes = (size_t)ctx->film_w*ctx->film_h*sizeof(float4);
    if (!ctx->film_front || ctx->film_front_w != ctx->film_w || ctx->film_front_h != ctx->film_h) {
        CK(cudaStreamSynchronize(ctx->d2h_stream));
        cudaFree(ctx->film_front); ctx->film_front = nullptr;
        CK(cudaMalloc((void**)&ctx->film_front, bytes));
        ctx->film_front_w = ctx->film_w; ctx->film_front_h = ctx->film_h;
        ctx->download_pending = false;
    }
    if (ctx->download_pending) CK(cudaStreamWaitEvent(ctx->stream, ctx->download_done, 0));   // the previous read-back still reads the front buffer
    CK(cudaMemcpyAsync(ctx->film_front, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->front_ready, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->d2h_stream, ctx->front_ready, 0));
    CK(cudaMemcpyAsync(out, ctx->film_front, bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));    // `out` should be page-locked (bpt_host_register)
    CK(cudaEventRecord(ctx->download_done, ctx->d2h_stream));
    ctx->download_pending = true;
    ctx->d2h_bytes += bytes;
    return BPT_OK;
}

int bpt_wait_download(bpt_ctx* ctx) {
    if (!ctx) { set_error("bpt_wait_download: null ctx"); return BPT_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    if (ctx->download_pending) CK(cudaEventSynchronize(ctx->download_done));
    return device_error(ctx, "bpt_wait_download");
}

int bpt_host_register(void* host_ptr, size_t bytes) {
    if (!host_ptr || bytes == 0) { set_error("bpt_host_register: bad arguments"); return BPT_ERR_ARG; }
    CK(cudaHostRegister(host_ptr, bytes, cudaHostRegisterDefault));
    return BPT_OK;
}

int bpt_host_unregister(void* host_ptr) {
    if (!host_ptr) return BPT_OK;
    CK(cudaHostUnregister(host_ptr));
    return BPT_OK;
}

int bpt_sync(bpt_ctx* ctx) {
    if (!ctx) { set_error("bpt_sync: null ctx"); return BPT_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return device_error(ctx, "bpt_sync");
}

int bpt_stats_enable(bpt_ctx* ctx, int enable) {
    if (!ctx) return BPT_ERR_ARG;
    ctx->stats_enabled = enable != 0;
    return BPT_OK;
}

int bpt_set_ray_prefilter(bpt_ctx* ctx, int enable) {
    if (!ctx) return BPT_ERR_ARG;
    ctx->ray_prefilter = enable != 0;
    return BPT_OK;
}

int bpt_get_ray_prefilter_stats(bpt_ctx* ctx, uint64_t out[2]) {
    if (!ctx || !out) { set_error("bpt_get_ray_prefilter_stats: null argument"); return BPT_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    DStats h;
    CK(cudaMemcpy(&h, ctx->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
    out[0] = h.v[17]; out[1] = h.v[18];
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
    { int rc_ = flip_to_pending_scene(ctx); if (rc_) return rc_; }
    if (!ctx->scene_ready) { set_error("bpt_trace: no scene uploaded"); return BPT_ERR_STATE; }
    if (mode != BPT_TRACE_CLOSEST && mode != BPT_TRACE_OCCLUSION) { set_error("bpt_trace: bad mode"); return BPT_ERR_ARG; }
    if (n == 0) return BPT_OK;
    CK(cudaSetDevice(ctx->device));
    // grow-only device buffers owned by the context (no allocation per call, nothing to leak on an error path)
    void* d = nullptr;
    int rc = device_slot(ctx->misc_slots, SL_TRACE_RAYS, (size_t)n*sizeof(bpt_ray), &d);   if (rc) return rc;  bpt_ray* d_rays = (bpt_ray*)d;
    rc = device_slot(ctx->misc_slots, SL_TRACE_HITS, (size_t)n*sizeof(bpt_hit), &d);       if (rc) return rc;  bpt_hit* d_hits = (bpt_hit*)d;
    rc = device_slot(ctx->misc_slots, SL_TRACE_CURSOR, 256, &d);                           if (rc) return rc;  uint32_t* d_cursor = (uint32_t*)d;
    CK(cudaMemcpyAsync(d_rays, rays, (size_t)n*sizeof(bpt_ray), cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += (uint64_t)n*sizeof(bpt_ray); ctx->d2h_bytes += (uint64_t)n*sizeof(bpt_hit);
    CK(cudaMemsetAsync(d_cursor, 0, 256, ctx->stream));
    uint32_t grid = grid_for(ctx, n, BPT_TRACE_THREADS, ctx->trace_ctas_per_sm);
    bool st = ctx->stats_enabled;
    if (mode == BPT_TRACE_CLOSEST) {
        if (st) k_trace_api<false, true ><<<grid, BPT_TRACE_THREADS, 0, ctx->stream>>>(ctx->sc, d_rays, n, ignored, d_hits, ctx->tri_original, d_cursor, ctx->refill, ctx->d_stats);
        else    k_trace_api<false, false><<<grid, BPT_TRACE_THREADS, 0, ctx->stream>>>(ctx->sc, d_rays, n, ignored, d_hits, ctx->tri_original, d_cursor, ctx->refill, ctx->d_stats);
    } else {
        if (st) k_trace_api<true, true ><<<grid, BPT_TRACE_THREADS, 0, ctx->stream>>>(ctx->sc, d_rays, n, ignored, d_hits, ctx->tri_original, d_cursor, ctx->refill, ctx->d_stats);
        else    k_trace_api<true, false><<<grid, BPT_TRACE_THREADS, 0, ctx->stream>>>(ctx->sc, d_rays, n, ignored, d_hits, ctx->tri_original, d_cursor, ctx->refill, ctx->d_stats);
    }
    ctx->total_launches += 1;
    mark_scene_use(ctx);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_hits, (size_t)n*sizeof(bpt_hit), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return device_error(ctx, "bpt_trace");
}

static int render_rows(bpt_ctx* ctx, int32_t x0, int32_t x1, const std::vector<int32_t>& rows,
                       uint32_t frame_count, uint32_t spp, uint32_t seed_mode, uint32_t seed_salt, const char* who) {
    { int rc_ = flip_to_pending_scene(ctx); if (rc_) return rc_; }     // a scene uploaded asynchronously becomes the active one here
    if (!ctx->scene_ready) { set_error("%s: no scene uploaded", who); return BPT_ERR_STATE; }
    if (!ctx->tables_ready) { set_error("%s: sampler tables not set (bpt_set_sampler_tables)", who); return BPT_ERR_STATE; }
    if (!ctx->film) { set_error("%s: no film (bpt_film_resize)", who); return BPT_ERR_STATE; }
    if (seed_mode != BPT_SEED_PER_PIXEL) { set_error("%s: unknown seed mode", who); return BPT_ERR_ARG; }
    if (x0 < 0 || x1 > (int32_t)ctx->film_w || x0 >= x1 || rows.empty() || spp == 0) { set_error("%s: bad rect/spp", who); return BPT_ERR_ARG; }
    for (int32_t y : rows) if (y < 0 || y >= (int32_t)ctx->film_h) { set_error("%s: bad rect/spp (row %d)", who, y); return BPT_ERR_ARG; }
    const bool recursive = ctx->sc.settings.integrator == BPT_INTEGRATOR_WHITTED || ctx->sc.settings.integrator == BPT_INTEGRATOR_GT_RECURSIVE;
    if (recursive && ctx->sc.settings.max_bounce_count > BPT_MAX_RECURSION) {
        set_error("%s: the recursive integrators support max_bounce_count <= %d", who, BPT_MAX_RECURSION);
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
    // 128 Mi path slots per pipeline (~30 GB of path state each at 12 bounces): few, large batches.  Every launch of a batch ends with
    // the machine draining behind its longest rays, so fewer, larger launches win: pass period of C2 / C3 at 64 spp / C4 at 64 spp with a
    // cap of 32 / 64 / 128 Mi slots: 54.2 / 52.4 / 50.7 ms, - / 107.6 / 105.6, - / 51.5 / 50.4.  Out of device memory halves the cap (below).
    uint64_t cap = 128ull << 20;
    if (const char* e = getenv("BPT_MAX_SLOTS")) { uint64_t v = strtoull(e, nullptr, 10); if (v >= 1024) cap = v; }
    uint32_t S, rows_per_batch;
    uint64_t n_batches;
    const int n_pipes_wanted = n_pipes;
    // The caller enqueues passes back to back (the previous one has not finished): a pass that fits one batch stays one
    // batch -- half the launches, half the kernel tails -- and consecutive passes alternate between the batch streams.
    // A caller that waits for every pass gets the pass split over the streams instead, so that they overlap inside it.
    bool back_to_back = false;
    if (n_pipes > 1 && ctx->pass_recorded && !ctx->last_pass_single) {
        back_to_back = cudaEventQuery(ctx->pass_end) == cudaErrorNotReady;
        cudaGetLastError();
    }
    if (const char* e = getenv("BPT_BACK_TO_BACK")) back_to_back = atoi(e) != 0 && n_pipes > 1;
retry_shape:
    n_pipes = n_pipes_wanted;             // a retry changes the batch count, and with it how many pipelines can be fed
    S = (uint32_t)std::min<uint64_t>(spp, std::max<uint64_t>(1, cap / rect_w));
    rows_per_batch = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(rect_h, cap / ((uint64_t)rect_w*S)));
    n_batches = (uint64_t)((spp + S - 1)/S) * ((rect_h + rows_per_batch - 1)/rows_per_batch);
    {
        // fewer batches than pipelines cannot overlap: split the rows so that every pipeline gets a batch and the
        // pipelines hide each other's kernel tails (only worth it when each batch still fills the machine)
        uint64_t row_batches = (rect_h + rows_per_batch - 1)/rows_per_batch;
        uint64_t want = std::max<uint64_t>(ctx->min_batches, back_to_back ? 1u : (uint64_t)n_pipes);
        uint64_t total = (uint64_t)rect_w*rect_h*spp;
        while (want > 1 && total/want < (1ull << 20) && want > ctx->min_batches) --want;
        if (S == spp && row_batches < want && rect_h >= want) {
            rows_per_batch = (uint32_t)((rect_h + want - 1)/want);
            n_batches = (rect_h + rows_per_batch - 1)/rows_per_batch;
        }
    }
    if (n_batches < (uint64_t)n_pipes && !back_to_back) n_pipes = (int)n_batches;
    if (n_pipes >= 2 && n_batches > 1) {
        // even out the batches of the last round over the pipelines
        uint32_t nb = (rect_h + rows_per_batch - 1)/rows_per_batch;
        uint32_t rounded = ((nb + n_pipes - 1)/n_pipes)*n_pipes;
        if (rounded != nb && rect_h >= rounded) rows_per_batch = (rect_h + rounded - 1)/rounded;
    }
    uint64_t slots64 = (uint64_t)rect_w*rows_per_batch*S;
    if (slots64 > 0x7FFFFFFFull) { set_error("%s: batch too large", who); return BPT_ERR_ARG; }
    // One-batch passes enqueued back to back (a rank's share of a multi-GPU frame) go round BPT_MAX_PIPES streams instead of two
    // when the path state of that many batches stays within a budget (32 GB): more passes in flight under each other's chains of
    // small launches.  Pass period of an 8 / 4-rank share of C2 on one GPU with 2 -> 4 streams: 7.36 -> 7.11, 13.78 -> 13.43 ms; on
    // 8 / 4 GPUs 7.73 -> 7.49, 14.06 -> 13.93 ms.  Larger batches stay on two streams: a 2-rank share gained 2 % on one GPU and
    // nothing on two GPUs for 59 GB of path state, the whole frame 1.2 % for 123 GB.
    {
        const uint64_t levels = std::min<uint32_t>(BPT_MATERIAL_STACK_DEPTH - 1, std::max<uint32_t>(1, ctx->sc.settings.max_bounce_count));
        const uint64_t state_bytes = slots64*(197ull + 2ull*levels);          // ensure_state: the arrays of one pipeline
        if (back_to_back && n_batches == 1 && !ctx->pipes_forced && (uint64_t)BPT_MAX_PIPES*state_bytes <= ctx->wide_pipes_budget) n_pipes = BPT_MAX_PIPES;
    }
    for (int p = 0; p < n_pipes; ++p) {
        const uint32_t levels = std::min<uint32_t>(BPT_MATERIAL_STACK_DEPTH - 1, std::max<uint32_t>(1, ctx->sc.settings.max_bounce_count));
        int rc = ensure_state(ctx, &ctx->pipes[p], (uint32_t)slots64, levels, want_records);
        if (rc) {
            // out of device memory for this batch size: halve the batch and retry (the film/scene stay resident)
            cudaGetLastError();
            for (auto& pp : ctx->pipes) { free_all(&pp.allocs); pp.max_slots = 0; pp.mstack_levels = 0; pp.has_record_arrays = false; }
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

    // The pipelines read: the scene set (flip_to_pending_scene), the row map, the sampler tables -- all enqueued on
    // ctx->stream.  A pass that finds them unchanged does not wait for ctx->stream at all, so its first batches overlap
    // the kernel tails of the previous pass (ctx->stream carries that pass's join).  The one-batch-at-a-time mode
    // (per-stage timing, records) keeps the full join on both sides so that its spans measure one kernel each.
    const bool single = ctx->detailed_timing || want_records;
    if (ctx->row_map_capacity < rect_h) {
        CK(cudaStreamSynchronize(ctx->stream));
        for (auto& pp : ctx->pipes) CK(cudaStreamSynchronize(pp.stream));
        cudaFree(ctx->d_row_map); ctx->d_row_map = nullptr; ctx->row_map_capacity = 0;
        CK(cudaMalloc((void**)&ctx->d_row_map, (size_t)std::max<uint32_t>(rect_h, 4096)*sizeof(int32_t)));
        ctx->row_map_capacity = std::max<uint32_t>(rect_h, 4096);
        ctx->row_map_host.clear();
    }
    if (ctx->row_map_host != rows) {
        // the previous pass may still be reading the old row map on the pipe streams: order the copy after them
        for (int p = 0; p < BPT_MAX_PIPES; ++p) { CK(cudaEventRecord(ctx->pipes[p].done, ctx->pipes[p].stream)); CK(cudaStreamWaitEvent(ctx->stream, ctx->pipes[p].done, 0)); }
        ctx->row_map_host = rows;         // the async copy below reads this vector: it stays untouched until the next change, which joins first
        CK(cudaMemcpyAsync(ctx->d_row_map, ctx->row_map_host.data(), (size_t)rect_h*sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        ctx->pipes_need_setup = true;
    }

    ctx->spans_used = 0;
    ctx->launches = 0; ctx->trace_launches = 0;
    CK(cudaEventRecord(ctx->pass_begin, ctx->stream));      // with overlapping passes: the end of the previous pass, so total_ms is the pass period
    if (ctx->pipes_need_setup || single || ctx->last_pass_single) {
        for (int p = 0; p < BPT_MAX_PIPES; ++p) CK(cudaStreamWaitEvent(ctx->pipes[p].stream, ctx->pass_begin, 0));
        ctx->pipes_need_setup = false;
    }
    ctx->last_pass_single = single;
    CK(cudaEventRecord(ctx->film_free, ctx->stream));
    const bool stats = ctx->stats_enabled;
    // the shadow-ray prefilter runs where it can decide: a TLAS that is one leaf (planes and the root test alone settle too few rays
    // to pay for its registers in k_shade: measured +0.7 ms on C3 / C4).  The counting pass traces every ray in the traversal kernels.
    uint32_t root_ref; memcpy(&root_ref, &sc.tlas_root_q1.z, 4);
    const bool prefilter = ctx->ray_prefilter && (root_ref & BPT_WREF_LEAF) != 0u && sc.settings.integrator == BPT_INTEGRATOR_ADVANCED;
    sc.prefilter = !prefilter ? 0u : (stats ? 2u : 1u);
    uint32_t max_bounce = sc.settings.max_bounce_count;
    if (sc.settings.integrator == BPT_INTEGRATOR_NORMALS || sc.settings.integrator == BPT_INTEGRATOR_DISTANCES) max_bounce = 1;   // one intersect_scene per sample
    uint32_t batch_index = 0;

    for (uint32_t sa = 0; sa < spp; sa += S) {
        uint32_t Sb = std::min(S, spp - sa);
        for (uint32_t row = 0; row < rect_h; row += rows_per_batch, ++batch_index) {
            const int pipe_index = single ? 0 : (int)(ctx->batch_serial++ % (uint64_t)n_pipes);
            bpt_ctx::Pipe& pp = ctx->pipes[pipe_index];
            cudaStream_t s = pp.stream;
            BatchDesc b;
            b.x0 = x0; b.row0 = row; b.row_map = ctx->d_row_map;
            b.rect_w = rect_w; b.rows = std::min(rows_per_batch, rect_h - row);
            b.sa = sa; b.S = Sb;
            b.frame_count = frame_count; b.salt = seed_salt;
            b.slots = rect_w*b.rows*Sb;
            b.want_records = want_records ? 1u : 0u;
            set_batch_magic(b);

            begin_span(ctx, ST_RAYGEN, s);
            k_raygen<<<grid_for(ctx, b.slots, 256, 8), 256, 0, s>>>(sc, pp.st, b);
            end_span(ctx, s);
            ctx->launches++;

            if (recursive) {
                // "Whitted" / "Ground Truth Recursive": every sample is evaluated to completion by one launch (recursive.cuh)
                begin_span(ctx, ST_TRACE, s);
                uint32_t rg = grid_for(ctx, b.slots, 128, 4);
                if (sc.settings.integrator == BPT_INTEGRATOR_WHITTED) k_recursive<true ><<<rg, 128, 0, s>>>(sc, pp.st, b, ctx->tail_refill, ctx->d_stats);
                else                                                  k_recursive<false><<<rg, 128, 0, s>>>(sc, pp.st, b, ctx->tail_refill, ctx->d_stats);
                debug_sync("k_recursive", 0, s);
                end_span(ctx, s);
                ctx->launches++; ctx->trace_launches++;
                max_bounce = 0;
            }
            if (max_bounce == 0 && !recursive) CK(cudaMemsetAsync(pp.st.radiance, 0, (size_t)b.slots*sizeof(float4), s));   // no bounce ever writes it (k_raygen leaves it to the first shade)
            uint32_t* counters = pp.q.counters;
            // Per bounce b:  trace { extension rays of b  +  shadow rays queued by bounce b-1 }  ->  shade b.
            // counters: [0]/[1] = active-queue sizes (ping-pong), [2] = shadow count, [3] = fetch cursor.
            // With the counting instantiations (stats) the two populations are traced by separate launches.
            // Per-kernel timing: ST_TRACE spans k_trace_closest (bounce 0) + k_trace_merged, ST_SHADOW the final k_trace_shadow.
            // Merging pays where kernel tails dominate (small batches: +14 % on an 8-rank share of C2) and costs where
            // they do not (C4 at 64 Mi slots: -15 %, the mixed population diverges more): merge below a batch size.
            const bool merged = !stats && ctx->merge_traces && b.slots <= ctx->merge_max_slots;
            const bool tail = !stats && ctx->tail_threshold > 0;
            if (max_bounce > 0) { k_reset_counters<<<1, 32, 0, s>>>(counters, (1 << 3) | (1 << 4)); ctx->launches++; }   // fetch cursors; later bounces: k_bounce_prepare
            for (uint32_t bounce = 0; bounce < max_bounce; ++bounce) {
                int in = bounce & 1, out = in ^ 1;
                const uint32_t* in_queue = bounce == 0 ? nullptr : pp.q.active[in];
                const uint32_t* in_count = bounce == 0 ? nullptr : counters + in;
                uint32_t work = b.slots;    // upper bound; kernels read the true counts on the device
                uint32_t tg = grid_for(ctx, work, 128, ctx->trace_ctas_per_sm);

                begin_span(ctx, ST_TRACE, s);
                if (merged && bounce > 0) {
                    k_trace_merged<<<tg, 128, 0, s>>>(sc, pp.st, in_queue, in_count, pp.q.shadow, counters + 2, counters + 3, ctx->refill);
                    debug_sync("k_trace_merged", bounce, s);
                } else {
                    if (stats) k_trace_closest<true ><<<tg, 128, 0, s>>>(sc, pp.st, in_queue, in_count, b.slots, counters + 3, ctx->refill, ctx->d_stats);
                    else       k_trace_closest<false><<<tg, 128, 0, s>>>(sc, pp.st, in_queue, in_count, b.slots, counters + 3, ctx->refill, ctx->d_stats);
                    debug_sync("k_trace_closest", bounce, s);
                }
                end_span(ctx, s);
                ctx->launches++; ctx->trace_launches++;

                // queue bookkeeping for the shade below, and -- few survivors -- the hand-over to k_tail, which finishes them
                // inside one launch; the wavefront launches of the remaining bounces then find empty queues
                k_bounce_prepare<<<1, 32, 0, s>>>(counters, in, out, (tail && bounce > 0) ? ctx->tail_threshold : 0u);
                ctx->launches++;
                if (tail && bounce > 0) {
                    begin_span(ctx, ST_TRACE, s);
                    k_tail<<<(ctx->tail_threshold + 127)/128, 128, 0, s>>>(sc, pp.st, b, bounce, pp.q.active[in], counters + 8, ctx->tail_refill, ctx->d_stats);
                    debug_sync("k_tail", bounce, s);
                    end_span(ctx, s);
                    ctx->launches++; ctx->trace_launches++;
                }
                begin_span(ctx, ST_SHADE, s);
                const uint32_t sh_threads = bounce == 0 ? BPT_SHADE_THREADS : ctx->shade_late_threads;
                const uint32_t sg = grid_for(ctx, work, sh_threads, BPT_SHADE_MIN_CTAS*BPT_SHADE_THREADS/sh_threads);
                if (prefilter) k_shade<true ><<<sg, sh_threads, 0, s>>>(sc, pp.st, b, bounce, in_queue, in_count, b.slots, pp.q.active[out], counters + out, pp.q.shadow, counters + 2, ctx->d_stats);
                else           k_shade<false><<<sg, sh_threads, 0, s>>>(sc, pp.st, b, bounce, in_queue, in_count, b.slots, pp.q.active[out], counters + out, pp.q.shadow, counters + 2, ctx->d_stats);
                debug_sync("k_shade", bounce, s);
                end_span(ctx, s);
                ctx->launches++;

                if (!merged || bounce + 1 == max_bounce) {
                    begin_span(ctx, ST_SHADOW, s);
                    if (stats) k_trace_shadow<true ><<<tg, 128, 0, s>>>(sc, pp.st, pp.q.shadow, counters + 2, counters + 4, ctx->refill, ctx->d_stats);
                    else       k_trace_shadow<false><<<tg, 128, 0, s>>>(sc, pp.st, pp.q.shadow, counters + 2, counters + 4, ctx->refill, ctx->d_stats);
                    debug_sync("k_trace_shadow", bounce, s);
                    end_span(ctx, s);
                    ctx->launches++; ctx->trace_launches++;
                }
            }

            CK(cudaStreamWaitEvent(s, ctx->film_free, 0));    // the film's readers / clears enqueued before this pass (and the previous pass itself) are through
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
    for (int p = 0; p < BPT_MAX_PIPES; ++p) {          // the join: what follows on ctx->stream (film readers, the next upload's last_use) sees the whole pass
        CK(cudaEventRecord(ctx->pipes[p].done, ctx->pipes[p].stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->pipes[p].done, 0));
    }
    CK(cudaEventRecord(ctx->pass_end, ctx->stream));
    mark_scene_use(ctx);
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
        if (!ctx->scene_ready && !ctx->pending) { set_error("bpt_render_pass: no scene uploaded"); return BPT_ERR_STATE; }
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
    const bool dump = getenv("BPT_DUMP_SPANS") != nullptr;
    static const char* stage_names[ST_COUNT] = {"raygen", "trace", "shade", "shadow", "splat"};
    for (size_t i = 0; i < ctx->spans_used; ++i) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->spans[i].a, ctx->spans[i].b) == cudaSuccess) acc[ctx->spans[i].stage] += ms;
        if (dump) fprintf(stderr, "span %3zu %-7s %9.4f ms\n", i, stage_names[ctx->spans[i].stage], ms);
    }
    out->raygen_ms = acc[ST_RAYGEN]; out->trace_ms = acc[ST_TRACE]; out->shade_ms = acc[ST_SHADE];
    out->shadow_ms = acc[ST_SHADOW]; out->splat_ms = acc[ST_SPLAT];
    out->kernel_launches = ctx->launches; out->trace_launches = ctx->trace_launches;
    return BPT_OK;
}

static int resolve_film(bpt_ctx* ctx, const float4* film, const bpt_post_settings* post, const uint8_t* dither_rgb8,
                        uint32_t dither_w, uint32_t dither_h, uint32_t* out_pixels, const char* who) {
    if (dither_rgb8 && (dither_w == 0 || dither_h == 0 || (dither_w & (dither_w - 1)) || (dither_h & (dither_h - 1)))) {
        set_error("%s: dither tile must have power-of-two width and height", who); return BPT_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    size_t n = (size_t)ctx->film_w*ctx->film_h;
    void* d = nullptr;
    int rc = device_slot(ctx->misc_slots, SL_RESOLVE_OUT, n*sizeof(uint32_t), &d);   if (rc) return rc;   uint32_t* d_out = (uint32_t*)d;
    uint8_t* d_dither = nullptr;
    if (dither_rgb8) {
        size_t db = (size_t)dither_w*dither_h*3;
        rc = device_slot(ctx->misc_slots, SL_RESOLVE_DITHER, db, &d);   if (rc) return rc;   d_dither = (uint8_t*)d;
        CK(cudaMemcpyAsync(d_dither, dither_rgb8, db, cudaMemcpyHostToDevice, ctx->stream));
        ctx->h2d_bytes += db;
    }
    k_resolve<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(film, ctx->film_w, ctx->film_h, *post, d_dither, dither_w, dither_h, d_out);
    ctx->total_launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_pixels, d_out, n*sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->d2h_bytes += n*sizeof(uint32_t);
    return device_error(ctx, who);
}

int bpt_resolve_bgra8(bpt_ctx* ctx, const bpt_post_settings* post, const uint8_t* dither_rgb8,
                      uint32_t dither_w, uint32_t dither_h, uint32_t* out_pixels) {
    if (!ctx || !post || !out_pixels) { set_error("bpt_resolve_bgra8: null argument"); return BPT_ERR_ARG; }
    if (!ctx->film) { set_error("bpt_resolve_bgra8: no film"); return BPT_ERR_STATE; }
    return resolve_film(ctx, ctx->film, post, dither_rgb8, dither_w, dither_h, out_pixels, "bpt_resolve_bgra8");
}

int bpt_resolve_reduced_bgra8(bpt_ctx* ctx, const bpt_post_settings* post, const uint8_t* dither_rgb8,
                              uint32_t dither_w, uint32_t dither_h, uint32_t* out_pixels) {
    if (!ctx || !post || !out_pixels) { set_error("bpt_resolve_reduced_bgra8: null argument"); return BPT_ERR_ARG; }
    if (!ctx->film_reduced || ctx->film_reduced_w != ctx->film_w || ctx->film_reduced_h != ctx->film_h) {
        set_error("bpt_resolve_reduced_bgra8: bpt_reduce_film has not run for this film"); return BPT_ERR_STATE;
    }
    return resolve_film(ctx, ctx->film_reduced, post, dither_rgb8, dither_w, dither_h, out_pixels, "bpt_resolve_reduced_bgra8");
}

int bpt_build_mesh_bvh_device(bpt_ctx* ctx, uint32_t n, const float* positions, int32_t method, bpt_bvh_node* nodes_out,
                              uint32_t node_capacity, uint32_t* node_count, uint32_t* indices_out, float* build_ms) {
    using namespace bpt::gbvh;
    static_assert(sizeof(OutNode) == sizeof(bpt_bvh_node), "node layout");
    if (!ctx || !positions || !nodes_out || !node_count || !indices_out || n == 0) { set_error("bpt_build_mesh_bvh_device: null argument"); return BPT_ERR_ARG; }
    if (node_capacity < 2) { set_error("bpt_build_mesh_bvh_device: node buffer too small"); return BPT_ERR_ARG; }
    if (method != BPT_BVH_SAH_BINNED && method != BPT_BVH_MIDPOINT_SPLIT) {
        set_error("bpt_build_mesh_bvh_device: method %d is not built on the device (BPT_BVH_SAH_FULL is O(n^2), host only)", method);
        return BPT_ERR_UNSUPPORTED;
    }
    const int midpoint = method == BPT_BVH_MIDPOINT_SPLIT ? 1 : 0;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t T = 256;
    auto blocks = [&](uint64_t work) { return (uint32_t)((work + T - 1)/T); };
    const uint32_t max_level_nodes = (uint32_t)std::max<uint64_t>(4, 2ull*n/5 + 4);
    const uint32_t tiles = (n + kScanTile - 1)/kScanTile;

    std::vector<void*> owned;
    auto dalloc = [&](void** p, size_t bytes) -> int { CK(cudaMalloc(p, std::max<size_t>(bytes, 256))); owned.push_back(*p); return BPT_OK; };
    float* d_pos = nullptr; GEntry* d_e[2] = {nullptr, nullptr}; uint32_t* d_seg[2] = {nullptr, nullptr};
    uint2 *d_flags = nullptr, *d_scan = nullptr, *d_tiles = nullptr;
    uint32_t *d_posA = nullptr, *d_posB = nullptr, *d_perm = nullptr, *d_next = nullptr, *d_idx = nullptr;
    GNode* d_nodes = nullptr; GAcc* d_acc = nullptr; OutNode* d_out = nullptr;
    int rc = 0;
    rc |= dalloc((void**)&d_pos, (size_t)n*9*sizeof(float));
    for (int k = 0; k < 2; ++k) { rc |= dalloc((void**)&d_e[k], (size_t)n*sizeof(GEntry)); rc |= dalloc((void**)&d_seg[k], (size_t)n*4); }
    rc |= dalloc((void**)&d_flags, (size_t)n*8);
    rc |= dalloc((void**)&d_scan, ((size_t)n + 1)*8);
    rc |= dalloc((void**)&d_tiles, (size_t)tiles*8);
    rc |= dalloc((void**)&d_posA, (size_t)n*4); rc |= dalloc((void**)&d_posB, (size_t)n*4); rc |= dalloc((void**)&d_perm, (size_t)n*4);
    rc |= dalloc((void**)&d_next, 256); rc |= dalloc((void**)&d_idx, (size_t)n*4);
    rc |= dalloc((void**)&d_nodes, ((size_t)2*n + 4)*sizeof(GNode));
    rc |= dalloc((void**)&d_acc, (size_t)max_level_nodes*sizeof(GAcc));
    rc |= dalloc((void**)&d_out, ((size_t)2*n + 2)*sizeof(OutNode));
    auto cleanup = [&]() { for (void* p : owned) cudaFree(p); };
    if (rc) { cleanup(); set_error("bpt_build_mesh_bvh_device: out of device memory"); return BPT_ERR_CUDA; }

    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    cudaError_t err = cudaMemcpyAsync(d_pos, positions, (size_t)n*9*sizeof(float), cudaMemcpyHostToDevice, s);
    ctx->h2d_bytes += (uint64_t)n*9*sizeof(float);
    cudaEventRecord(ev0, s);
    k_make_entries<<<blocks(n), T, 0, s>>>(d_pos, n, d_e[0], d_seg[0]);
    k_init_root<<<1, 1, 0, s>>>(d_nodes, n);
    std::vector<std::pair<uint32_t, uint32_t>> levels;      // (first breadth-first id, count)
    uint32_t lvl_start = 0, lvl_count = 1;
    int cur = 0;
    uint64_t launches = 2;
    while (lvl_count > 0 && err == cudaSuccess) {
        if (levels.size() >= 65536 || lvl_count > max_level_nodes) { err = cudaErrorUnknown; break; }    // deeper than any tree of <= 2^32 entries can be split sensibly
        levels.push_back({lvl_start, lvl_count});
        GNode* lv = d_nodes + lvl_start;
        uint32_t next_base = lvl_start + lvl_count;
        cudaMemsetAsync(d_next, 0, 4, s);
        k_init_acc<<<blocks(lvl_count), T, 0, s>>>(d_acc, lvl_count);
        k_bounds<<<blocks(n), T, 0, s>>>(d_e[cur], d_seg[cur], n, d_acc);
        k_decide<<<blocks(lvl_count), T, 0, s>>>(lv, d_acc, lvl_count, midpoint);
        if (!midpoint) {
            k_bin<<<blocks(n), T, 0, s>>>(d_e[cur], d_seg[cur], n, d_acc);
            k_sah<<<blocks(lvl_count), T, 0, s>>>(lv, d_acc, lvl_count);
        }
        k_flags<<<blocks(n), T, 0, s>>>(d_e[cur], d_seg[cur], n, d_acc, d_flags, d_perm);
        k_scan_tiles<<<tiles, kScanBlock, 0, s>>>(d_flags, n, d_scan, d_tiles);
        k_scan_tile_sums<<<1, 1024, 0, s>>>(d_tiles, tiles, d_scan, n);
        k_scan_add<<<blocks(n), T, 0, s>>>(d_scan, n, d_tiles);
        k_scatter<<<blocks(n), T, 0, s>>>(d_seg[cur], n, lv, d_acc, d_flags, d_scan, d_posA, d_posB);
        k_pair<<<blocks(n), T, 0, s>>>(d_seg[cur], n, lv, d_acc, d_scan, d_posA, d_posB, d_perm);
        k_split<<<blocks(lvl_count), T, 0, s>>>(lv, d_acc, lvl_count, d_scan, d_posA, d_posB, d_nodes + next_base, next_base, d_next);
        k_apply<<<blocks(n), T, 0, s>>>(d_e[cur], d_e[cur ^ 1], d_perm, d_seg[cur], d_seg[cur ^ 1], n, lv, d_acc);
        launches += 13;
        uint32_t next_count = 0;
        err = cudaMemcpyAsync(&next_count, d_next, 4, cudaMemcpyDeviceToHost, s);
        if (err == cudaSuccess) err = cudaStreamSynchronize(s);
        cur ^= 1;
        lvl_start = next_base; lvl_count = next_count;
    }
    uint32_t total = lvl_start;
    uint32_t inner_root = 0;
    if (err == cudaSuccess) {
        for (size_t l = levels.size(); l-- > 0;) k_inner_count<<<blocks(levels[l].second), T, 0, s>>>(d_nodes, levels[l].first, levels[l].second);
        for (size_t l = 0; l < levels.size(); ++l) k_rank<<<blocks(levels[l].second), T, 0, s>>>(d_nodes, levels[l].first, levels[l].second);
        launches += 2*levels.size();
        GNode root;
        err = cudaMemcpyAsync(&root, d_nodes, sizeof(GNode), cudaMemcpyDeviceToHost, s);
        if (err == cudaSuccess) err = cudaStreamSynchronize(s);
        inner_root = root.inner_count;
    }
    uint32_t out_count = 2u + 2u*inner_root;
    if (err == cudaSuccess && out_count > node_capacity) { cleanup(); cudaEventDestroy(ev0); cudaEventDestroy(ev1); set_error("bpt_build_mesh_bvh_device: node buffer too small (%u nodes)", out_count); return BPT_ERR_ARG; }
    if (err == cudaSuccess) {
        cudaMemsetAsync(d_out, 0, (size_t)out_count*sizeof(OutNode), s);
        k_emit<<<blocks(total), T, 0, s>>>(d_nodes, total, d_out);
        k_indices<<<blocks(n), T, 0, s>>>(d_e[cur], n, d_idx);
        launches += 2;
        cudaEventRecord(ev1, s);
        err = cudaMemcpyAsync(nodes_out, d_out, (size_t)out_count*sizeof(OutNode), cudaMemcpyDeviceToHost, s);
        if (err == cudaSuccess) err = cudaMemcpyAsync(indices_out, d_idx, (size_t)n*4, cudaMemcpyDeviceToHost, s);
        if (err == cudaSuccess) err = cudaStreamSynchronize(s);
        if (err == cudaSuccess) err = cudaGetLastError();
        ctx->d2h_bytes += (uint64_t)out_count*sizeof(OutNode) + (uint64_t)n*4;
    }
    float ms = 0.0f;
    if (err == cudaSuccess) cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    cleanup();
    ctx->total_launches += launches;
    if (err != cudaSuccess) { set_error("bpt_build_mesh_bvh_device: %s", cudaGetErrorString(err)); cudaGetLastError(); return BPT_ERR_CUDA; }
    *node_count = out_count;
    if (build_ms) *build_ms = ms;
    return BPT_OK;
}

// ---- multi-GPU: NCCL is resolved at run time so that single-GPU users of the library do not need it ------------------------
namespace {
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, bpt_nccl_id, int) = nullptr;     // ncclUniqueId is passed by value: same 128-byte POD
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
    if (!api.handle) { set_error("multi-GPU: cannot load libnccl.so.2 (%s)", dlerror()); return nullptr; }
    #define SYM(field, name) do { *(void**)&api.field = dlsym(api.handle, name); if (!api.field) { set_error("multi-GPU: %s missing from libnccl", name); return nullptr; } } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommInitAll, "ncclCommInitAll");
    SYM(CommDestroy, "ncclCommDestroy"); SYM(Reduce, "ncclReduce"); SYM(GetErrorString, "ncclGetErrorString");
    #undef SYM
    api.ok = true;
    return &api;
}
enum { kNcclFloat32 = 7, kNcclSum = 0 };        // ncclFloat32 / ncclSum (nccl.h: ncclDataType_t, ncclRedOp_t)
#define NK(api, call, what) do { int r_ = (call); if (r_ != 0) { set_error("%s: NCCL error %d (%s)", what, r_, (api)->GetErrorString(r_)); return BPT_ERR_CUDA; } } while (0)
} // namespace

int bpt_nccl_get_unique_id(bpt_nccl_id* out) {
    if (!out) { set_error("bpt_nccl_get_unique_id: null argument"); return BPT_ERR_ARG; }
    NcclApi* api = nccl_api(); if (!api) return BPT_ERR_UNSUPPORTED;
    NK(api, api->GetUniqueId(out), "ncclGetUniqueId");
    return BPT_OK;
}

int bpt_nccl_comm_init_rank(bpt_ctx* ctx, const bpt_nccl_id* id, int nranks, int rank, void** out_comm) {
    if (!ctx || !id || !out_comm || nranks < 1 || rank < 0 || rank >= nranks) { set_error("bpt_nccl_comm_init_rank: bad arguments"); return BPT_ERR_ARG; }
    NcclApi* api = nccl_api(); if (!api) return BPT_ERR_UNSUPPORTED;
    CK(cudaSetDevice(ctx->device));
    NK(api, api->CommInitRank(out_comm, nranks, *id, rank), "ncclCommInitRank");
    return BPT_OK;
}

int bpt_nccl_comm_init_all(int ndev, const int* devices, void** out_comms) {
    if (ndev < 1 || !out_comms) { set_error("bpt_nccl_comm_init_all: bad arguments"); return BPT_ERR_ARG; }
    NcclApi* api = nccl_api(); if (!api) return BPT_ERR_UNSUPPORTED;
    NK(api, api->CommInitAll(out_comms, ndev, devices), "ncclCommInitAll");
    return BPT_OK;
}

int bpt_nccl_comm_destroy(void* comm) {
    if (!comm) return BPT_OK;
    NcclApi* api = nccl_api(); if (!api) return BPT_ERR_UNSUPPORTED;
    NK(api, api->CommDestroy(comm), "ncclCommDestroy");
    return BPT_OK;
}

int bpt_reduce_film(bpt_ctx* ctx, void* nccl_comm, int root) {
    if (!ctx || !nccl_comm || root < 0) { set_error("bpt_reduce_film: bad arguments"); return BPT_ERR_ARG; }
    if (!ctx->film) { set_error("bpt_reduce_film: no film"); return BPT_ERR_STATE; }
    NcclApi* api = nccl_api(); if (!api) return BPT_ERR_UNSUPPORTED;
    CK(cudaSetDevice(ctx->device));
    size_t count = (size_t)ctx->film_w*ctx->film_h*4;
    // every rank owns a destination buffer (NCCL only writes the root's): the caller does not have to know its rank here
    if (!ctx->film_reduced || ctx->film_reduced_w != ctx->film_w || ctx->film_reduced_h != ctx->film_h) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->film_reduced); ctx->film_reduced = nullptr;
        CK(cudaMalloc((void**)&ctx->film_reduced, count*sizeof(float)));
        CK(cudaMemsetAsync(ctx->film_reduced, 0, count*sizeof(float), ctx->stream));
        ctx->film_reduced_w = ctx->film_w; ctx->film_reduced_h = ctx->film_h;
    }
    // on the context's stream: behind every pass enqueued so far (render_rows joins its pipeline streams into it), and
    // the next pass starts behind it -- no host synchronisation anywhere
    NK(api, api->Reduce(ctx->film, ctx->film_reduced, count, kNcclFloat32, kNcclSum, root, nccl_comm, ctx->stream), "ncclReduce");
    return BPT_OK;
}

int bpt_reduced_film_device_ptr(bpt_ctx* ctx, void** out) {
    if (!ctx || !out) { set_error("bpt_reduced_film_device_ptr: null argument"); return BPT_ERR_ARG; }
    *out = ctx->film_reduced;
    return ctx->film_reduced ? BPT_OK : BPT_ERR_STATE;
}

int bpt_download_reduced_film(bpt_ctx* ctx, float* out) {
    if (!ctx || !out) { set_error("bpt_download_reduced_film: null argument"); return BPT_ERR_ARG; }
    if (!ctx->film_reduced) { set_error("bpt_download_reduced_film: bpt_reduce_film has not run"); return BPT_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    size_t bytes = (size_t)ctx->film_reduced_w*ctx->film_reduced_h*sizeof(float4);
    CK(cudaMemcpyAsync(out, ctx->film_reduced, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->d2h_bytes += bytes;
    return device_error(ctx, "bpt_download_reduced_film");
}

int bpt_set_tail_threshold(bpt_ctx* ctx, uint32_t paths) {
    if (!ctx) { set_error("bpt_set_tail_threshold: null ctx"); return BPT_ERR_ARG; }
    if (paths > (1u << 22)) { set_error("bpt_set_tail_threshold: at most %u paths", 1u << 22); return BPT_ERR_ARG; }
    ctx->tail_threshold = paths;
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
