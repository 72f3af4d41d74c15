// Wavefront kernels of the path-tracing core: ray generation, closest-hit / occlusion traversal, shading
// (the body of advanced_integrator, Raytracer/integrators.cpp:581-821), and film splatting
// (render_tile's tail + splat_filter, Raytracer/raytracer.cpp:187-259, :469-488).
#pragma once
#include "trace.cuh"

namespace bpt {

// Path-state accesses.  Every array is streamed -- written by one kernel, read once by the next -- while the BVH nodes and
// triangles the traversal kernels walk are re-read all the time: -DBPT_STREAM_HINTS=1 marks the path-state traffic
// evict-first (ld.global.cs / st.global.cs) so that it does not push the scene out of L2.
#ifndef BPT_STREAM_HINTS
#define BPT_STREAM_HINTS 0
#endif
#if BPT_STREAM_HINTS
#define PSL(p) __ldcs(p)
#define PSS(p, v) __stcs((p), (v))
#else
#define PSL(p) (*(p))
#define PSS(p, v) (*(p) = (v))
#endif

// One batch = pixel rows [ya, yb) of the pass rect x samples [sa, sb).  slot = (row-major pixel in batch)*S + (s - sa)
struct BatchDesc {
    int32_t  x0;              // first pixel column of the rect
    uint32_t row0;            // first entry of row_map this batch covers
    const int32_t* row_map;   // device list of the pass's image rows (a rect, or a rank's interleaved row blocks)
    uint32_t rect_w, rows;    // rect width, rows in this batch
    uint32_t sa, S;           // first sample of this batch, samples per pixel in this batch
    uint32_t frame_count;     // AccumulationBuffer::frame_count of the pass
    uint32_t salt;
    uint32_t slots;           // rect_w*rows*S
    uint32_t want_records;
    uint32_t magic_S, magic_w;   // udiv_magic multipliers of S and rect_w (set_batch_magic)
};

inline void set_batch_magic(BatchDesc& b) {
    auto magic = [](uint32_t d) { uint64_t m = (1ull << 32)/d; return (uint32_t)(m > 0xFFFFFFFFull ? 0xFFFFFFFFull : m); };
    b.magic_S = magic(b.S); b.magic_w = magic(b.rect_w);
}

BPT_D SamplerCtx make_sampler(const DScene& sc, const BatchDesc& b, uint32_t slot) {
    uint32_t s, px;
    uint32_t pix = udiv_magic(slot, b.S, b.magic_S, s);
    uint32_t row = udiv_magic(pix, b.rect_w, b.magic_w, px);
    SamplerCtx c;
    c.strata_perm = sc.strata_perm; c.bn_sobol = sc.bn_sobol; c.bn_scramble = sc.bn_scramble; c.bn_rank = sc.bn_rank;
    c.strategy = sc.settings.sampling_strategy;
    c.index = b.frame_count + b.sa + s;
    c.x = (uint32_t)(b.x0 + (int32_t)px);
    c.y = (uint32_t)__ldg(&b.row_map[b.row0 + row]);
    return c;
}

// ---- camera (raytracer.cpp:86-123, :372-461) -------------------------------------------------------------------------

BPT_D V2 brown_conrady(V2 uv, float amount, float woh) {
    uv.y = uv.y / woh;
    float b1 = 0.1f*amount, b2 = -0.025f*amount;
    float r2 = uv.x*uv.x + uv.y*uv.y;
    float s = (1.0f + r2*b1) + r2*r2*b2;
    uv.x = uv.x*s; uv.y = uv.y*s;
    uv.y = uv.y*woh;
    return uv;
}

BPT_D void apply_lens_distortion(float amount, uint32_t w, uint32_t h, float& u, float& v) {
    float woh = (float)w / (float)h;
    V2 zero; zero.x = 0.0f; zero.y = 0.0f;
    V2 one;  one.x = 1.0f;  one.y = 1.0f;
    V2 mn = brown_conrady(zero, amount, woh);
    V2 mx = brown_conrady(one, amount, woh);
    V2 uv; uv.x = u; uv.y = v;
    uv = brown_conrady(uv, amount, woh);
    if (amount > 0.0f) {
        uv.x = (uv.x - mn.x) / (mn.x + mx.x);
        uv.y = (uv.y - mn.y) / (mn.y + mx.y);
    }
    u = uv.x; v = uv.y;
}

BPT_D V2 transform_bokeh_sample(V2 o, float f, float n, float phi_shutter_max) {
    float abx = (o.x*2.0f) - 1.0f, aby = (o.y*2.0f) - 1.0f;
    float px, py;
    if ((abx*abx) > (aby*aby)) {
        px = (fabsf(abx) > 1e-8f) ? ((kPi*0.25f)*(aby / abx)) : 0.0f;
        py = abx;
    } else {
        px = (fabsf(aby) > 1e-8f) ? ((kPi*0.5f) - ((kPi*0.25f)*(abx / aby))) : 0.0f;
        py = aby;
    }
    px += f*phi_shutter_max;
    float scale = 1.0f;
    if (f > 0.0f) {
        float seg = (2.0f*(kPi / n))*floorf(((n*px) + kPi) / (2.0f*kPi));
        scale = pow_f(cos_f(kPi / n) / cos_f(px - seg), f);
    }
    py *= scale;
    V2 r; r.x = cos_f(px)*py; r.y = sin_f(px)*py;
    return r;
}

// ---- shading helpers (integrators.cpp:11-19, :58-119, :235-308) ------------------------------------------------------

BPT_D void get_tangents(V3 n, V3& b1, V3& b2) {
    float sign = copy_sign(1.0f, n.z);
    float a = -1.0f / (sign + n.z);
    float b = n.x*n.y*a;
    b1 = v3(1.0f + sign*n.x*n.x*a, sign*b, -sign*n.x);
    b2 = v3(b, sign + n.y*n.y*a, -n.y);
}

BPT_D V3 oriented_around_normal(V3 v, V3 N) {
    V3 T, B;
    get_tangents(N, T, B);
    return (v.x*B + v.y*N + v.z*T);
}

BPT_D V3 map_to_hemisphere(V3 N, V2 u) {
    float azimuth = kTau*u.x;
    float y = u.y;
    float s = sqrtf(1.0f - y*y);
    float sn, cs;
    sincos_f(azimuth, sn, cs);
    V3 hemi = v3(cs*s, y, sn*s);
    return oriented_around_normal(hemi, N);
}

BPT_D V3 map_to_cosine_weighted_hemisphere(V3 N, V2 u) {
    float azimuth = kTau*u.x;
    float y = u.y;
    float s = sqrtf(1.0f - y);
    float sn, cs;
    sincos_f(azimuth, sn, cs);
    V3 hemi = v3(cs*s, sqrtf(y), sn*s);
    return oriented_around_normal(hemi, N);
}

BPT_D float fresnel_dielectric(float cos_i, float eta_i, float eta_t, float ratio, float& cos_t_out) {
    float sin_i = sqrtf(max_t(0.0f, 1.0f - cos_i*cos_i));
    float sin_t = ratio*sin_i;
    float cos_t = sqrtf(max_t(0.0f, 1.0f - sin_t*sin_t));
    cos_t_out = cos_t;
    if (sin_t >= 1.0f) return 1.0f;
    float r_par  = (((eta_t*cos_i) - (eta_i*cos_t)) / ((eta_t*cos_i) + (eta_i*cos_t)));
    float r_perp = (((eta_i*cos_i) - (eta_t*cos_t)) / ((eta_i*cos_i) + (eta_t*cos_t)));
    return 0.5f*(r_par*r_par + r_perp*r_perp);
}

BPT_D V3 sample_sky(const DScene& sc, V3 d) {
    if (sc.skydome) {
        float rcp_pi = 1.0f / kPi, rcp_2pi = 0.5f / kPi;
        float phi = atan2_f(d.z, d.x);
        float theta = asin_f(d.y);
        float u = 0.5f + rcp_2pi*phi;
        float v = 0.5f + rcp_pi*theta;
        int w = (int)sc.skydome_w, h = (int)sc.skydome_h;
        int sx = (int)(u*(float)w) % w;
        int sy = (int)(v*(float)h) % h;
        if (sx < 0) sx = 0;          // the reference reads out of bounds here (Appendix A #9); clamp instead
        if (sy < 0) sy = 0;
        float4 px = __ldg(&sc.skydome[sy*w + sx]);
        return v3(px.x, px.y, px.z);
    }
    float sky_t = fabsf(d.y);
    return lerp_v(v3(sc.bot_sky), v3(sc.top_sky), sky_t);
}

struct MatView {
    uint32_t flags;
    V3 albedo, checker, emission, absorb;
    float ior, metallic, roughness;
    int medium;
};

BPT_D MatView load_material(const DScene& sc, uint32_t id) {
    const DMaterial* m = sc.materials + id;
    MatView v;
    v.flags = __ldg(&m->flags);
    v.albedo = v3(__ldg(&m->albedo[0]), __ldg(&m->albedo[1]), __ldg(&m->albedo[2]));
    v.checker = v3(__ldg(&m->checker_color[0]), __ldg(&m->checker_color[1]), __ldg(&m->checker_color[2]));
    v.emission = v3(__ldg(&m->emission_color[0]), __ldg(&m->emission_color[1]), __ldg(&m->emission_color[2]));
    v.absorb = v3(__ldg(&m->absorb[0]), __ldg(&m->absorb[1]), __ldg(&m->absorb[2]));
    v.ior = __ldg(&m->ior); v.metallic = __ldg(&m->metallic); v.roughness = __ldg(&m->roughness);
    v.medium = __ldg(&m->is_participating_medium);
    return v;
}

// :NormalCalculation (intersection.cpp:526-591): hit point + world normal from a hit record
BPT_D void hit_geometry(const DScene& sc, V3 ro, V3 rd, const HitRecord& h, V3& I, V3& N, uint32_t& material) {
    float t = h.t;
    I = ro + t*rd;
    V3 n = v3(0.0f);
    if (h.prim & BPT_HIT_PLANE) {
        const DPlane& pl = sc.planes[h.prim & ~BPT_HIT_PLANE];
        n = v3(__ldg(&pl.n[0]), __ldg(&pl.n[1]), __ldg(&pl.n[2]));
        material = __ldg(&pl.material);
        // planes carry the shared identity transform (scene.cpp:76, :86-90)
        float4 id[3] = {make_float4(1, 0, 0, 0), make_float4(0, 1, 0, 0), make_float4(0, 0, 1, 0)};
        N = noz(xform_normal(id, n));
        return;
    }
    const DPrimitive* prim = sc.primitives + h.prim;
    float4 m[3] = {__ldg(&prim->inv[0]), __ldg(&prim->inv[1]), __ldg(&prim->inv[2])};
    material = __ldg(&prim->material);
    uint32_t type = __ldg(&prim->type);
    if (type == BPT_PRIM_MESH) {
        const DMesh* mesh = sc.meshes + __ldg(&prim->mesh);
        if (__ldg(&mesh->has_normals)) {
            const float4* nt = sc.normals + (size_t)h.tri*3;
            V3 na = v3(__ldg(&nt[0])), nb = v3(__ldg(&nt[1])), nc = v3(__ldg(&nt[2]));
            float ux = 1.0f - h.v - h.w;
            n = (ux*na + h.v*nb + h.w*nc);
        } else {
            const DTriangle* tri = sc.triangles + h.tri;
            V3 e1 = normalize(v3(__ldg(&tri->e1)));
            V3 e2 = normalize(v3(__ldg(&tri->e2)));
            n = cross(e1, e2);
        }
    } else {
        V3 oo = xform(m, ro, 1.0f), od = xform(m, rd, 0.0f);
        V3 op = oo + t*od;
        if (type == BPT_PRIM_SPHERE) {
            n = op;
        } else if (type == BPT_PRIM_BOX) {
            V3 rel = op / v3(__ldg(&prim->box_r[0]), __ldg(&prim->box_r[1]), __ldg(&prim->box_r[2]));
            int axis = 0;
            float largest = fabsf(rel.x);
            if (fabsf(rel.y) > largest) { axis = 1; largest = fabsf(rel.y); }
            if (fabsf(rel.z) > largest) { axis = 2; largest = fabsf(rel.z); }
            n = v3(axis == 0 ? sign_of(rel.x) : 0.0f, axis == 1 ? sign_of(rel.y) : 0.0f, axis == 2 ? sign_of(rel.z) : 0.0f);
        }
    }
    N = noz(xform_normal(m, n));
}

// the per-sample RandomSeries as k_raygen leaves it: seeded per pixel and sample, advanced by the AA and the
// DOF draw (every get_next_sample_* call advances the series exactly once, samplers.h:36-45)
BPT_D uint4 primary_rng_state(const SamplerCtx& sm, const BatchDesc& b) {
    uint4 rng = random_seed(hash_coordinate3(sm.x, sm.y, sm.index) ^ b.salt);
    next_set(rng); next_set(rng);
    return rng;
}

// raytracer.cpp:469-474
BPT_D float primary_vignette(const DScene& sc, V3 ray_d) {
    float vignette = dot(ray_d, v3(sc.camera.z));
    vignette = vignette*vignette*vignette*vignette;
    return lerp_f(1.0f, vignette, sc.settings.vignette_strength);
}

// material_stack[level] of a path: level 0 is always the integrator's local "air" (integrators.cpp:597-600) and is not
// stored; level L >= 1 lives in plane L-1 of DPathState::mstack
BPT_D uint32_t mstack_get(const DScene& sc, const DPathState& st, const BatchDesc& b, uint32_t slot, int level) {
    return level <= 0 ? sc.air_material : (uint32_t)st.mstack[(size_t)(level - 1)*b.slots + slot];
}

// An NEE shadow ray that never reaches a mesh BLAS is settled where it is generated.  This is the part of
// intersect_shadow_ray that precedes the first intersect_mesh call -- the planes, the TLAS root, and, when the whole TLAS
// is one leaf, its spheres / boxes and the root box of each mesh (intersection.cpp:424-454, :456-520) -- executed with the
// functions the traversal kernels use, in their order, so it decides exactly what persistent_trace would:
//   2  occluded for certain: a plane is hit (occlusion mode does not return on it, but the ray ends with a hit record whatever
//      follows), or a sphere / box of the leaf is hit;
//   1  unoccluded for certain: the TLAS root is missed, or every item of the leaf is missed (mesh: its root box);
//   0  undecided: a mesh BLAS (or a TLAS with inner nodes, a leaf of more than 7 items, a ray with a zero / denormal / huge
//      direction component) -> the ray goes to the shadow queue and the traversal kernel starts over with it.
// In k_shade all 32 lanes run it together; in the traversal kernel the same work is a refill + a TLAS-item step at ~18 of 32
// lanes, plus 96 bytes of queue traffic per ray.  `bytes`: the algorithmic bytes of the visits made (SURVEY 8d units).
// (The same test for PRIMARY rays inside k_raygen was measured and dropped: k_raygen 1.7 -> 3.6 ms, bounce-0 traversal
// 8.0 -> 6.0 ms on C2 -- an even trade, the settled rays were not costing the traversal kernel more than they cost here.)
BPT_D int shadow_tlas_head(const DScene& sc, V3 o, V3 d, float max_t, uint32_t ignored, uint32_t& bytes) {
    RayT ray;
    make_ray(ray, o, d);
    bytes = 0u;
    float t = max_t;
    if (!sc.tame_bounds || !(ray.neg & BPT_RAY_TAME)) return 0;
    for (uint32_t i = 0; i < sc.plane_count; ++i) {
        const DPlane& pl = sc.planes[i];
        if (plane_test(ray, v3(__ldg(&pl.n[0]), __ldg(&pl.n[1]), __ldg(&pl.n[2])), __ldg(&pl.d), t)) return 2;
    }
    float tn;
    bytes = 32u;                                                                                   // the TLAS root is popped
    const bool root = slab_test(ray, sc.tlas_root_q0.x, sc.tlas_root_q0.y, sc.tlas_root_q0.z, sc.tlas_root_q0.w, sc.tlas_root_q1.x, sc.tlas_root_q1.y, tn);
    if (!(root && tn < t)) return 1;
    const uint32_t ref = __float_as_uint(sc.tlas_root_q1.z);
    if (!(ref & BPT_WREF_LEAF)) return 0;
    const uint32_t count = (ref >> 28) & 7u, first = ref & BPT_WREF_INDEX_MASK;
    if (count == 0u) return 0;
    for (uint32_t k = 0; k < count; ++k) {
        const uint32_t prim_index = __ldg(&sc.tlas_indices[first + k]);
        if (prim_index == ignored) continue;
        const DPrimitive* prim = sc.primitives + prim_index;
        float4 m[3] = {__ldg(&prim->inv[0]), __ldg(&prim->inv[1]), __ldg(&prim->inv[2])};
        RayT oray;
        oray.o = xform(m, ray.o, 1.0f); oray.d = xform(m, ray.d, 0.0f);
        oray.inv = v3(0.0f); oray.neg = 0u;
        bytes += 104u;
        const uint32_t type = __ldg(&prim->type);
        if (type == BPT_PRIM_SPHERE) {
            if (sphere_test(oray, __ldg(&prim->sphere_r), t)) return 2;
        } else {
            make_ray(oray, oray.o, oray.d);
            if (!(oray.neg & BPT_RAY_TAME)) return 0;
            if (type == BPT_PRIM_BOX) {
                if (box_test(oray, __ldg(&prim->box_r[0]), __ldg(&prim->box_r[1]), __ldg(&prim->box_r[2]), t)) return 2;
            } else if (type == BPT_PRIM_MESH) {
                const DMesh* mesh = sc.meshes + __ldg(&prim->mesh);
                float4 q0 = __ldg(&mesh->root_q0), q1 = __ldg(&mesh->root_q1);
                bytes += 32u;                                                                      // the BLAS root is popped
                if (slab_test(oray, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tn) && tn < t) return 0;
            }
        }
    }
    return 1;
}

// ---- kernels ---------------------------------------------------------------------------------------------------------------

// render_tile's per-sample ray setup (raytracer.cpp:372-461) with per-pixel counter-based seeding
__global__ void __launch_bounds__(256)
k_raygen(DScene sc, DPathState st, BatchDesc b) {
    for (uint32_t slot = blockIdx.x*blockDim.x + threadIdx.x; slot < b.slots; slot += gridDim.x*blockDim.x) {
        SamplerCtx sm = make_sampler(sc, b, slot);
        uint4 rng = random_seed(hash_coordinate3(sm.x, sm.y, sm.index) ^ b.salt);

        const bpt_camera& cam = sc.camera;
        V3 cam_p = v3(cam.p), cam_x = v3(cam.x), cam_y = v3(cam.y), cam_z = v3(cam.z);
        float focus = cam.focus_distance, lens_radius = cam.lens_radius;
        float half_w = cam.half_film_w*focus, half_h = cam.half_film_h*focus;
        float film_distance = focus*cam.film_distance;
        V3 film_center = (cam_p - film_distance*cam_z);
        float pixel_w = 1.0f / (float)sc.film_w, pixel_h = 1.0f / (float)sc.film_h;

        float v = 1.0f - 2.0f*(float)(int32_t)sm.y*pixel_h;
        float u = 1.0f - 2.0f*(float)(int32_t)sm.x*pixel_w;
        apply_lens_distortion(sc.settings.lens_distortion, sc.film_w, sc.film_h, u, v);

        V2 aa = sample_2d(sm, rng, Sample_AA, 0);
        float jx = aa.x - 0.5f, jy = aa.y - 0.5f;
        // With lens_radius == 0 the lens sample is multiplied by zero: dof_x / dof_y are +-0 and
        // cam_p + (+-0)*cam_x + (+-0)*cam_y is cam_p bit for bit -- unless a component of cam_p is -0.0 (the sum's zero would
        // take its sign from the sample) or the camera basis is not finite.  In that (usual) case the DOF draw and the
        // polygonal-bokeh transform (cos / sin / pow in double) are skipped; the RNG state the first shade re-derives
        // (primary_rng_state) does not depend on them being executed here.
        const bool lens_inert = lens_radius == 0.0f &&
                                __float_as_uint(cam_p.x) != 0x80000000u && __float_as_uint(cam_p.y) != 0x80000000u && __float_as_uint(cam_p.z) != 0x80000000u &&
                                fabsf(cam_x.x) + fabsf(cam_x.y) + fabsf(cam_x.z) + fabsf(cam_y.x) + fabsf(cam_y.y) + fabsf(cam_y.z) + fabsf(half_w*pixel_w) + fabsf(half_h*pixel_h) < 3.0e38f;
        float dof_x = 0.0f, dof_y = 0.0f;
        if (!lens_inert) {
            V2 dof = sample_2d(sm, rng, Sample_DOF, 0);
            dof = transform_bokeh_sample(dof, sc.settings.f_factor, sc.settings.diaphragm_edges, kPi*sc.settings.phi_shutter_max);
            dof_x = half_w*pixel_w*lens_radius*dof.x;
            dof_y = half_h*pixel_h*lens_radius*dof.y;
        }

        V3 film_p = film_center;
        film_p = film_p + (u + pixel_w*jx)*half_w*cam_x;
        film_p = film_p + (v + pixel_h*jy)*half_h*cam_y;
        V3 lens_p = (cam_p + dof_x*cam_x + dof_y*cam_y);
        V3 ray_o = lens_p;
        V3 ray_d = normalize(film_p - lens_p);

        PSS(st.ray_o + slot, make_float4(ray_o.x, ray_o.y, ray_o.z, 3.402823466e+38f));    // make_ray's FLT_MAX far clip
        PSS(st.ray_d + slot, make_float4(ray_d.x, ray_d.y, ray_d.z, 0.0f));
        PSS(st.jitter + slot, make_float2(jx, jy));
        // Everything else IntegratorState starts with (integrators.cpp:587-600) is a function of the slot, so the first
        // bounce's shading re-derives it instead of reading it back from HBM: throughput = 1, radiance = 0, no previous
        // normal ("specular"), material stack = {air}, the RNG state = this seed advanced by the two draws above
        // (primary_rng_state), the vignette = primary_vignette(ray_d).  40 of 139 bytes per sample are written here.
        if (b.want_records) {
            st.primary_d[slot] = make_float4(ray_d.x, ray_d.y, ray_d.z, 0.0f);
            st.primary_o[slot] = make_float4(ray_o.x, ray_o.y, ray_o.z, 0.0f);
        }
    }
}

BPT_D void flush_counters(DStats* stats, const TraceCounters& c, bool shadow) {
    // warp-aggregate, then one atomic per counter per warp (rays themselves are counted in k_shade / k_trace_api)
    unsigned long long vals[7] = {c.tlas_pops, c.instances, c.mesh_calls, c.blas_pops, c.blas_inner, c.blas_leaves, c.tris};
    #pragma unroll
    for (int k = 0; k < 7; ++k) {
        unsigned long long v = vals[k];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0 && v) {
            atomicAdd(&stats->v[2 + k], v);
            if (shadow) atomicAdd(&stats->v[10 + k], v);
        }
    }
}

BPT_D void flush_ray_counts(DStats* stats, uint32_t rays, uint32_t shadow) {
    rays = __reduce_add_sync(0xFFFFFFFFu, rays);
    shadow = __reduce_add_sync(0xFFFFFFFFu, shadow);
    if ((threadIdx.x & 31) == 0) {
        if (rays) atomicAdd(&stats->v[0], (unsigned long long)rays);
        if (shadow) atomicAdd(&stats->v[1], (unsigned long long)shadow);
    }
}

#ifndef BPT_TRACE_MIN_CTAS
#define BPT_TRACE_MIN_CTAS 9      // 56 registers, no spills since the per-ray state moved to shared memory: traversal 43.7 ms on C2 against 45.0 at 8 CTAs; 10 CTAs (51 registers) spill and lose (47 ms)
#endif

struct ClosestSrc {        // rays come from the path state (through the active queue), hits go back to it
    DPathState st;
    const uint32_t* queue;
    bool store_w;
    BPT_D void load(uint32_t i, V3& o, V3& d, float& max_t, uint32_t& ignored, bool& occ) const {
        uint32_t slot = queue ? queue[i] : i;
        float4 ro = PSL(st.ray_o + slot), rd = PSL(st.ray_d + slot);
        o = v3(ro); d = v3(rd); max_t = ro.w; ignored = 0u;       // intersect_scene passes PrimitiveID 0 (intersection.cpp:608)
        occ = false;
    }
    BPT_D void store(uint32_t i, const HitRecord& h) const {
        uint32_t slot = queue ? queue[i] : i;
        PSS(st.hit + slot, make_float4(h.t, __uint_as_float(h.prim), __uint_as_float(h.tri), h.v));
        if (store_w) st.hit_w[slot] = h.w;      // barycentric w is only read for meshes with vertex normals
    }
};

struct ShadowSrc {         // NEE shadow rays; an unoccluded ray releases its pending contribution (integrators.cpp:756-769)
    DPathState st;
    const DShadowItem* items;
    BPT_D void load(uint32_t i, V3& o, V3& d, float& max_t, uint32_t& ignored, bool& occ) const {
        float4 ro = PSL(&items[i].o_maxt), rd = PSL(&items[i].d_light);
        o = v3(ro); d = v3(rd); max_t = ro.w; ignored = __float_as_uint(rd.w);
        occ = true;
    }
    BPT_D void store(uint32_t i, const HitRecord& h) const {
        if (h.prim == BPT_HIT_MISS) {
            float4 c = PSL(&items[i].contrib_slot);
            uint32_t slot = __float_as_uint(c.w);
            float4 r = PSL(st.radiance + slot);
            r.x = r.x + c.x; r.y = r.y + c.y; r.z = r.z + c.z;
            PSS(st.radiance + slot, r);
        }
    }
};

// One launch for two ray populations: the extension rays of bounce b (closest hit) and the NEE shadow rays that
// bounce b-1's shading queued.  Both are produced by the same k_shade launch and are independent of each other, so
// tracing them together halves the number of traversal launches -- and with it the number of kernel tails, which are
// set by the latency chain of the longest single ray and do not shrink with the batch (they cap multi-GPU scaling).
// Closest-hit rays come first in the index space (they are the longer ones); shadow rays fill in behind them.
struct MergedSrc {
    ClosestSrc closest;
    ShadowSrc shadow;
    uint32_t n_closest;
    BPT_D void load(uint32_t i, V3& o, V3& d, float& max_t, uint32_t& ignored, bool& occ) const {
        if (i < n_closest) closest.load(i, o, d, max_t, ignored, occ);
        else shadow.load(i - n_closest, o, d, max_t, ignored, occ);
    }
    BPT_D void store(uint32_t i, const HitRecord& h) const {
        if (i < n_closest) closest.store(i, h);
        else shadow.store(i - n_closest, h);
    }
};

__global__ void __launch_bounds__(128, BPT_TRACE_MIN_CTAS)
k_trace_merged(DScene sc, DPathState st, const uint32_t* __restrict__ in_queue, const uint32_t* __restrict__ n_closest_ptr,
               const DShadowItem* __restrict__ items, const uint32_t* __restrict__ n_shadow_ptr,
               uint32_t* cursor, uint32_t refill) {
    uint32_t nc = *n_closest_ptr, ns = *n_shadow_ptr;
    TraceCounters ctr = {};
    MergedSrc src = {{st, in_queue, sc.normals != nullptr}, {st, items}, nc};
    persistent_trace<TRACE_MODE_MIXED, false, false>(sc, src, nc + ns, cursor, refill, ctr);
}

// intersect_scene for every active path (integrators.cpp:615).  in_queue == nullptr means "slot = i".
template <bool STATS>
__global__ void __launch_bounds__(128, BPT_TRACE_MIN_CTAS)
k_trace_closest(DScene sc, DPathState st, const uint32_t* __restrict__ in_queue, const uint32_t* __restrict__ n_ptr,
                uint32_t n_fixed, uint32_t* cursor, uint32_t refill, DStats* stats) {
    uint32_t n = n_ptr ? *n_ptr : n_fixed;
    TraceCounters ctr = {};
    ClosestSrc src = {st, in_queue, sc.normals != nullptr};
    persistent_trace<TRACE_MODE_CLOSEST, STATS, false>(sc, src, n, cursor, refill, ctr);
    if (STATS) flush_counters(stats, ctr, false);
}

// intersect_shadow_ray for every queued NEE sample (integrators.cpp:756)
template <bool STATS>
__global__ void __launch_bounds__(128, BPT_TRACE_MIN_CTAS)
k_trace_shadow(DScene sc, DPathState st, const DShadowItem* __restrict__ items, const uint32_t* __restrict__ n_ptr,
               uint32_t* cursor, uint32_t refill, DStats* stats) {
    uint32_t n = *n_ptr;
    TraceCounters ctr = {};
    ShadowSrc src = {st, items};
    persistent_trace<TRACE_MODE_OCCLUSION, STATS, false>(sc, src, n, cursor, refill, ctr);
    if (STATS) flush_counters(stats, ctr, true);
}

// warp-aggregated append: returns this lane's index in the destination queue
BPT_D uint32_t queue_append(uint32_t* counter, bool want) {
    uint32_t mask = __ballot_sync(__activemask(), want);
    if (!want) return 0;
    uint32_t lane = threadIdx.x & 31;
    uint32_t leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

BPT_D uint32_t direction_octant(V3 d) { return (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u); }

// The reference's single-ray and ground-truth integrators (g_integrators[], integrators.cpp:823-830), one bounce of ONE
// path:  "Normals" (:544-561) and "Distances" (:563-580) end at their first hit; "Ground Truth Iterative" (:486-542)
// is the plain path tracer -- Fresnel reflection or uniform-hemisphere diffuse, no NEE, no roulette, one
// random_unilaterals() per hit.  They use the surface material as is (eta_i = 1, no material stack, no absorption).
BPT_D void shade_path_simple(const DScene& sc, const DPathState& st, const BatchDesc& b, uint32_t bounce, uint32_t slot,
                             bool& alive, uint32_t& octant) {
    const int integrator = sc.settings.integrator;
    float4 ro4 = PSL(st.ray_o + slot), rd4 = PSL(st.ray_d + slot), h4 = PSL(st.hit + slot);
    V3 ro = v3(ro4), rd = v3(rd4);
    HitRecord h;
    h.t = h4.x; h.prim = __float_as_uint(h4.y); h.tri = __float_as_uint(h4.z); h.v = h4.w; h.w = sc.normals ? st.hit_w[slot] : 0.0f;
    const bool first = bounce == 0;                       // state the ray generation did not write (see k_raygen)
    V3 throughput = first ? v3(1.0f) : v3(PSL(st.throughput + slot));
    float4 rad4 = first ? make_float4(0.0f, 0.0f, 0.0f, primary_vignette(sc, rd)) : PSL(st.radiance + slot);   // .w carries the vignette to the splat
    V3 total = v3(rad4);
    if (b.want_records) { float4 pd = st.primary_d[slot]; pd.w = __uint_as_float(__float_as_uint(pd.w) + 1u); st.primary_d[slot] = pd; }
    alive = false;

    if (h.prim == BPT_HIT_MISS) {
        if (integrator == BPT_INTEGRATOR_GT_ITERATIVE) total = total + throughput*sample_sky(sc, rd);            // :538
        else total = sample_sky(sc, rd);                                                                         // :559, :578
    } else {
        V3 I, N;
        uint32_t surface_id;
        hit_geometry(sc, ro, rd, h, I, N, surface_id);
        if (integrator == BPT_INTEGRATOR_NORMALS) {
            total = 0.5f*(v3(1.0f) + N);                                                                         // :557
        } else if (integrator == BPT_INTEGRATOR_DISTANCES) {
            total = v3(1.0f - clamp_t(h.t / 15.0f, 0.0f, 1.0f));                                                 // :576
        } else {
            MatView m = load_material(sc, surface_id);
            if (m.flags & BPT_MATERIAL_EMISSIVE) {
                total = total + throughput*m.emission;                                                           // :505-508
            } else {
                uint4 rng = first ? primary_rng_state(make_sampler(sc, b, slot), b) : PSL(st.rng + slot);
                next_set(rng);                                                                                   // random_unilaterals :510
                float rx = unilateral(rng.x);
                V2 ryz; ryz.x = unilateral(rng.y); ryz.y = unilateral(rng.z);
                float eta_i = 1.0f, eta_t = m.ior;
                float ratio = eta_i / eta_t;
                float cos_i = -dot(rd, N);
                float cos_t;
                float reflectance = fresnel_dielectric(cos_i, eta_i, eta_t, ratio, cos_t);
                V3 next_o, next_d;
                if (rx < reflectance) {                                                                          // :522-523
                    V3 refl = reflect(rd, N);
                    next_o = I + refl*kEps; next_d = refl;
                } else {                                                                                         // :525-533
                    V3 albedo = m.albedo;
                    if (m.flags & BPT_MATERIAL_CHECKERS) {
                        int checker = (((int)floorf(0.25f*I.x)) ^ ((int)floorf(0.25f*I.z))) & 1;
                        if (checker) albedo = m.checker;
                    }
                    V3 brdf = albedo*(1.0f / kPi);
                    throughput = throughput*brdf;
                    V3 R = map_to_hemisphere(N, ryz);
                    next_o = I + N*kEps; next_d = R;
                    throughput = throughput*dot(R, N);
                    throughput = throughput*(2.0f*kPi);
                }
                alive = bounce + 1 < sc.settings.max_bounce_count;
                if (alive) {
                    PSS(st.ray_o + slot, make_float4(next_o.x, next_o.y, next_o.z, 3.402823466e+38f));
                    PSS(st.ray_d + slot, make_float4(next_d.x, next_d.y, next_d.z, 0.0f));
                    PSS(st.rng + slot, rng);
                    PSS(st.throughput + slot, make_float4(throughput.x, throughput.y, throughput.z, 0.0f));
                    octant = direction_octant(next_d);
                }
            }
        }
    }
    PSS(st.radiance + slot, make_float4(total.x, total.y, total.z, rad4.w));
}

// One bounce of advanced_integrator (integrators.cpp:612-818) for ONE path: reads the path's state and the hit of its
// current ray, accumulates emission / sky, draws the next direction, and reports whether the path continues
// (its next ray is then in st.ray_o/ray_d) and whether it queued an NEE shadow ray (`sh`).
// PRE: compile the shadow-ray prefilter in (k_shade<true>, launched for scenes whose TLAS is one leaf -- elsewhere it
// decides next to nothing and its registers cost k_shade 3 %).
template <bool PRE>
BPT_D void shade_path(const DScene& sc, const DPathState& st, const BatchDesc& b, uint32_t bounce, uint32_t slot,
                      bool& alive, bool& want_shadow, DShadowItem& sh, uint32_t& octant, uint32_t& settled_shadow, uint32_t& settled_bytes) {
    const bpt_settings& set = sc.settings;
    if (set.integrator != BPT_INTEGRATOR_ADVANCED) { shade_path_simple(sc, st, b, bounce, slot, alive, octant); return; }
    float4 ro4 = PSL(st.ray_o + slot), rd4 = PSL(st.ray_d + slot), h4 = PSL(st.hit + slot);
    V3 ro = v3(ro4), rd = v3(rd4);
    HitRecord h;
    h.t = h4.x; h.prim = __float_as_uint(h4.y); h.tri = __float_as_uint(h4.z); h.v = h4.w; h.w = sc.normals ? st.hit_w[slot] : 0.0f;
    const bool first = bounce == 0;                       // state the ray generation did not write (see k_raygen)
    V3 throughput = first ? v3(1.0f) : v3(PSL(st.throughput + slot));
    float4 rad4 = first ? make_float4(0.0f, 0.0f, 0.0f, primary_vignette(sc, rd)) : PSL(st.radiance + slot);   // .w carries the vignette to the splat
    V3 total = v3(rad4);
    float4 pd = make_float4(0, 0, 0, 0);
    if (b.want_records) pd = st.primary_d[slot];
    uint32_t ray_count = __float_as_uint(pd.w) + 1u;      // this bounce's intersect_scene call
    bool total_changed = false;

    if (h.prim == BPT_HIT_MISS) {
        total = total + throughput*sample_sky(sc, rd);                                     // :813
        total_changed = true;
    } else {
        SamplerCtx sm = make_sampler(sc, b, slot);
        uint4 rng = first ? primary_rng_state(sm, b) : PSL(st.rng + slot);
        float4 pn4 = first ? make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(1u)) : PSL(st.prev_n + slot);   // is_specular_bounce starts true
        V3 prev_N = v3(pn4);
        bool is_specular = (__float_as_uint(pn4.w) & 1u) != 0;
        int stack_at = first ? 0 : (int)st.mstack_at[slot];

        V3 I, N;
        uint32_t surface_id;
        hit_geometry(sc, ro, rd, h, I, N, surface_id);
        float t = h.t;

        float cos_i = -dot(rd, N);                                                          // :618
        bool inside = (cos_i < 0.0f);
        uint32_t id_i, id_t;
        if (inside) {
            id_i = surface_id;
            int below = stack_at - 1; if (below < 0) below = 0;
            id_t = mstack_get(sc, st, b, slot, below);
            cos_i = -cos_i;
            N = -N;
        } else {
            id_i = mstack_get(sc, st, b, slot, stack_at);
            id_t = surface_id;
        }
        MatView mi = load_material(sc, id_i);
        MatView mt = load_material(sc, id_t);

        if (mi.medium) {                                                                     // :640-649 Beer
            // A zero coefficient (the integrator's own "air" is a medium with absorb = 0, :597-599) gives expf(-+0) = 1 and
            // x*1 = x bit for bit (t is the finite distance of a hit): the three double-precision exps are skipped for it.
            if (mi.absorb.x != 0.0f) throughput.x = throughput.x*exp_f(-mi.absorb.x*t);
            if (mi.absorb.y != 0.0f) throughput.y = throughput.y*exp_f(-mi.absorb.y*t);
            if (mi.absorb.z != 0.0f) throughput.z = throughput.z*exp_f(-mi.absorb.z*t);
        }

        if (mt.flags & BPT_MATERIAL_EMISSIVE) {                                              // :651-670
            bool allow_direct = (!set.next_event_estimation ||
                                 ((set.caustics || (bounce < 2)) && is_specular));
            if (allow_direct) {
                total = total + throughput*mt.emission;
                total_changed = true;
            } else if (bounce > 0 && set.use_mis) {
                float light_distance_sq = t*t;
                float light_pdf = light_distance_sq / cos_i;
                float brdf_pdf = (set.importance_sample_diffuse ? dot(prev_N, rd) / kPi : 1.0f / (2.0f*kPi));
                float mis_pdf = light_pdf + brdf_pdf;
                total = total + (1.0f / mis_pdf)*throughput*mt.emission;
                total_changed = true;
            }
        } else {
            alive = true;
            float eta_i = mi.ior, eta_t = mt.ior;
            float ratio = eta_i / eta_t;
            float cos_t;
            float reflectance = fresnel_dielectric(cos_i, eta_i, eta_t, ratio, cos_t);
            float reflect_test = sample_1d(sm, rng, Sample_Reflectance, bounce);
            reflectance = lerp_f(reflectance, 1.0f, mt.metallic);
            is_specular = true;
            V3 next_o, next_d;

            if (reflect_test < reflectance) {                                                // :684-696
                V3 refl = reflect(rd, N);
                if (mt.roughness > 0.0f) {
                    V3 rs;
                    do {                                                                     // random_in_unit_sphere :11-19
                        next_set(rng);
                        rs = v3(bilateral(rng.x), bilateral(rng.y), bilateral(rng.z));
                    } while (length_sq(rs) >= 1.0f);
                    refl = normalize((1.0f + kEps)*refl + mt.roughness*rs);
                }
                next_o = I + kEps*refl; next_d = refl;
                throughput = throughput*lerp_v(v3(1.0f), mt.albedo, mt.metallic);
            } else if (mt.medium) {                                                          // :698-717 refract
                if (inside) {
                    if (stack_at > 0) --stack_at;
                } else if (stack_at < (BPT_MATERIAL_STACK_DEPTH - 1)) {
                    ++stack_at;
                    st.mstack[(size_t)(stack_at - 1)*b.slots + slot] = (uint16_t)id_t;
                }
                V3 refr = ratio*rd + N*(ratio*cos_i - cos_t);
                next_o = I + refr*kEps; next_d = refr;
            } else {                                                                         // :719-790 diffuse
                is_specular = false;
                V3 albedo = mt.albedo;
                if (mt.flags & BPT_MATERIAL_CHECKERS) {
                    int checker = (((int)floorf(0.25f*I.x)) ^ ((int)floorf(0.25f*I.z))) & 1;
                    if (checker) albedo = mt.checker;
                }
                V3 brdf = (1.0f / kPi)*albedo;

                if (set.next_event_estimation && (sc.light_count > 0)) {
                    float pick = sample_1d(sm, rng, Sample_LightSelection, bounce);
                    // pick_random_light (:135-192)
                    uint32_t light_id = 0;
                    float pick_pdf = 0.0f;
                    if (set.importance_sample_lights) {
                        float sum = 0.0f;
                        for (uint32_t li = 0; li < sc.light_count; ++li) {
                            const DPrimitive* lp = sc.primitives + __ldg(&sc.lights[li]);
                            V3 lv = v3(__ldg(&lp->fwd[0].w), __ldg(&lp->fwd[1].w), __ldg(&lp->fwd[2].w)) - I;
                            float dsq = length_sq(lv);
                            const DMaterial* lm = sc.materials + __ldg(&lp->material);
                            float l = max3(v3(__ldg(&lm->emission_color[0]), __ldg(&lm->emission_color[1]), __ldg(&lm->emission_color[2])));
                            float r = __ldg(&lp->sphere_r);
                            float psa = (__ldg(&lp->type) == BPT_PRIM_SPHERE) ? (kPi*r*r / dsq) : 0.0f;
                            sum += l*psa;
                        }
                        float e = sum*pick;
                        float cdf = 0.0f, pdf = 0.0f;
                        uint32_t li = 0;
                        for (;; ++li) {
                            const DPrimitive* lp = sc.primitives + __ldg(&sc.lights[li]);
                            V3 lv = v3(__ldg(&lp->fwd[0].w), __ldg(&lp->fwd[1].w), __ldg(&lp->fwd[2].w)) - I;
                            float dsq = length_sq(lv);
                            const DMaterial* lm = sc.materials + __ldg(&lp->material);
                            float l = max3(v3(__ldg(&lm->emission_color[0]), __ldg(&lm->emission_color[1]), __ldg(&lm->emission_color[2])));
                            float r = __ldg(&lp->sphere_r);
                            float psa = (__ldg(&lp->type) == BPT_PRIM_SPHERE) ? (kPi*r*r / dsq) : 0.0f;
                            pdf = l*psa;
                            cdf = cdf + pdf;
                            if (!(cdf < e) || li + 1 >= sc.light_count) break;
                        }
                        pick_pdf = pdf / sum;
                        light_id = __ldg(&sc.lights[li]);
                    } else {
                        pick_pdf = 1.0f / (float)sc.light_count;
                        uint32_t li = (uint32_t)(pick*(float)sc.light_count - kEps);
                        light_id = __ldg(&sc.lights[li]);
                    }

                    V2 ds = sample_2d(sm, rng, Sample_DirectLighting, bounce);
                    const DPrimitive* lp = sc.primitives + light_id;
                    if (__ldg(&lp->type) == BPT_PRIM_SPHERE) {
                        // random_point_on_light (:199-228)
                        float4 f[3] = {__ldg(&lp->fwd[0]), __ldg(&lp->fwd[1]), __ldg(&lp->fwd[2])};
                        float lr = __ldg(&lp->sphere_r);
                        V3 light_p = v3(f[0].w, f[1].w, f[2].w);
                        V3 towards = normalize(light_p - I);
                        V3 Nl = map_to_hemisphere(-towards, ds);
                        V3 p = Nl*lr;
                        V3 p_world = xform(f, p, 1.0f);
                        V3 L = p_world - I;
                        float dist_sq = length_sq(L);
                        float dist = sqrtf(dist_sq);
                        L = L / dist;
                        float A = 2.0f*kPi*lr*lr;

                        float N_dot_L = dot(N, L);
                        float neg_Nl_dot_L = -dot(Nl, L);
                        if (N_dot_L > 0.0f && neg_Nl_dot_L > 0.0f) {
                            float solid_angle = (neg_Nl_dot_L*A) / dist_sq;
                            float pdf;
                            if (set.use_mis) {
                                float light_pdf = 1.0f / solid_angle;
                                float brdf_pdf = (set.importance_sample_diffuse ? N_dot_L / kPi : 1.0f / (2.0f*kPi));
                                pdf = light_pdf + brdf_pdf;
                            } else {
                                pdf = 1.0f / solid_angle;
                            }
                            pdf *= pick_pdf;
                            const DMaterial* lm = sc.materials + __ldg(&lp->material);
                            V3 emission = v3(__ldg(&lm->emission_color[0]), __ldg(&lm->emission_color[1]), __ldg(&lm->emission_color[2]));
                            V3 contrib = throughput*(dot(N, L) / pdf)*brdf*emission;
                            V3 so = I + L*kEps;
                            sh.o_maxt = make_float4(so.x, so.y, so.z, dist - 2*kEps);
                            sh.d_light = make_float4(L.x, L.y, L.z, __uint_as_float(light_id));
                            sh.contrib_slot = make_float4(contrib.x, contrib.y, contrib.z, __uint_as_float(slot));
                            want_shadow = true;
                            ray_count += 1u;
                        }
                    }
                }

                V2 is = sample_2d(sm, rng, Sample_IndirectLighting, bounce);
                V3 R;
                if (set.importance_sample_diffuse) {
                    R = map_to_cosine_weighted_hemisphere(N, is);
                    throughput = throughput*kPi;
                } else {
                    R = map_to_hemisphere(N, is);
                    throughput = throughput*(2.0f*kPi*dot(N, R));
                }
                throughput = throughput*brdf;
                next_o = I + N*kEps; next_d = R;
            }

            if (set.russian_roulette && !is_specular) {                                      // :801-811
                float p = clamp_t(max3(throughput), 0.1f, 0.9f);
                float e = sample_1d(sm, rng, Sample_Roulette, bounce);
                if (e > p) alive = false;
                else throughput = throughput*(1.0f / p);
            }

            if (bounce + 1 >= set.max_bounce_count) alive = false;
            if (alive) {
                PSS(st.ray_o + slot, make_float4(next_o.x, next_o.y, next_o.z, 3.402823466e+38f));
                PSS(st.ray_d + slot, make_float4(next_d.x, next_d.y, next_d.z, 0.0f));
                PSS(st.rng + slot, rng);
                PSS(st.prev_n + slot, make_float4(N.x, N.y, N.z, __uint_as_float(is_specular ? 1u : 0u)));
                st.mstack_at[slot] = (uint8_t)stack_at;
                PSS(st.throughput + slot, make_float4(throughput.x, throughput.y, throughput.z, 0.0f));
                octant = direction_octant(next_d);
            }
        }
    }
    if (PRE && want_shadow && sc.prefilter) {
        // settle the shadow ray here when it never reaches a BLAS (shadow_tlas_head): its contribution joins `total` behind this
        // bounce's other terms, which is where the traversal kernel's store would have added it
        uint32_t bytes;
        const int verdict = shadow_tlas_head(sc, v3(sh.o_maxt), v3(sh.d_light), sh.o_maxt.w, __float_as_uint(sh.d_light.w), bytes);
        if (verdict != 0) {
            settled_shadow = 1u; settled_bytes = bytes;
            if (sc.prefilter == 1u) {
                want_shadow = false;
                if (verdict == 1) { total = total + v3(sh.contrib_slot); total_changed = true; }
            }
        }
    }
    if (total_changed || first) PSS(st.radiance + slot, make_float4(total.x, total.y, total.z, rad4.w));
    if (b.want_records) { pd.w = __uint_as_float(ray_count); st.primary_d[slot] = pd; }
}

// One bounce for every active path of the batch (wavefront form of the loop at integrators.cpp:612-818).
// Survivors are appended to the next bounce's queue GROUPED BY THE OCTANT OF THEIR NEW DIRECTION: a block sorts its (up
// to 256) survivors by octant in shared memory and appends them with one atomic.  Neighbouring queue entries already
// start from neighbouring surface points (the queue follows pixel order); with equal direction signs they also take the
// same near/far decisions at every node, so the lanes of a traversal warp stay in the same phase for longer.  The order
// of the queue does not enter any per-path result.
#ifndef BPT_SHADE_THREADS
#define BPT_SHADE_THREADS 256     // 512: +2 % shade time (ncu: the sort's barriers are the top stall), 1024: +8 %
#endif
#ifndef BPT_SHADE_MIN_CTAS
#define BPT_SHADE_MIN_CTAS 4      // x 256 threads = 64 registers: measured 14.7 ms vs 18.6 ms at 128 registers on C2 (latency-bound on path state)
#endif
#ifndef BPT_SHADE_SORT
#define BPT_SHADE_SORT 1
#endif
#ifndef BPT_SHADE_SORT_BOUNCES
#define BPT_SHADE_SORT_BOUNCES 64     // bounces < this group their survivors by octant with the block-local sort; later ones append per warp
#endif
#ifndef BPT_SHADE_BLOCK_SHADOW
#define BPT_SHADE_BLOCK_SHADOW 1
#endif
#ifndef BPT_SHADE_SORT_MATERIAL
#define BPT_SHADE_SORT_MATERIAL 0     // measured on B200: shade time C2 +0.8 %, C3 -2.5 %, C4 +4 % -> off (the kernel is bound by
#endif                                // path-state traffic, not by branch divergence; parity tests pass with it on)
template <bool PRE>
__global__ void __launch_bounds__(BPT_SHADE_THREADS, BPT_SHADE_MIN_CTAS)
k_shade(DScene sc, DPathState st, BatchDesc b, uint32_t bounce,
        const uint32_t* __restrict__ in_queue, const uint32_t* __restrict__ n_ptr, uint32_t n_fixed,
        uint32_t* __restrict__ out_queue, uint32_t* out_count,
        DShadowItem* __restrict__ shadow_items, uint32_t* shadow_count, DStats* stats) {
    __shared__ uint32_t s_count[8], s_start[8], s_base;
    __shared__ uint32_t s_wshadow[BPT_SHADE_THREADS/32], s_shadow_base;     // shadow-queue append: per-warp counts, the block's base
    __shared__ uint32_t s_slots[BPT_SHADE_THREADS];
    uint32_t n = n_ptr ? *n_ptr : n_fixed;
    uint32_t n_rays = 0, n_shadow = 0;
    unsigned long long n_settled = 0, n_settled_bytes = 0;

    for (uint32_t blk0 = blockIdx.x*blockDim.x; blk0 < n; blk0 += gridDim.x*blockDim.x) {
        uint32_t i = blk0 + threadIdx.x;
        bool alive = false;          // path continues to the next bounce
        bool want_shadow = false;
        uint32_t slot = 0, octant = 0;
        DShadowItem sh;
        const bool block_sort = BPT_SHADE_SORT != 0 && bounce < (uint32_t)BPT_SHADE_SORT_BOUNCES;      // uniform over the grid
        if (block_sort) {
            if (threadIdx.x < 8) s_count[threadIdx.x] = 0;
            __syncthreads();
        }

        bool have = i < n;
        if (have) slot = in_queue ? in_queue[i] : i;
#if BPT_SHADE_SORT_MATERIAL
        // Sort-by-material: from the second bounce on the queue order is a mix of misses (sky lookup only), light hits
        // (path ends) and the materials of the scene.  The block re-deals its entries so that neighbouring lanes shade the
        // same kind of hit: key = miss | material id (folded to 3 bits), counting sort in shared memory.  The first
        // bounce runs in pixel order, which is coherent already.
        if (in_queue) {
            uint32_t key = 0;
            if (have) {
                uint32_t prim = __float_as_uint(PSL(st.hit + slot).y);
                if (prim != BPT_HIT_MISS) {
                    uint32_t mat = (prim & BPT_HIT_PLANE) ? __ldg(&sc.planes[prim & ~BPT_HIT_PLANE].material)
                                                          : __ldg(&sc.primitives[prim].material);
                    key = 1u + (mat % 7u);
                }
            }
            uint32_t r = 0;
            if (have) r = atomicAdd(&s_count[key], 1u);
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t run = 0;
                for (int k = 0; k < 8; ++k) { s_start[k] = run; run += s_count[k]; s_count[k] = 0; }
                s_base = run;                                    // entries this block holds in this round
            }
            __syncthreads();
            if (have) s_slots[s_start[key] + r] = slot;
            __syncthreads();
            have = threadIdx.x < s_base;
            if (have) slot = s_slots[threadIdx.x];
            __syncthreads();
        }
#endif
        uint32_t settled = 0u, settled_bytes = 0u;   // a shadow ray shade_path settled itself (counted as a traced ray all the same)
        if (have) shade_path<PRE>(sc, st, b, bounce, slot, alive, want_shadow, sh, octant, settled, settled_bytes);
        if (block_sort) {
            // block-local counting sort of the survivors by octant (rank within the octant from a shared-memory atomic)
            uint32_t rank = 0;
            if (alive) rank = atomicAdd(&s_count[octant], 1u);
            // the shadow queue is appended per block as well, in thread order: one global atomic per block instead of one
            // per warp on a single address (BPT_SHADE_BLOCK_SHADOW)
            const uint32_t shadow_mask = __ballot_sync(0xFFFFFFFFu, want_shadow);
            if (BPT_SHADE_BLOCK_SHADOW && (threadIdx.x & 31u) == 0u) s_wshadow[threadIdx.x >> 5] = __popc(shadow_mask);
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t run = 0;
                for (int k = 0; k < 8; ++k) { s_start[k] = run; run += s_count[k]; }
                s_base = run ? atomicAdd(out_count, run) : 0u;
                s_count[0] = run;                                    // total survivors of this round
                if (BPT_SHADE_BLOCK_SHADOW) {
                    uint32_t total_shadow = 0;
                    for (uint32_t wv = 0; wv < blockDim.x/32; ++wv) { uint32_t c = s_wshadow[wv]; s_wshadow[wv] = total_shadow; total_shadow += c; }
                    s_shadow_base = total_shadow ? atomicAdd(shadow_count, total_shadow) : 0u;
                }
            }
            __syncthreads();
            if (BPT_SHADE_BLOCK_SHADOW && want_shadow) {
                uint32_t si = s_shadow_base + s_wshadow[threadIdx.x >> 5] + __popc(shadow_mask & ((1u << (threadIdx.x & 31u)) - 1u));
                PSS(&shadow_items[si].o_maxt, sh.o_maxt); PSS(&shadow_items[si].d_light, sh.d_light); PSS(&shadow_items[si].contrib_slot, sh.contrib_slot);
            }
            if (alive) s_slots[s_start[octant] + rank] = slot;
            __syncthreads();
            uint32_t total = s_count[0];
            if (threadIdx.x < total) out_queue[s_base + threadIdx.x] = s_slots[threadIdx.x];
        } else {
            // no grouping: warp-aggregated append, no block barrier (the lanes of a late bounce take very different paths
            // through shade_path; ncu showed warps of a block waiting 7-10 issue slots per instruction at the sort's barriers)
            uint32_t qi = queue_append(out_count, alive);
            if (alive) out_queue[qi] = slot;
        }

        if (!(BPT_SHADE_BLOCK_SHADOW && block_sort)) {
            uint32_t si = queue_append(shadow_count, want_shadow);
            if (want_shadow) { PSS(&shadow_items[si].o_maxt, sh.o_maxt); PSS(&shadow_items[si].d_light, sh.d_light); PSS(&shadow_items[si].contrib_slot, sh.contrib_slot); }
        }
        const uint32_t settled_here = sc.prefilter == 1u ? settled : 0u;     // mode 2 only classifies: the ray is in the queue as well
        n_rays += (have ? 1u : 0u) + (want_shadow ? 1u : 0u) + settled_here;
        n_shadow += (want_shadow ? 1u : 0u) + settled_here;
        n_settled += settled; n_settled_bytes += settled_bytes;
        if (block_sort) __syncthreads();
    }
    flush_ray_counts(stats, n_rays, n_shadow);
    if (sc.prefilter == 2u) {             // the counting pass: how many shadow rays a timed pass settles in here, and their bytes
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) { n_settled += __shfl_xor_sync(0xFFFFFFFFu, n_settled, o); n_settled_bytes += __shfl_xor_sync(0xFFFFFFFFu, n_settled_bytes, o); }
        if ((threadIdx.x & 31) == 0 && n_settled) { atomicAdd(&stats->v[17], n_settled); atomicAdd(&stats->v[18], n_settled_bytes); }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused tail.  Once only a few paths of a batch are still alive, a wavefront bounce costs the latency of its longest
// ray four launches over (the machine is empty), bounce after bounce.  k_bounce_prepare (one thread, between the trace
// and the shade of a bounce) hands the survivors over when they fit: it moves the active count to counters[8] and
// zeroes it, so the remaining wavefront launches of the batch find empty queues.  k_tail then carries each surviving
// path to its end inside ONE launch: lane = path, and shading is simply what an idle lane does to get its next rays
// (persistent_trace in LOCAL mode): shade_path -- the same function k_shade runs -- then the lane's shadow ray, then
// its extension ray, then shade again ...  No lane waits for another path's bounce.  Per path the sequence of
// operations, and of float additions into its radiance, is the one the wavefront would have executed.
struct TailSrc {
    const DScene* sc;
    const BatchDesc* b;
    DPathState st;
    DShadowItem sh;
    uint32_t slot, bounce;
    uint32_t n_rays, n_shadow;
    bool alive;                  // the path's current extension ray has been traced and wants shading
    bool has_shadow, has_closest, cur_shadow, store_w;
    BPT_D bool pending() const { return has_shadow || has_closest || alive; }
    BPT_D bool next(V3& o, V3& d, float& max_t, uint32_t& ignored, bool& occ) {
        if (!has_shadow && !has_closest) {
            bool cont = false, want_shadow = false;
            uint32_t octant_unused = 0;
            uint32_t settled = 0u, settled_bytes_unused = 0u;
            shade_path<false>(*sc, st, *b, bounce, slot, cont, want_shadow, sh, octant_unused, settled, settled_bytes_unused);     // the few shadow rays of a tail all go through its own traversal
            n_rays += 1u + (want_shadow ? 1u : 0u) + settled;          // k_tail never runs in the counting mode (shadow_prefilter is 0 or 1 here)
            n_shadow += (want_shadow ? 1u : 0u) + settled;
            ++bounce;
            has_shadow = want_shadow; has_closest = cont; alive = cont;
            if (!want_shadow && !cont) return false;
        }
        if (has_shadow) {
            o = v3(sh.o_maxt); d = v3(sh.d_light); max_t = sh.o_maxt.w; ignored = __float_as_uint(sh.d_light.w);
            occ = true; has_shadow = false; cur_shadow = true;
        } else {
            float4 ro = PSL(st.ray_o + slot), rd = PSL(st.ray_d + slot);
            o = v3(ro); d = v3(rd); max_t = ro.w; ignored = 0u;
            occ = false; has_closest = false; cur_shadow = false;
        }
        return true;
    }
    BPT_D void store(uint32_t, const HitRecord& h) const {
        if (cur_shadow) {
            if (h.prim == BPT_HIT_MISS) {
                float4 r = PSL(st.radiance + slot);
                r.x = r.x + sh.contrib_slot.x; r.y = r.y + sh.contrib_slot.y; r.z = r.z + sh.contrib_slot.z;
                PSS(st.radiance + slot, r);
            }
        } else {
            PSS(st.hit + slot, make_float4(h.t, __uint_as_float(h.prim), __uint_as_float(h.tri), h.v));
            if (store_w) st.hit_w[slot] = h.w;
        }
    }
};

// Between the trace and the shade of a bounce (one thread): (1) the bookkeeping that used to be two launches -- zero the
// count of the queue the shade is about to fill, the shadow count (consumed by the merged trace, or not yet produced)
// and both fetch cursors (this bounce's trace and the previous bounce's shadow trace are done with them); (2) the
// hand-over to k_tail.  counters: the batch's DQueues::counters; in / out = active counts of this / the next bounce.
__global__ void k_bounce_prepare(uint32_t* counters, int in, int out, uint32_t threshold) {
    if (threadIdx.x == 0) {
        counters[out] = 0u; counters[2] = 0u; counters[3] = 0u; counters[4] = 0u;
        uint32_t n = counters[in];
        if (n != 0u && n <= threshold) { counters[8] = n; counters[in] = 0u; }
        else counters[8] = 0u;
    }
}

#ifndef BPT_TAIL_MIN_CTAS
#define BPT_TAIL_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(128, BPT_TAIL_MIN_CTAS)
k_tail(DScene sc, DPathState st, BatchDesc b, uint32_t first_bounce, const uint32_t* __restrict__ in_queue,
       const uint32_t* __restrict__ tail_count, uint32_t refill, DStats* stats) {
    const uint32_t n = *tail_count;
    const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i - (threadIdx.x & 31u) >= n) return;                   // whole warps only: the trace loop is warp-synchronous
    TailSrc src;
    src.sc = &sc; src.b = &b; src.st = st;
    src.alive = i < n;
    src.slot = src.alive ? in_queue[i] : 0u;
    src.bounce = first_bounce;
    src.n_rays = src.n_shadow = 0u;
    src.has_shadow = src.has_closest = src.cur_shadow = false;
    src.store_w = sc.normals != nullptr;
    TraceCounters ctr = {};
    persistent_trace<TRACE_MODE_MIXED, false, true>(sc, src, 0u, nullptr, refill, ctr);
    flush_ray_counts(stats, src.n_rays, src.n_shadow);
}

} // namespace bpt
#include "recursive.cuh"
namespace bpt {

// render_tile's tail (raytracer.cpp:469-488) + splat_filter (:187-259): one thread per pixel of the batch walks that
// pixel's samples, keeps the (2r+1)^2 footprint in registers, and flushes it with vector atomics.
template <int R>
__global__ void __launch_bounds__(128)
k_splat(DScene sc, DPathState st, BatchDesc b, float4* __restrict__ film) {
    constexpr int SPAN = 2*R + 1;
    uint32_t pixels = b.rect_w*b.rows;
    for (uint32_t pix = blockIdx.x*blockDim.x + threadIdx.x; pix < pixels; pix += gridDim.x*blockDim.x) {
        int x = b.x0 + (int)(pix % b.rect_w), y = __ldg(&b.row_map[b.row0 + pix / b.rect_w]);
        float4 acc[SPAN*SPAN];
        #pragma unroll
        for (int k = 0; k < SPAN*SPAN; ++k) acc[k] = make_float4(0, 0, 0, 0);
        float kernel_scale = (float)(sc.filter_lut_size - 1) / (float)sc.filter_radius;
        for (uint32_t s = 0; s < b.S; ++s) {
            uint32_t slot = pix*b.S + s;
            float4 rad = PSL(st.radiance + slot);
            float vig = rad.w;
            float2 j = PSL(st.jitter + slot);
            V3 c = v3(rad)*vig;
            float wx[SPAN], wy[SPAN];
            #pragma unroll
            for (int i = 0; i < SPAN; ++i) {
                int ix = (int)fabsf(0.5f + kernel_scale*((float)(i - R) - j.x));
                int iy = (int)fabsf(0.5f + kernel_scale*((float)(i - R) - j.y));
                wx[i] = __ldg(&sc.filter_lut[ix]);
                wy[i] = __ldg(&sc.filter_lut[iy]);
            }
            #pragma unroll
            for (int yy = 0; yy < SPAN; ++yy) {
                #pragma unroll
                for (int xx = 0; xx < SPAN; ++xx) {
                    float f = wx[xx]*wy[yy];
                    float4& a = acc[yy*SPAN + xx];
                    a.x = a.x + f*c.x; a.y = a.y + f*c.y; a.z = a.z + f*c.z; a.w = a.w + f;
                }
            }
        }
        #pragma unroll
        for (int yy = 0; yy < SPAN; ++yy) {
            int py = y + yy - R;
            if (py < 0 || py >= (int)sc.film_h) continue;
            #pragma unroll
            for (int xx = 0; xx < SPAN; ++xx) {
                int px = x + xx - R;
                if (px < 0 || px >= (int)sc.film_w) continue;
                atomicAdd(&film[(size_t)py*sc.film_w + px], acc[yy*SPAN + xx]);
            }
        }
    }
}

// generic-radius splat (filters up to r = 12) and the Box path (raytracer.cpp:485-488); one thread per sample
__global__ void __launch_bounds__(128)
k_splat_generic(DScene sc, DPathState st, BatchDesc b, float4* __restrict__ film) {
    int R = (int)sc.filter_radius;
    for (uint32_t slot = blockIdx.x*blockDim.x + threadIdx.x; slot < b.slots; slot += gridDim.x*blockDim.x) {
        uint32_t pix = slot / b.S;
        int x = b.x0 + (int)(pix % b.rect_w), y = __ldg(&b.row_map[b.row0 + pix / b.rect_w]);
        float4 rad = PSL(st.radiance + slot);
        float vig = rad.w;
        V3 c = v3(rad)*vig;
        if (sc.filter_lut_size == 0) {
            atomicAdd(&film[(size_t)y*sc.film_w + x], make_float4(c.x, c.y, c.z, 1.0f));
            continue;
        }
        float2 j = PSL(st.jitter + slot);
        float kernel_scale = (float)(sc.filter_lut_size - 1) / (float)sc.filter_radius;
        for (int yy = 0; yy <= 2*R; ++yy) {
            int py = y + yy - R;
            if (py < 0 || py >= (int)sc.film_h) continue;
            float fy = __ldg(&sc.filter_lut[(int)fabsf(0.5f + kernel_scale*((float)(yy - R) - j.y))]);
            for (int xx = 0; xx <= 2*R; ++xx) {
                int px = x + xx - R;
                if (px < 0 || px >= (int)sc.film_w) continue;
                float fx = __ldg(&sc.filter_lut[(int)fabsf(0.5f + kernel_scale*((float)(xx - R) - j.x))]);
                float f = fx*fy;
                atomicAdd(&film[(size_t)py*sc.film_w + px], make_float4(f*c.x, f*c.y, f*c.z, f));
            }
        }
    }
}

__global__ void k_write_records(DPathState st, BatchDesc b, bpt_sample_record* __restrict__ out) {
    for (uint32_t slot = blockIdx.x*blockDim.x + threadIdx.x; slot < b.slots; slot += gridDim.x*blockDim.x) {
        float4 o = st.primary_o[slot], d = st.primary_d[slot], r = PSL(st.radiance + slot);
        bpt_sample_record rec;
        rec.ray_o[0] = o.x; rec.ray_o[1] = o.y; rec.ray_o[2] = o.z;
        rec.ray_d[0] = d.x; rec.ray_d[1] = d.y; rec.ray_d[2] = d.z;
        rec.radiance[0] = r.x; rec.radiance[1] = r.y; rec.radiance[2] = r.z;
        rec.rays = __float_as_uint(d.w);
        out[slot] = rec;
    }
}

// bpt_trace: host ray batch -> bpt_hit (diagnostics / parity), closest or occlusion, through the same persistent loop
template <bool OCC>
struct ApiSrc {
    DScene sc;
    const bpt_ray* rays;
    bpt_hit* out;
    const uint32_t* tri_original;
    uint32_t ignored;
    BPT_D void load(uint32_t i, V3& o, V3& d, float& max_t, uint32_t& ign, bool& occ) const {
        bpt_ray r = rays[i];
        o = v3(r.o); d = v3(r.d); max_t = r.max_t; ign = OCC ? ignored : 0u;
        occ = OCC;
    }
    BPT_D void store(uint32_t i, const HitRecord& h) const {
        bpt_ray r = rays[i];
        bpt_hit res;
        res.t = OCC ? r.max_t : h.t;
        res.primitive = h.prim;
        res.triangle = 0xFFFFFFFFu;
        res.n[0] = res.n[1] = res.n[2] = 0.0f;
        res.p[0] = res.p[1] = res.p[2] = 0.0f;
        if (!OCC && h.prim != BPT_HIT_MISS) {
            V3 I, N; uint32_t mat;
            hit_geometry(sc, v3(r.o), v3(r.d), h, I, N, mat);
            res.n[0] = N.x; res.n[1] = N.y; res.n[2] = N.z;
            res.p[0] = I.x; res.p[1] = I.y; res.p[2] = I.z;
            if (h.tri != 0xFFFFFFFFu && !(h.prim & BPT_HIT_PLANE) && sc.primitives[h.prim].type == BPT_PRIM_MESH) {
                res.triangle = tri_original[h.tri];
            }
        }
        out[i] = res;
    }
};

template <bool OCC, bool STATS>
__global__ void __launch_bounds__(128)
k_trace_api(DScene sc, const bpt_ray* __restrict__ rays, uint32_t n, uint32_t ignored, bpt_hit* __restrict__ out,
            const uint32_t* __restrict__ tri_original, uint32_t* cursor, uint32_t refill, DStats* stats) {
    TraceCounters ctr = {};
    ApiSrc<OCC> src = {sc, rays, out, tri_original, ignored};
    persistent_trace<OCC ? TRACE_MODE_OCCLUSION : TRACE_MODE_CLOSEST, STATS, false>(sc, src, n, cursor, refill, ctr);
    if (STATS) flush_counters(stats, ctr, OCC);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&stats->v[0], (unsigned long long)n);
        if (OCC) atomicAdd(&stats->v[1], (unsigned long long)n);
    }
}

// the display loop's per-pixel resolve (raytracer.cpp:2113-2172)
BPT_D float sigmoidal_contrast(float x, float contrast, float midpoint) {       // raytracer.cpp:69-84
    float curve;
    if (x < midpoint) {
        float scale = (1.0f / midpoint)*x;
        curve = midpoint*(scale*scale);
    } else {
        float y = (1.0f / (1.0f - midpoint));
        float scale = y - y*x;
        curve = 1.0f - (1.0f - midpoint)*(scale*scale);
    }
    return lerp_f(x, curve, contrast);
}

BPT_D float remap_tpdf(float x) {                                                 // raytracer.cpp:125-132
    float orig = 2.0f*x - 1.0f;
    // the reference uses the SSE rsqrt approximation (12-bit); an exact 1/sqrt differs by < 1e-3 -> at most 1 LSB after dither
    x = orig*rsqrtf(fabsf(orig));
    x = max_t(-1.0f, x);
    x = x - sign_of(x);
    return x;
}

__global__ void k_resolve(const float4* __restrict__ film, uint32_t w, uint32_t h, bpt_post_settings post,
                          const uint8_t* __restrict__ dither, uint32_t dw, uint32_t dh, uint32_t* __restrict__ out) {
    uint32_t n = w*h;
    for (uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) {
        float4 s = film[i];
        V3 c = v3(0.0f);
        if ((s.x != s.x) || (s.y != s.y) || (s.z != s.z) || (s.w != s.w)) {
            c = v3(0.0f, 255.0f, 255.0f);
        } else if (s.w > 0.001f) {
            c = v3(s.x, s.y, s.z) / s.w;
            c = v3(max_t(c.x, 0.0f), max_t(c.y, 0.0f), max_t(c.z, 0.0f));
            if (post.exposure != 0.0f) c = c*pow_f(2.0f, post.exposure);
            if (post.tonemapping) c = v3(1.0f - exp_f(-c.x), 1.0f - exp_f(-c.y), 1.0f - exp_f(-c.z));
            if (post.srgb_transform) c = v3(pow_f(c.x, 1.0f / 2.23333f), pow_f(c.y, 1.0f / 2.23333f), pow_f(c.z, 1.0f / 2.23333f));
            if (post.contrast != 0.0f) c = v3(sigmoidal_contrast(c.x, post.contrast, post.midpoint),
                                              sigmoidal_contrast(c.y, post.contrast, post.midpoint),
                                              sigmoidal_contrast(c.z, post.contrast, post.midpoint));
            c = c*255.0f;
            if (dither) {
                uint32_t x = i % w, y = i / w;
                const uint8_t* d = dither + ((size_t)(y & (dh - 1))*dw + (x & (dw - 1)))*3;
                c = c + v3(0.5f + remap_tpdf((1.0f / 255.0f)*(float)d[0]),
                           0.5f + remap_tpdf((1.0f / 255.0f)*(float)d[1]),
                           0.5f + remap_tpdf((1.0f / 255.0f)*(float)d[2]));
            }
        } else if (s.w < -0.01f) {
            c = v3(-255.0f*s.w, 0.0f, -255.0f*s.w);
        }
        uint32_t r = (uint32_t)(uint8_t)clamp_t(c.x, 0.0f, 255.0f);
        uint32_t g = (uint32_t)(uint8_t)clamp_t(c.y, 0.0f, 255.0f);
        uint32_t b = (uint32_t)(uint8_t)clamp_t(c.z, 0.0f, 255.0f);
        out[i] = (255u << 24) | (r << 16) | (g << 8) | b;
    }
}

// upload-time re-layout on the device: leaf-ordered 36-byte triangles -> 48-byte {a, idx}{b-a}{c-a} records (the two
// edge subtractions are the first operations of ray_intersect_triangle, intersection.cpp:145-146: same IEEE ops, done
// once here), and caller-order vertex normals -> leaf order.
__global__ void k_build_triangles(const float* __restrict__ raw, const uint32_t* __restrict__ original, uint32_t n,
                                  DTriangle* __restrict__ out, const float* __restrict__ raw_normals, float4* __restrict__ out_normals) {
    for (uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) {
        const float* p = raw + (size_t)i*9;
        uint32_t orig = original[i];
        DTriangle t;
        t.a_idx = make_float4(p[0], p[1], p[2], __uint_as_float(orig));
        t.e1 = make_float4(p[3] - p[0], p[4] - p[1], p[5] - p[2], 0.0f);
        t.e2 = make_float4(p[6] - p[0], p[7] - p[1], p[8] - p[2], 0.0f);
        out[i] = t;
        if (raw_normals) {
            const float* nn = raw_normals + (size_t)orig*9;
            out_normals[(size_t)i*3 + 0] = make_float4(nn[0], nn[1], nn[2], 0.0f);
            out_normals[(size_t)i*3 + 1] = make_float4(nn[3], nn[4], nn[5], 0.0f);
            out_normals[(size_t)i*3 + 2] = make_float4(nn[6], nn[7], nn[8], 0.0f);
        }
    }
}

__global__ void k_expand_rgb(const float* __restrict__ rgb, uint32_t n, float4* __restrict__ out) {
    for (uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x)
        out[i] = make_float4(rgb[(size_t)i*3], rgb[(size_t)i*3 + 1], rgb[(size_t)i*3 + 2], 0.0f);
}

__global__ void k_reset_counters(uint32_t* counters, int which_mask) {
    if (threadIdx.x < 8 && (which_mask >> threadIdx.x) & 1) counters[threadIdx.x] = 0;
}

} // namespace bpt
