// The reference's two RECURSIVE integrators, "Whitted" (raytrace_recursively, integrators.cpp:310-415) and
// "Ground Truth Recursive" (pathtrace_recursively, :431-472), on the device.
//
// A recursive call tree does not fit a wavefront (one sample spawns up to 2^depth rays whose results combine in
// call order), so each SAMPLE becomes a small coroutine owned by one lane: an explicit frame stack replaces the C++
// call stack, and "I need a ray traced" suspends the coroutine -- the lane then takes part in the warp's
// persistent_trace loop (LOCAL mode, trace.cuh) like any other ray owner, and resumes when its hit is back.
// The evaluation order is the reference's depth-first order, so the sampler / RNG draws happen in the same sequence and
// every float expression keeps its association: results are compared per sample with the reference (tests).
// Included from kernels.cuh (needs hit_geometry, load_material, sample_sky, fresnel_dielectric, map_to_hemisphere).
#pragma once

namespace bpt {

#define BPT_MAX_RECURSION 32      // the UI's maximum for max_bounce_count (raytracer.cpp:1972); deeper settings are refused

struct RecursiveFrame {           // a suspended caller
    V3 a;                         // Whitted medium: throughput            | Whitted reflective: diffuse_light   | GT: brdf
    V3 b;                         // Whitted medium: refracted_light (once known) | reflective: metallic_color
    V3 refl_o, refl_d;            // Whitted medium: the reflected child ray, launched after the refracted subtree
    float f;                      // Whitted: reflectance                  | GT: max(0, N.R)
    int kind;                     // see K_* below
};

template <bool WHITTED>
struct RecursiveSrc {
    enum { K_MEDIUM_REFRACTED = 0, K_MEDIUM_REFLECTED = 1, K_REFLECTIVE = 2, K_GT_PASSTHROUGH = 3, K_GT_DIFFUSE = 4 };
    enum { ST_ENTER, ST_HIT, ST_LIGHT_NEXT, ST_LIGHT_RESULT, ST_AFTER_LIGHTS, ST_RETURN, ST_DONE };
    static constexpr uint32_t NO_MATERIAL = 0xFFFFFFFFu;

    const DScene* sc;
    const BatchDesc* b;
    DPathState st;
    uint32_t slot;
    SamplerCtx sm;
    uint4 rng;
    int state;
    int depth;                    // frames in use == recursion levels above the call being evaluated
    uint32_t max_depth;           // scene->settings.max_bounce_count
    uint32_t n_rays, n_shadow;
    HitRecord hit;                // result of the ray this lane asked for last
    // the call being evaluated: (ray, previous_material)
    V3 ray_o, ray_d;
    uint32_t prev_mat;
    // its locals that live across the shadow rays of the light loop (Whitted)
    V3 I, N, throughput, illumination, pending_light;
    uint32_t mat_id, light_index;
    float cos_i, eta_i, eta_t;
    // the value being returned up the stack
    V3 value;
    RecursiveFrame frames[BPT_MAX_RECURSION];

    BPT_D void init(const DScene& scene, const DPathState& state_arrays, const BatchDesc& batch, uint32_t s, bool valid) {
        sc = &scene; b = &batch; st = state_arrays; slot = s;
        state = valid ? ST_ENTER : ST_DONE;
        depth = 0; max_depth = scene.settings.max_bounce_count;
        n_rays = n_shadow = 0u;
        prev_mat = NO_MATERIAL;
        if (valid) {
            sm = make_sampler(scene, batch, s);
            rng = primary_rng_state(sm, batch);             // k_raygen leaves everything but the ray to its consumers
            ray_o = v3(st.ray_o[s]); ray_d = v3(st.ray_d[s]);
        }
    }

    BPT_D bool pending() const { return state != ST_DONE; }

    BPT_D void store(uint32_t, const HitRecord& h) { hit = h; }

    BPT_D void emit_closest(V3& o, V3& d, float& max_t, uint32_t& ignored, bool& occ) {
        o = ray_o; d = ray_d; max_t = 3.402823466e+38f; ignored = 0u; occ = false;
        n_rays += 1u;
        state = ST_HIT;
    }

    BPT_D V3 reflected_direction(const MatView& m) {                                        // :385-388, :395-398
        V3 refl = reflect(ray_d, N);
        if (m.roughness > 0.0f) {
            V3 rs;
            do {                                                                            // random_in_unit_sphere :11-19
                next_set(rng);
                rs = v3(bilateral(rng.x), bilateral(rng.y), bilateral(rng.z));
            } while (length_sq(rs) >= 1.0f);
            refl = normalize((1.0f + kEps)*refl + m.roughness*rs);
        }
        return refl;
    }

    BPT_D V3 evaluate_material(const MatView& m) const {                                    // :297-308
        V3 albedo = m.albedo;
        if (m.flags & BPT_MATERIAL_CHECKERS) {
            int checker = (((int)floorf(0.25f*I.x)) ^ ((int)floorf(0.25f*I.z))) & 1;
            if (checker) albedo = m.checker;
        }
        return albedo;
    }

    // Runs the coroutine until it needs a ray (returns true with the ray) or the sample is finished (false).
    BPT_D bool next(V3& o, V3& d, float& max_t, uint32_t& ignored, bool& occ) {
        for (;;) {
            switch (state) {
            case ST_ENTER: {
                if ((uint32_t)depth >= max_depth) {                                         // recursion == 0
                    value = WHITTED ? v3(0.0f) : sample_sky(*sc, ray_d);                    // :414 | :471
                    state = ST_RETURN;
                    break;
                }
                emit_closest(o, d, max_t, ignored, occ);
                return true;
            }
            case ST_HIT: {
                if (hit.prim == BPT_HIT_MISS) { value = sample_sky(*sc, ray_d); state = ST_RETURN; break; }   // :411 | :471
                uint32_t surface_id;
                hit_geometry(*sc, ray_o, ray_d, hit, I, N, surface_id);
                MatView m = load_material(*sc, surface_id);
                if (m.flags & BPT_MATERIAL_EMISSIVE) { value = m.emission; state = ST_RETURN; break; }        // :320 | :440
                if (WHITTED) {
                    cos_i = -dot(ray_d, N);                                                 // :324-345
                    bool inside = (cos_i < 0.0f);
                    eta_i = 1.0f; eta_t = m.ior;
                    throughput = v3(1.0f);
                    mat_id = surface_id;
                    if (inside) {
                        N = -N; cos_i = -cos_i;
                        float tmp = eta_i; eta_i = eta_t; eta_t = tmp;
                        if (prev_mat != NO_MATERIAL) { mat_id = prev_mat; m = load_material(*sc, mat_id); }
                    }
                    if (inside && m.medium) {
                        throughput.x = throughput.x*exp_f(-m.absorb.x*hit.t);
                        throughput.y = throughput.y*exp_f(-m.absorb.y*hit.t);
                        throughput.z = throughput.z*exp_f(-m.absorb.z*hit.t);
                    }
                    illumination = v3(0.0f);
                    light_index = 0u;
                    state = ST_LIGHT_NEXT;
                } else {
                    next_set(rng);                                                          // random_unilaterals :443
                    float rx = unilateral(rng.x);
                    V2 ryz; ryz.x = unilateral(rng.y); ryz.y = unilateral(rng.z);
                    float e_i = 1.0f, e_t = m.ior;
                    float ratio = e_i / e_t;
                    float c_i = -dot(ray_d, N);
                    float c_t;
                    float reflectance = fresnel_dielectric(c_i, e_i, e_t, ratio, c_t);
                    RecursiveFrame& fr = frames[depth];
                    if (rx < reflectance) {                                                 // :452-454
                        V3 refl = reflect(ray_d, N);
                        fr.kind = K_GT_PASSTHROUGH;
                        ray_o = I + refl*kEps; ray_d = refl;
                    } else {                                                                // :456-467
                        fr.a = evaluate_material(m)*(1.0f / kPi);
                        V3 R = map_to_hemisphere(N, ryz);
                        fr.f = bpt::max_t(0.0f, dot(N, R));
                        fr.kind = K_GT_DIFFUSE;
                        ray_o = I + N*kEps; ray_d = R;
                    }
                    ++depth;
                    state = ST_ENTER;
                }
                break;
            }
            case ST_LIGHT_NEXT: {                                                           // :349-368, one light per visit
                if (light_index >= sc->light_count) { state = ST_AFTER_LIGHTS; break; }
                uint32_t light_id = __ldg(&sc->lights[light_index]);
                const DPrimitive* lp = sc->primitives + light_id;
                V2 ds = sample_2d(sm, rng, Sample_DirectLighting, 0u);                      // bounce index 0, always (:354)
                bool shoot = false;
                if (__ldg(&lp->type) == BPT_PRIM_SPHERE) {                                  // random_point_on_light :199-228
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
                        const DMaterial* lm = sc->materials + __ldg(&lp->material);
                        V3 emission = v3(__ldg(&lm->emission_color[0]), __ldg(&lm->emission_color[1]), __ldg(&lm->emission_color[2]));
                        pending_light = (neg_Nl_dot_L*A*N_dot_L*emission) / dist_sq;              // :364
                        V3 so = I + L*kEps;
                        o = so; d = L; max_t = dist - 2*kEps; ignored = light_id; occ = true;
                        shoot = true;
                    }
                }
                if (shoot) { n_rays += 1u; n_shadow += 1u; state = ST_LIGHT_RESULT; return true; }
                ++light_index;
                break;
            }
            case ST_LIGHT_RESULT: {
                if (hit.prim == BPT_HIT_MISS) illumination = illumination + pending_light;
                ++light_index;
                state = ST_LIGHT_NEXT;
                break;
            }
            case ST_AFTER_LIGHTS: {                                                         // :370-408
                MatView m = load_material(*sc, mat_id);
                illumination = illumination*(1.0f / 1.0f);
                illumination = illumination + v3(sc->ambient_light);
                V3 brdf = (1.0f / kPi)*evaluate_material(m);
                V3 metallic_color = lerp_v(v3(1.0f), m.albedo, m.metallic);
                float ratio = eta_i / eta_t;
                float cos_t;
                float reflectance = fresnel_dielectric(cos_i, eta_i, eta_t, ratio, cos_t);
                reflectance = lerp_f(reflectance, 1.0f, m.metallic);
                RecursiveFrame& fr = frames[depth];
                if (m.medium) {
                    V3 refracted_d = ratio*ray_d + N*(ratio*cos_i - cos_t);                 // refract :260-264
                    V3 reflected_d = reflected_direction(m);
                    fr.kind = K_MEDIUM_REFRACTED;
                    fr.a = throughput; fr.f = reflectance;
                    fr.refl_o = I + reflected_d*kEps; fr.refl_d = reflected_d;
                    ray_o = I + refracted_d*kEps; ray_d = refracted_d;
                    prev_mat = mat_id;
                    ++depth;
                    state = ST_ENTER;
                } else if (reflectance > 0.05f) {
                    V3 reflected_d = reflected_direction(m);
                    fr.kind = K_REFLECTIVE;
                    fr.a = throughput*brdf*illumination;                                    // diffuse_light
                    fr.b = metallic_color; fr.f = reflectance;
                    ray_o = I + reflected_d*kEps; ray_d = reflected_d;
                    prev_mat = NO_MATERIAL;
                    ++depth;
                    state = ST_ENTER;
                } else {
                    value = throughput*brdf*illumination;
                    state = ST_RETURN;
                }
                break;
            }
            case ST_RETURN: {
                if (depth == 0) {
                    st.radiance[slot] = make_float4(value.x, value.y, value.z, primary_vignette(*sc, v3(st.ray_d[slot])));   // .w: vignette for the splat
                    if (b->want_records) { float4 pd = st.primary_d[slot]; pd.w = __uint_as_float(n_rays); st.primary_d[slot] = pd; }
                    state = ST_DONE;
                    return false;
                }
                --depth;
                RecursiveFrame& fr = frames[depth];
                if (fr.kind == K_MEDIUM_REFRACTED) {                                        // :390-391: now the reflected subtree
                    fr.b = value;
                    fr.kind = K_MEDIUM_REFLECTED;
                    ray_o = fr.refl_o; ray_d = fr.refl_d;
                    prev_mat = NO_MATERIAL;
                    ++depth;
                    state = ST_ENTER;
                } else if (fr.kind == K_MEDIUM_REFLECTED) {
                    value = lerp_v(fr.a*fr.b, value, fr.f);                                 // :393
                } else if (fr.kind == K_REFLECTIVE) {
                    value = lerp_v(fr.a, fr.b*value, fr.f);                                 // :400-403
                } else if (fr.kind == K_GT_DIFFUSE) {
                    V3 light = value*fr.f;                                                  // :464-467
                    value = ((2.0f*kPi)*light)*fr.a;
                }                                                                           // K_GT_PASSTHROUGH: value unchanged
                break;
            }
            default:
                return false;
            }
        }
    }
};

// One launch evaluates every sample of a batch to completion (the primary rays come from k_raygen).
template <bool WHITTED>
__global__ void __launch_bounds__(128)
k_recursive(DScene sc, DPathState st, BatchDesc b, uint32_t refill, DStats* stats) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x*blockDim.x) >> 5;
    uint32_t n_rays = 0, n_shadow = 0;
    TraceCounters ctr = {};
    RecursiveSrc<WHITTED> src;
    for (uint32_t base = ((blockIdx.x*blockDim.x + threadIdx.x) >> 5)*32u; base < b.slots; base += warps*32u) {
        uint32_t slot = base + lane;
        src.init(sc, st, b, slot, slot < b.slots);
        persistent_trace<TRACE_MODE_MIXED, false, true>(sc, src, 0u, nullptr, refill, ctr);
        n_rays += src.n_rays; n_shadow += src.n_shadow;
    }
    flush_ray_counts(stats, n_rays, n_shadow);
}

} // namespace bpt
