// Device restatement of the ~30 MathLib / sampler functions the hot path touches, op-for-op
// (MathLib/my_math.h, Raytracer/samplers.h, Raytracer/samplers.cpp:18-138).
//
// Numerics contract: this translation unit is compiled with --fmad=false and the
// default IEEE div/sqrt, so every + - * / sqrt below rounds exactly like the oracle's SSE scalar code
// (g++ -O2 -ffp-contract=off).  min/max are the reference's ternaries, not fminf/fmaxf.  libm calls
// (sin/cos/exp/atan2/asin) are evaluated in double and rounded once to float, which agrees with glibc's
// nearly-correctly-rounded float routines except on rare 1-ulp cases (stated in the radiance tolerance).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bpt {

#define BPT_D __device__ __forceinline__

constexpr float kPi  = 3.14159265359f;   // my_math.h:15
constexpr float kTau = 6.28318530717f;   // my_math.h:16
constexpr float kEps = 0.001f;           // common.h:35

struct V3 { float x, y, z; };
struct V2 { float x, y; };

BPT_D V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
BPT_D V3 v3(float x) { return v3(x, x, x); }
BPT_D V3 v3(const float* p) { return v3(p[0], p[1], p[2]); }
BPT_D V3 v3(float4 q) { return v3(q.x, q.y, q.z); }
BPT_D V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
BPT_D V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
BPT_D V3 operator*(V3 a, V3 b) { return v3(a.x*b.x, a.y*b.y, a.z*b.z); }
BPT_D V3 operator/(V3 a, V3 b) { return v3(a.x/b.x, a.y/b.y, a.z/b.z); }
BPT_D V3 operator*(V3 a, float b) { return v3(a.x*b, a.y*b, a.z*b); }
BPT_D V3 operator*(float a, V3 b) { return v3(a*b.x, a*b.y, a*b.z); }
BPT_D V3 operator/(V3 a, float b) { return v3(a.x/b, a.y/b, a.z/b); }
BPT_D V3 operator/(float a, V3 b) { return v3(a/b.x, a/b.y, a/b.z); }
BPT_D V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }

BPT_D float min_t(float a, float b) { return a < b ? a : b; }     // my_math.h:77-85
BPT_D float max_t(float a, float b) { return a > b ? a : b; }
BPT_D float clamp_t(float n, float a, float b) { return max_t(a, min_t(b, n)); }   // :107-110
BPT_D float lerp_f(float a, float b, float t) { return a*(1.0f - t) + b*t; }       // :69-72
BPT_D V3 lerp_v(V3 a, V3 b, float t) { return a*(1.0f - t) + b*t; }                // :448-451
BPT_D float dot(V3 a, V3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }                // :453-456
BPT_D V3 cross(V3 a, V3 b) { return v3(a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x); }
BPT_D V3 reflect(V3 v, V3 n) { return v - 2.0f*dot(v, n)*n; }                      // :471-474
BPT_D float length_sq(V3 a) { return dot(a, a); }
BPT_D V3 normalize(V3 a) { float rcp = 1.0f / sqrtf(dot(a, a)); return a*rcp; }    // :486-490
BPT_D V3 noz(V3 a) {                                                               // :492-500
    V3 r = v3(0.0f);
    float lsq = dot(a, a);
    if ((lsq > 0.0001f) && (lsq < __int_as_float(0x7f800000))) r = a / sqrtf(lsq);
    return r;
}
BPT_D V3 abs_v(V3 a) { return v3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
BPT_D float max3(V3 a) { return max_t(a.x, max_t(a.y, a.z)); }                     // :541-545
BPT_D float copy_sign(float value_of, float sign_of) {                             // :185-194
    return __int_as_float((__float_as_int(sign_of) & 0x80000000) | (__float_as_int(value_of) & 0x7fffffff));
}
BPT_D float sign_of(float x) { return x < 0.0f ? -1.0f : 1.0f; }                      // :167-170

// transform / transform_normal on the 3 stored rows (my_math.h:947-963; note :956-963 is the reference's own,
// not-quite-transpose formula -- reproduced verbatim)
BPT_D V3 xform(const float4* m, V3 p, float pw) {
    return v3(p.x*m[0].x + p.y*m[0].y + p.z*m[0].z + pw*m[0].w,
              p.x*m[1].x + p.y*m[1].y + p.z*m[1].z + pw*m[1].w,
              p.x*m[2].x + p.y*m[2].y + p.z*m[2].z + pw*m[2].w);
}
BPT_D V3 xform_normal(const float4* m, V3 n) {
    return v3(n.x*m[0].x + n.y*m[0].y + n.z*m[2].x,
              n.x*m[0].y + n.y*m[1].y + n.z*m[2].y,
              n.x*m[0].z + n.y*m[1].z + n.z*m[2].z);
}

// libm stand-ins: one rounding from a double evaluation
BPT_D float sin_f(float x)  { return (float)sin((double)x); }
BPT_D float cos_f(float x)  { return (float)cos((double)x); }
BPT_D float exp_f(float x)  { return (float)exp((double)x); }
BPT_D float atan2_f(float y, float x) { return (float)atan2((double)y, (double)x); }
BPT_D float asin_f(float x) { return (float)asin((double)x); }
BPT_D float pow_f(float x, float y) { return (float)pow((double)x, (double)y); }
// sin and cos of a float, evaluated in double and rounded once, for the arguments the integrator produces (azimuths in
// [0, 2 pi]): Cody-Waite reduction by pi/2 (two FMAs, exact to ~1e-16 for |x| <= 64) and the fdlibm kernels on
// [-pi/4, pi/4].  Max error 1.7e-16 against the true value, so the float results are those of sincos((double)x) -- which
// is what runs for every other argument -- at a third of its instructions (the library call was 9 % of k_shade's).
BPT_D void sincos_f(float x, float& s, float& c) {
    if (!(fabsf(x) <= 64.0f)) { double ds, dc; sincos((double)x, &ds, &dc); s = (float)ds; c = (float)dc; return; }
    const double dx = (double)x;
    const double q = rint(dx*0.6366197723675814);
    double r = fma(-q, 1.5707963267948966, dx);
    r = fma(-q, 6.123233995736766e-17, r);
    const double z = r*r;
    double ps = 1.58969099521155010221e-10;
    ps = fma(ps, z, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    const double sn = fma(z*r, fma(z, ps, -1.66666666666666324348e-01), r);
    double pc = -1.13596475577881948265e-11;
    pc = fma(pc, z, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    const double cs = fma(z*z, pc, fma(-0.5, z, 1.0));
    const int n = __double2int_rn(q);
    const double a = (n & 1) ? cs : sn, b = (n & 1) ? sn : cs;           // quadrant: sin(r + n pi/2), cos(r + n pi/2)
    s = (float)((n & 2) ? -a : a);
    c = (float)(((n + 1) & 2) ? -b : b);
}

// n / d for a divisor known on the host: m = min(floor(2^32 / d), 2^32 - 1).  n*m / 2^32 > n/d - n / 2^32 > n/d - 1, so the
// estimate is the quotient or one below it (brute-forced in tests/test_partition_closed_form.py::test_udiv_magic_is_exact)
BPT_D uint32_t udiv_magic(uint32_t n, uint32_t d, uint32_t m, uint32_t& rem) {
    uint32_t q = __umulhi(n, m);
    uint32_t r = n - q*d;
    if (r >= d) { ++q; r -= d; }
    rem = r;
    return q;
}

// ---- RNG (samplers.h:3-108): four xorshift32 lanes ----------------------------------------------------------------
BPT_D uint32_t wang_hash(uint32_t key) {
    key += ~(key << 15);
    key ^=  (key >> 10);
    key +=  (key << 3);
    key ^=  (key >> 6);
    key += ~(key << 11);
    key ^=  (key >> 16);
    return key;
}
BPT_D uint32_t hash_coordinate3(uint32_t x, uint32_t y, uint32_t z) { return (x*73856093u) ^ (y*83492791u) ^ (z*871603259u); }
BPT_D uint32_t hash_coordinate2(uint32_t x, uint32_t y) {
    uint32_t qx = 1103515245u*((x >> 1) ^ y);
    uint32_t qy = 1103515245u*((y >> 1) ^ x);
    return 1103515245u*(qx ^ (qy >> 3));
}
BPT_D uint32_t xorshift(uint32_t v) { v ^= v << 13; v ^= v >> 17; v ^= v << 5; return v; }
// random_unilaterals (samplers.h:36-45) advances four independent xorshift32 lanes; no integrator ever reads the fourth
// (the draws use .x, .xy or .xyz), and a lane's value never enters another lane, so it is not advanced here
BPT_D void next_set(uint4& s) { s.x = xorshift(s.x); s.y = xorshift(s.y); s.z = xorshift(s.z); }
BPT_D float unilateral(uint32_t bits) { return __int_as_float((127 << 23) | (bits >> 9)) - 1.0f; }
BPT_D float bilateral(uint32_t bits) { return unilateral(bits)*2.0f - 1.0f; }

BPT_D uint4 random_seed(uint32_t seed) {
    if (seed == 0) seed = 0xFFFFFFFFu;
    uint32_t v = wang_hash(seed);
    uint32_t a = xorshift(v), b = xorshift(a), c = xorshift(b), d = xorshift(c);
    return make_uint4(wang_hash(a), wang_hash(b), wang_hash(c), d);
}

// ---- samplers (samplers.cpp:18-138) ---------------------------------------------------------------------------------
enum SampleDimension { Sample_DirectLighting, Sample_IndirectLighting, Sample_LightSelection, Sample_Reflectance,
                       Sample_DOF, Sample_AA, Sample_Roulette };

struct SamplerCtx {
    const uint8_t* strata_perm;
    const uint8_t* bn_sobol;
    const uint8_t* bn_scramble;
    const uint8_t* bn_rank;
    int      strategy;
    uint32_t index;
    uint32_t x, y;
};

// the vendored Heitz et al. lookup (blue_noise_samplers/..._256spp.cpp:17-33)
BPT_D float blue_noise_value(const SamplerCtx& c, int dim) {
    int pi = (int)c.x & 127, pj = (int)c.y & 127, si = (int)c.index & 255;
    dim &= 255;
    int ranked = si ^ (int)__ldg(&c.bn_rank[dim + (pi + pj*128)*8]);
    int value = (int)__ldg(&c.bn_sobol[dim + ranked*256]);
    value ^= (int)__ldg(&c.bn_scramble[(dim % 8) + (pi + pj*128)*8]);
    return (float)value / 256.0f;
}

BPT_D int effective_strategy(const SamplerCtx& c, int dimension) {
    int s = c.strategy;
    if (s == BPT_SAMPLING_OPTIMIZED_BLUE_NOISE && c.index > 256) s = BPT_SAMPLING_STRATIFIED;
    if (s == BPT_SAMPLING_OPTIMIZED_BLUE_NOISE && dimension >= 4) s = BPT_SAMPLING_STRATIFIED;
    return s;
}

BPT_D V2 sample_2d(const SamplerCtx& c, uint4& rng, int dimension, uint32_t bounce) {
    next_set(rng);                       // every path below draws exactly one random_unilaterals()
    float rx = unilateral(rng.x), ry = unilateral(rng.y);
    V2 s; s.x = rx; s.y = ry;
    if (bounce == 0) {
        int strat = effective_strategy(c, dimension);
        if (strat == BPT_SAMPLING_OPTIMIZED_BLUE_NOISE) {
            s.x = (1.0f / 256.0f)*rx + blue_noise_value(c, 2*dimension);
            s.y = (1.0f / 256.0f)*ry + blue_noise_value(c, 2*dimension + 1);
        } else if (strat == BPT_SAMPLING_STRATIFIED) {
            uint32_t offset = (73856093u*(uint32_t)dimension) ^ hash_coordinate2(c.x, c.y);
            uint32_t strata = __ldg(&c.strata_perm[(offset & 255)*64 + (c.index % 64)]);
            float sx = (float)(strata % 8)*(1.0f / 8.0f);
            float sy = (float)(strata / 8)*(1.0f / 8.0f);
            s.x = sx + rx*(1.0f / 8.0f);
            s.y = sy + ry*(1.0f / 8.0f);
        }
    }
    return s;
}

BPT_D float sample_1d(const SamplerCtx& c, uint4& rng, int dimension, uint32_t bounce) {
    next_set(rng);
    float rx = unilateral(rng.x);
    float s = rx;
    if (bounce == 0) {
        int strat = effective_strategy(c, dimension);
        if (strat == BPT_SAMPLING_OPTIMIZED_BLUE_NOISE) {
            s = (1.0f / 256.0f)*rx + blue_noise_value(c, 2*dimension);
        } else if (strat == BPT_SAMPLING_STRATIFIED) {
            uint32_t offset = (73856093u*(uint32_t)dimension) ^ hash_coordinate2(c.x, c.y);
            uint32_t strata = __ldg(&c.strata_perm[(offset & 255)*64 + (c.index % 64)]);
            s = (float)strata*(1.0f / 64.0f) + rx*(1.0f / 64.0f);
        }
    }
    return s;
}

} // namespace bpt
