// ORACLE-ONLY (test infrastructure).  Wrapper TU around the reference's integrators.cpp (unmodified):
// exposes a few of its static helpers as known-answer functions for device-math unit parity.
#include "integrators.cpp"
#include "ref_api.h"

extern "C" BPT_API void
ref_kat_cosine_hemisphere(const float n[3], const float u[2], float out[3]) {
    V3 r = map_to_cosine_weighted_hemisphere(v3(n[0], n[1], n[2]), v2(u[0], u[1]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
extern "C" BPT_API void
ref_kat_hemisphere(const float n[3], const float u[2], float out[3]) {
    V3 r = map_to_hemisphere(v3(n[0], n[1], n[2]), v2(u[0], u[1]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
extern "C" BPT_API float
ref_kat_fresnel(float cos_i, float eta_i, float eta_t, float* cos_t) {
    return fresnel_dielectric(cos_i, eta_i, eta_t, eta_i / eta_t, cos_t);
}
