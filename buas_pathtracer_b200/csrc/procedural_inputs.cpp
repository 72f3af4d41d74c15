// Procedural inputs of the BASELINE.json configurations: the displaced icosphere and the closed-form HDR
// environment.  Self-contained on purpose: this one source is compiled into libbpt.so (bpt_make_* of include/bpt.h) AND,
// with -DBPT_INPUTS_STANDALONE, into oracle/_ref/libbpt_inputs.so, so that the reference arm of bench.py and the oracle
// side of the tests get byte-identical inputs without loading the product library.
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <vector>

#ifdef BPT_INPUTS_STANDALONE
#define BPT_INPUTS_API extern "C" __attribute__((visibility("default")))
#define BPT_INPUTS_NAME(x) inputs_##x
static void input_error(const char*) {}
static const int kInputErrArg = -1;
#else
#include "host_scene.h"
#define BPT_INPUTS_API extern "C"
#define BPT_INPUTS_NAME(x) bpt_##x
static void input_error(const char* m) { bpt::set_error("%s", m); }
static const int kInputErrArg = BPT_ERR_ARG;
#endif

namespace {
const float kPi = 3.14159265359f;
struct F3 { float x, y, z; };
inline F3 f3(const float* v) { return {v[0], v[1], v[2]}; }
inline void st(float* d, F3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
inline float dot3(F3 a, F3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }
inline F3 normalize3(F3 a) {
    float rcp = 1.0f / sqrtf(dot3(a, a));
    return {a.x*rcp, a.y*rcp, a.z*rcp};
}
} // namespace

BPT_INPUTS_API uint32_t BPT_INPUTS_NAME(make_displaced_icosphere)(uint32_t level, float amplitude, float* positions) {
    if (level > 10) { input_error("icosphere level > 10"); return 0; }
    uint32_t count = 20;
    for (uint32_t l = 0; l < level; ++l) count *= 4;
    if (!positions) return count;

    const float t = 1.61803398875f;
    const float V[12][3] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
                            {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
    const int F[20][3] = {{0,11,5},{0,5,1},{0,1,7},{0,7,10},{0,10,11},{1,5,9},{5,11,4},{11,10,2},{10,7,6},{7,1,8},
                          {3,9,4},{3,4,2},{3,2,6},{3,6,8},{3,8,9},{4,9,5},{2,4,11},{6,2,10},{8,6,7},{9,8,1}};
    std::vector<float> cur((size_t)20*9), next;
    for (int f = 0; f < 20; ++f) for (int v = 0; v < 3; ++v) {
        F3 p = normalize3({V[F[f][v]][0], V[F[f][v]][1], V[F[f][v]][2]});
        st(&cur[(size_t)f*9 + v*3], p);
    }
    for (uint32_t l = 0; l < level; ++l) {
        size_t n = cur.size()/9;
        next.resize(n*4*9);
        for (size_t i = 0; i < n; ++i) {
            F3 a = f3(&cur[i*9]), b = f3(&cur[i*9 + 3]), c = f3(&cur[i*9 + 6]);
            // midpoints are symmetric in their endpoints, so shared edges stay watertight
            F3 ab = normalize3({(a.x + b.x)*0.5f, (a.y + b.y)*0.5f, (a.z + b.z)*0.5f});
            F3 bc = normalize3({(b.x + c.x)*0.5f, (b.y + c.y)*0.5f, (b.z + c.z)*0.5f});
            F3 ca = normalize3({(c.x + a.x)*0.5f, (c.y + a.y)*0.5f, (c.z + a.z)*0.5f});
            F3 out[4][3] = {{a, ab, ca}, {ab, b, bc}, {ca, bc, c}, {ab, bc, ca}};
            for (int k = 0; k < 4; ++k) for (int v = 0; v < 3; ++v) st(&next[(i*4 + k)*9 + v*3], out[k][v]);
        }
        cur.swap(next);
    }
    for (size_t i = 0; i < cur.size(); i += 3) {
        float x = cur[i], y = cur[i + 1], z = cur[i + 2];
        float s = 1.0f + amplitude*sinf(9.0f*x)*sinf(7.0f*y)*sinf(11.0f*z);
        positions[i] = x*s; positions[i + 1] = y*s; positions[i + 2] = z*s;
    }
    return count;
}

BPT_INPUTS_API int BPT_INPUTS_NAME(make_procedural_skydome)(uint32_t w, uint32_t h, float* pixels) {
    if (!pixels || w == 0 || h == 0) { input_error("bpt_make_procedural_skydome: bad arguments"); return kInputErrArg; }
    const F3 sun = normalize3({0.45f, 0.55f, -0.70f});
    for (uint32_t y = 0; y < h; ++y) {
        float v = ((float)y + 0.5f) / (float)h;
        float theta = (v - 0.5f)*kPi;                  // latitude, matches sample_sky's v = 0.5 + asin(d.y)/pi
        float cy = cosf(theta), sy = sinf(theta);
        for (uint32_t x = 0; x < w; ++x) {
            float u = ((float)x + 0.5f) / (float)w;
            float phi = (u - 0.5f)*2.0f*kPi;           // u = 0.5 + atan2(d.z, d.x)/2pi
            F3 d = {cy*cosf(phi), sy, cy*sinf(phi)};
            float up = d.y > 0.0f ? d.y : 0.0f;
            float down = d.y < 0.0f ? -d.y : 0.0f;
            float r = 0.55f*(1.0f - up) + 0.10f*up, g = 0.65f*(1.0f - up) + 0.25f*up, b = 0.80f*(1.0f - up) + 0.90f*up;
            float gr = 1.0f - 0.75f*down;              // darker "ground" hemisphere
            r *= gr; g *= gr*0.95f; b *= gr*0.85f;
            float c = dot3(d, sun);
            float halo = expf(-(1.0f - c)*60.0f)*4.0f;
            float disc = c > 0.9995f ? 400.0f : 0.0f;  // HDR sun
            float* px = &pixels[((size_t)y*w + x)*3];
            px[0] = r + (halo + disc)*1.00f;
            px[1] = g + (halo + disc)*0.92f;
            px[2] = b + (halo + disc)*0.80f;
        }
    }
    return 0;
}

