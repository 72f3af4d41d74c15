// ORACLE-ONLY (test infrastructure).  Wrapper TU around the reference's samplers.cpp (unmodified) so the
// file-static blue-noise tables can be exported as bytes and the samplers called for known-answer tests.
#include "samplers.cpp"
#include "ref_api.h"

extern "C" BPT_API int
ref_get_sampler_tables(uint8_t* perm, uint8_t* sobol, uint8_t* scramble, uint8_t* rank) {
    if (perm)     memcpy(perm, g_strata_permutation_sets, 256*64);
    if (sobol)    for (int i = 0; i < 256*256;   ++i) sobol[i]    = (uint8_t)sobol_256spp_256d[i];
    if (scramble) for (int i = 0; i < 128*128*8; ++i) scramble[i] = (uint8_t)scramblingTile[i];
    if (rank)     for (int i = 0; i < 128*128*8; ++i) rank[i]     = (uint8_t)rankingTile[i];
    int ok = 1;
    for (int i = 0; i < 256*256;   ++i) ok &= (sobol_256spp_256d[i] >= 0 && sobol_256spp_256d[i] < 256);
    for (int i = 0; i < 128*128*8; ++i) ok &= (scramblingTile[i] >= 0 && scramblingTile[i] < 256 && rankingTile[i] >= 0 && rankingTile[i] < 256);
    return ok ? 0 : -1;
}

extern "C" BPT_API void
ref_kat_random_seed(uint32_t seed, uint32_t out_state[4]) {
    RandomSeries s = random_seed(seed);
    memcpy(out_state, s.e, 16);
}

extern "C" BPT_API void
ref_kat_sample_2d(uint32_t state[4], int strategy, uint32_t index, uint32_t x, uint32_t y, int dim, uint32_t bounce, float out[2]) {
    RandomSeries series; memcpy(series.e, state, 16);
    Sampler sampler = {};
    sampler.entropy = &series; sampler.strategy = (SamplingStrategy)strategy;
    sampler.sample_index = index; sampler.x = x; sampler.y = y;
    V2 r = get_next_sample_2d(&sampler, (SampleDimension)dim, bounce);
    out[0] = r.x; out[1] = r.y;
    memcpy(state, series.e, 16);
}

extern "C" BPT_API float
ref_kat_sample_1d(uint32_t state[4], int strategy, uint32_t index, uint32_t x, uint32_t y, int dim, uint32_t bounce) {
    RandomSeries series; memcpy(series.e, state, 16);
    Sampler sampler = {};
    sampler.entropy = &series; sampler.strategy = (SamplingStrategy)strategy;
    sampler.sample_index = index; sampler.x = x; sampler.y = y;
    float r = get_next_sample_1d(&sampler, (SampleDimension)dim, bounce);
    memcpy(state, series.e, 16);
    return r;
}
