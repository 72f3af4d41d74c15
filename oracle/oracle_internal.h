// ORACLE-ONLY (test infrastructure). Shared between the wrapper TUs; include AFTER the reference headers.
#pragma once
#include <vector>
#include "ref_api.h"

struct ref_scene {
    Scene scene;                 // the reference's own Scene (Raytracer/scene.h:91-120)
    Arena temp;                  // "transient_arena" of SDL_main (raytracer.cpp:1606)
    std::vector<Mesh> meshes;    // Mesh values handed to add_mesh (instances share triangles + bvh)
    EnvironmentMap skydome;
    bpt_filter_cache filter;     // copy of g_filter_cache as loaded for this scene
    bool latched;
};

static inline M4x4Inv to_ref(const bpt_m4x4inv* m) { M4x4Inv r; memcpy(&r, m, sizeof(r)); return r; }
static inline V3 to_v3(const float* f) { return v3(f[0], f[1], f[2]); }

// counters fed by the --wrap'd intersect_scene / intersect_shadow_ray (see tu_intersection.cpp)
extern volatile int  g_ref_count_rays;
extern volatile u64  g_ref_rays;
extern volatile u64  g_ref_shadow_rays;
