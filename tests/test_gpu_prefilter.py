"""GPU tests (-m gpu) of the ray prefilter: k_shade settles the NEE shadow rays that never reach a mesh BLAS itself
(kernels.cuh shadow_tlas_head = the head of intersect_shadow_ray, intersection.cpp:424-520) instead of queueing them for the
traversal kernels.  That moves work between kernels and must not move a single bit of a result: per-sample radiance and
ray counts with the prefilter on equal those with it off, bit for bit, on every scene type (parity with the reference
itself is asserted by the other GPU tests, which run with the prefilter on -- the default)."""
import numpy as np
import pytest

from buas_pathtracer_b200 import scenes
from helpers import bits

pytestmark = pytest.mark.gpu

CASES = [
    ("c1 spheres + plane (one-leaf TLAS, no mesh: every shadow ray is settled)", scenes.c1_week3, {}, 160, 90, 4),
    ("c2 icosphere (one-leaf TLAS: mesh + light)", scenes.c2_icosphere, dict(level=4), 160, 90, 4),
    ("c3 instances (TLAS with inner nodes: only plane hits / root misses are settled)", scenes.c3_instances,
     dict(level=3, grid=4, sky_size=(64, 32)), 160, 90, 4),
    ("c4 nested dielectrics", scenes.c4_nested_dielectrics, {}, 128, 72, 4),
    ("kitchen sink (three lights, rough metal, boxes, rotated mesh with vertex normals)", scenes.kitchen_sink, {}, 128, 72, 4),
]


def _records(bpt, scene, w, h, spp, prefilter, tail=True):
    r = bpt.Renderer(0)
    r.set_ray_prefilter(prefilter)
    r.upload_scene(scene)
    r.film_resize(w, h)
    rec = r.attach_records(w * h * spp)
    r.get_stats(reset=True)
    r.render_pass(spp)
    r.sync()
    rec = rec.copy()
    st = r.get_stats(reset=True)
    r.attach_records(0)
    # the same pass without records: two pipelines, merged traversal launches, the fused tail -- film only
    r.film_clear()
    r.render_pass(spp)
    film = r.download_film()
    r.close()
    return rec, st, film


@pytest.mark.parametrize("what,recipe,kw,w,h,spp", CASES, ids=[c[0].split(" ")[0] + "_" + c[0].split(" ")[1] for c in CASES])
def test_prefilter_changes_no_result(bpt, what, recipe, kw, w, h, spp):
    s = bpt.Scene()
    recipe(s, w, h, **kw)
    on, st_on, film_on = _records(bpt, s, w, h, spp, True)
    off, st_off, film_off = _records(bpt, s, w, h, spp, False)
    assert np.array_equal(bits(on["radiance"]), bits(off["radiance"])), what
    assert np.array_equal(on["rays"], off["rays"]), what
    assert st_on.rays == st_off.rays and st_on.shadow_rays == st_off.shadow_rays, what
    assert np.allclose(film_on, film_off, rtol=1e-5, atol=1e-6), what


def test_counting_pass_reports_what_a_normal_pass_settles(bpt):
    """With stats enabled every shadow ray goes through the traversal kernels (bpt_stats stays in the reference's units) and
    the rays a normal pass would settle in k_shade are only counted: all of them on C1 (no mesh), a part of them on C2 (the
    shadow rays of ground-plane points miss the mesh's root box, those of points on the mesh do not)."""
    w, h, spp = 160, 90, 4
    got = {}
    for name, recipe, kw in (("c1", scenes.c1_week3, {}), ("c2", scenes.c2_icosphere, dict(level=4))):
        s = bpt.Scene()
        recipe(s, w, h, **kw)
        r = bpt.Renderer(0)
        r.upload_scene(s)
        r.film_resize(w, h)
        r.stats_enable(True)
        r.get_stats(reset=True)
        r.render_pass(spp)
        r.sync()
        settled, nbytes = r.ray_prefilter_stats()
        st = r.get_stats(reset=True)
        got[name] = (settled, nbytes, st.shadow_rays, st.shadow_tlas_node_pops)
        r.close()
    settled, nbytes, shadow, pops = got["c1"]
    assert shadow > 0 and settled == shadow, got["c1"]            # spheres and a plane only: nothing needs a BLAS
    assert pops > 0 and nbytes > 0                                 # and the counting pass still traced them all in the traversal kernel
    settled, nbytes, shadow, pops = got["c2"]
    assert 0 < settled < shadow, got["c2"]                        # ground-plane points: settled; points on the mesh: traced
