"""GPU parity tests proper (run with -m gpu on the B200): the CUDA path, called through the C ABI, against the
reference's own CPU code on the same inputs.

Bars (BASELINE.json north_star):  hit primitive/triangle ids, t, hit point and normal: bit-exact.
Per-sample radiance: FP32 tolerance (see RADIANCE_* below).  Film: relative RMSE.
"""
import numpy as np
import pytest

from buas_pathtracer_b200 import capi, scenes
from helpers import (build_both, camera_rays, secondary_rays, assert_hits_equal, bits, rel_rmse)

pytestmark = pytest.mark.gpu

# Tolerances, stated once.  Integer work (RNG, sampler tables, hit ids) is bit-exact.  Float work that only uses
# + - * / sqrt is bit-exact too (no FMA contraction on either side).  The remaining difference is libm: the device
# evaluates sin/cos/exp/atan2/asin in double and rounds once, glibc's float routines are within ~0.56 ulp, so a
# small fraction of samples sees a 1-ulp change in a direction; through the path that stays a relative error of
# a few 1e-6 unless it flips a discrete decision (Russian roulette, Fresnel pick, a different triangle), which the
# outlier budget below covers.
RADIANCE_REL_TOL = 2e-4        # max relative error per sample, |gpu-ref| <= tol*max(|ref|, 1e-3*mean)
RADIANCE_OUTLIER_FRAC = 2e-3   # samples allowed to diverge discretely
FILM_REL_RMSE = 2e-3           # low-spp film (includes those outliers)


def _trace_both(renderer, ref_scene, rays, mode, ignored=0):
    g = renderer.trace(rays, mode, ignored)
    r = ref_scene.trace(rays, mode, ignored)
    return g, r


@pytest.fixture(scope="module")
def c1(bpt, oracle, renderer):
    a, b = build_both(bpt, oracle, scenes.c1_week3, 640, 360)
    return a, b


@pytest.fixture(scope="module")
def ico6(bpt, oracle):
    a, b = build_both(bpt, oracle, scenes.c2_icosphere, 320, 180, level=6)
    return a, b


@pytest.fixture(scope="module")
def inst(bpt, oracle):
    a, b = build_both(bpt, oracle, scenes.c3_instances, 320, 180, level=4, grid=4, sky_size=(256, 128))
    return a, b


@pytest.fixture(scope="module")
def nested(bpt, oracle):
    a, b = build_both(bpt, oracle, scenes.c4_nested_dielectrics, 320, 180)
    return a, b


def _hit_parity(renderer, pair, w, h, n=60000, light_pos=None, light_id=None):
    a, b = pair
    renderer.upload_scene(a)
    cam = a.get_camera()
    rays = camera_rays(cam, w, h, n)
    g, r = _trace_both(renderer, b, rays, capi.TRACE_CLOSEST)
    assert_hits_equal(g, r, capi.TRACE_CLOSEST, "primary")
    assert np.count_nonzero(g["primitive"] != capi.HIT_MISS) > n // 4
    # incoherent secondary rays from agreed hit points
    sec = secondary_rays(g, rays, n, seed=5)
    g2, r2 = _trace_both(renderer, b, sec, capi.TRACE_CLOSEST)
    assert_hits_equal(g2, r2, capi.TRACE_CLOSEST, "secondary")
    # a third generation, even less coherent
    ter = secondary_rays(g2, sec, n, seed=6)
    g3, r3 = _trace_both(renderer, b, ter, capi.TRACE_CLOSEST)
    assert_hits_equal(g3, r3, capi.TRACE_CLOSEST, "tertiary")
    if light_pos is not None:
        sh = secondary_rays(g, rays, n, seed=7, toward=light_pos)
        g4, r4 = _trace_both(renderer, b, sh, capi.TRACE_OCCLUSION, light_id)
        assert_hits_equal(g4, r4, capi.TRACE_OCCLUSION, "shadow")
        occluded = np.count_nonzero(g4["primitive"] != capi.HIT_MISS)
        assert 0 < occluded < n, "shadow batch should contain both occluded and free rays"
    return g, g2


def test_hit_ids_bit_exact_spheres_planes(renderer, c1):
    a, _ = c1
    _hit_parity(renderer, c1, 640, 360, light_pos=(8, 16, -8), light_id=a.counts()["primitives"] - 1)


def test_hit_ids_bit_exact_mesh(renderer, ico6):
    a, _ = ico6
    g, g2 = _hit_parity(renderer, ico6, 320, 180, light_pos=(9, 14, -9), light_id=a.counts()["primitives"] - 1)
    assert np.count_nonzero(g["triangle"] != 0xFFFFFFFF) > 1000


def test_hit_ids_bit_exact_instances(renderer, inst):
    a, _ = inst
    _hit_parity(renderer, inst, 320, 180, light_pos=(-14, 22, -10), light_id=a.counts()["primitives"] - 1)


def test_hit_ids_bit_exact_nested_boxes_spheres(renderer, nested):
    a, _ = nested
    _hit_parity(renderer, nested, 320, 180, light_pos=(0, 26, 12), light_id=a.counts()["primitives"] - 1)


def test_hit_ids_axis_aligned_and_degenerate_rays(renderer, ico6, inst):
    """rays with exactly-zero direction components (inv_d = inf, 0*inf = NaN inside the slab test, SURVEY Appendix A
    #2) take the exact ternary min/max path; denormal-small components too; results still match bit for bit"""
    rng = np.random.RandomState(21)
    n = 3000
    for pair in (ico6, inst):
        a, b = pair
        renderer.upload_scene(a)
        rays = np.zeros(n, capi.RAY_DTYPE)
        o = (rng.rand(n, 3).astype(np.float32) - 0.5) * np.float32(12) + np.array([0, 9, 0], np.float32)
        d = np.zeros((n, 3), np.float32)
        kind = rng.randint(0, 4, n)
        d[kind == 0] = (0, -1, 0)                                  # straight down
        d[kind == 1] = np.array((0.6, -0.8, 0), np.float32)        # one zero component
        hz = (kind == 1) & (rng.rand(n) < 0.5)
        d[hz] = np.array((-0.24058728, 0.0, 0.9706275), np.float32)  # horizon ray: the reference pops ~6 M nodes for one of these
        d[kind == 2] = np.array((1e-39, -1, 1e-30), np.float32)    # denormal / tiny components
        m3 = kind == 3
        d[m3] = rng.randn(int(m3.sum()), 3); d[m3] /= np.linalg.norm(d[m3], axis=1, keepdims=True)
        # snap some origins onto node-box planes so that (o - p) == 0 happens with inv = inf
        nodes, _, _ = a.mesh_bvh(0)
        pick = nodes[rng.randint(0, len(nodes), n)]
        snap = rng.rand(n) < 0.5
        o[snap, 0] = (pick["bv_p"][snap, 0] * np.float32(3.5)).astype(np.float32)
        o[hz, 1] = rng.rand(int(hz.sum())).astype(np.float32) * np.float32(3.0)      # horizon rays through the meshes
        rays["o"] = o; rays["d"] = d; rays["max_t"] = np.finfo(np.float32).max
        g = renderer.trace(rays, capi.TRACE_CLOSEST)
        r = b.trace(rays, capi.TRACE_CLOSEST)
        assert_hits_equal(g, r, capi.TRACE_CLOSEST, "axis-aligned")
        assert np.count_nonzero(g["primitive"] != capi.HIT_MISS) > n // 10


def test_traversal_counters_equal_g_stats(renderer, ico6, oracle):
    """TraversalStats (intersection.h:33-40): same pops / inner / leaf counts as the reference's binary traversal."""
    a, b = ico6
    renderer.upload_scene(a)
    rays = camera_rays(a.get_camera(), 320, 180, 20000, seed=11)
    renderer.stats_enable(True)
    renderer.get_stats(reset=True)
    oracle.get_stats(reset=True)
    g = renderer.trace(rays, capi.TRACE_CLOSEST)
    gs = renderer.get_stats(reset=True).as_dict()
    # the oracle's trace wrapper re-runs intersect_mesh for the triangle id but restores g_stats around it
    b.trace(rays, capi.TRACE_CLOSEST)
    rs = oracle.get_stats(reset=True).as_dict()
    renderer.stats_enable(False)
    for k in ("mesh_intersection_count", "mesh_bvh_traversals", "mesh_node_traversals", "mesh_leaf_traversals"):
        assert gs[k] == rs[k], (k, gs[k], rs[k])
    assert gs["rays"] == 20000 and gs["triangles_tested"] > 0


def _render_both(renderer, pair, w, h, spp, salt=0, frame_count=0):
    a, b = pair
    renderer.upload_scene(a)
    renderer.film_resize(w, h)
    # two schedules of the same work: pure wavefront (one round of launches per bounce) and the default, which hands
    # the last survivors of a batch to the fused k_tail launch.  Per path they execute the same operations in the
    # same order, so the per-sample records must agree bit for bit; the reference is then compared with the default.
    renderer.set_tail_threshold(0)
    rec0 = renderer.attach_records(w * h * spp)
    renderer.render_pass(spp, frame_count=frame_count, salt=salt)
    renderer.sync()
    rec0 = rec0.copy()
    renderer.attach_records(0)
    renderer.set_tail_threshold(65536)
    renderer.film_clear()
    rec = renderer.attach_records(w * h * spp)
    renderer.render_pass(spp, frame_count=frame_count, salt=salt)
    film = renderer.download_film()
    renderer.attach_records(0)
    rec = rec.copy()
    for k in ("radiance", "ray_o", "ray_d", "rays"):
        assert np.array_equal(bits(rec0[k]) if rec0[k].dtype == np.float32 else rec0[k],
                              bits(rec[k]) if rec[k].dtype == np.float32 else rec[k]), f"wavefront vs fused tail: {k} differs"
    rfilm, rrec = b.render_parity(w, h, spp, frame_count=frame_count, salt=salt, records=True)
    return film, rec, rfilm, rrec


def _check_records(rec, rrec, what):
    n = rec.shape[0]
    # ray generation (sampler + camera) has no libm on this path apart from the unused bokeh cos/sin: bit-exact
    assert np.array_equal(bits(rec["ray_o"]), bits(rrec["ray_o"])), f"{what}: primary ray origins differ"
    nbad = np.count_nonzero(np.any(bits(rec["ray_d"]) != bits(rrec["ray_d"]), axis=1))
    assert nbad == 0, f"{what}: {nbad}/{n} primary ray directions differ in their bits"
    g, r = rec["radiance"].astype(np.float64), rrec["radiance"].astype(np.float64)
    scale = max(float(np.mean(np.abs(r))), 1e-12)
    err = np.max(np.abs(g - r) / np.maximum(np.abs(r), 1e-3 * scale), axis=1)
    exact = np.count_nonzero(np.all(bits(rec["radiance"]) == bits(rrec["radiance"]), axis=1))
    outliers = np.count_nonzero(err > RADIANCE_REL_TOL)
    same_rays = np.count_nonzero(rec["rays"] == rrec["rays"])
    print(f"{what}: {n} samples, bit-exact radiance {exact/n:.4%}, same ray count {same_rays/n:.4%}, "
          f"outliers(>{RADIANCE_REL_TOL:g}) {outliers} ({outliers/n:.4%}), "
          f"max rel err among inliers {float(err[err <= RADIANCE_REL_TOL].max()):.3g}")
    assert outliers <= max(2, RADIANCE_OUTLIER_FRAC * n), f"{what}: {outliers}/{n} samples beyond {RADIANCE_REL_TOL}"
    assert exact >= 0.90 * n, f"{what}: only {exact}/{n} samples bit-exact"


def _check_film(film, rfilm, what):
    # weights use no libm: sum-of-weights agrees to float-add reordering only
    wg, wr = film[..., 3].astype(np.float64), rfilm[..., 3].astype(np.float64)
    assert np.allclose(wg, wr, rtol=2e-5, atol=1e-6), f"{what}: filter weight sums differ"
    e = rel_rmse(film[..., :3].astype(np.float64), rfilm[..., :3].astype(np.float64))
    print(f"{what}: film relRMSE {e:.3g}")
    assert e <= FILM_REL_RMSE, f"{what}: film relRMSE {e}"


def test_render_parity_c1_small(renderer, c1):
    film, rec, rfilm, rrec = _render_both(renderer, c1, 640, 360, 1)
    _check_records(rec, rrec, "C1 640x360x1")
    _check_film(film, rfilm, "C1 640x360x1")


def test_render_parity_c1_multi_spp_and_progressive(renderer, c1, bpt, oracle):
    a, b = build_both(bpt, oracle, scenes.c1_week3, 160, 90)
    film, rec, rfilm, rrec = _render_both(renderer, (a, b), 160, 90, 8, salt=0x9e3779b9)
    _check_records(rec, rrec, "C1 160x90x8")
    _check_film(film, rfilm, "C1 160x90x8")
    # second progressive pass accumulates on top (frame_count advances by spp, raytracer.cpp:721-722)
    renderer.render_pass(8, frame_count=8, salt=0x9e3779b9)
    film2 = renderer.download_film()
    rfilm2, _ = b.render_parity(160, 90, 8, frame_count=8, salt=0x9e3779b9, film=rfilm.copy())
    _check_film(film2, rfilm2, "C1 160x90 pass 2")


def test_render_parity_mesh(renderer, bpt, oracle):
    pair = build_both(bpt, oracle, scenes.c2_icosphere, 160, 90, level=5)
    film, rec, rfilm, rrec = _render_both(renderer, pair, 160, 90, 4)
    _check_records(rec, rrec, "icosphere L5 160x90x4")
    _check_film(film, rfilm, "icosphere L5 160x90x4")


def test_render_parity_instances_envmap(renderer, bpt, oracle):
    pair = build_both(bpt, oracle, scenes.c3_instances, 160, 90, level=3, grid=4, sky_size=(256, 128))
    film, rec, rfilm, rrec = _render_both(renderer, pair, 160, 90, 4)
    _check_records(rec, rrec, "instances 160x90x4")
    _check_film(film, rfilm, "instances 160x90x4")


def test_render_parity_nested_dielectrics(renderer, bpt, oracle):
    pair = build_both(bpt, oracle, scenes.c4_nested_dielectrics, 160, 90)
    film, rec, rfilm, rrec = _render_both(renderer, pair, 160, 90, 4)
    _check_records(rec, rrec, "nested dielectrics 160x90x4")
    _check_film(film, rfilm, "nested dielectrics 160x90x4")


@pytest.mark.parametrize("name", ["Normals", "Distances", "Ground Truth Iterative", "Ground Truth Recursive", "Whitted"])
def test_render_parity_other_integrators(renderer, bpt, oracle, name):
    """g_integrators[] entries beyond the default (integrators.cpp:310-580), selected by the reference's own names."""
    for recipe, kw, what in ((scenes.c1_week3, {}, "C1"), (scenes.c2_icosphere, dict(level=4), "icosphere L4"),
                             (scenes.c4_nested_dielectrics, {}, "nested dielectrics"),
                             (scenes.whitted_showcase, {}, "whitted showcase")):
        a, b = build_both(bpt, oracle, recipe, 128, 72, **kw)
        for s in (a, b):
            s.update_settings(integrator=name)
            if name == "Whitted" and recipe is scenes.c4_nested_dielectrics:
                s.update_settings(max_bounce_count=8)       # 2^depth rays per sample through the glass: keep the CPU side short
        film, rec, rfilm, rrec = _render_both(renderer, (a, b), 128, 72, 3)
        _check_records(rec, rrec, f"{what} / {name}")
        _check_film(film, rfilm, f"{what} / {name}")


@pytest.mark.parametrize("strategy", [capi.SAMPLING_UNIFORM, capi.SAMPLING_BLUE_NOISE])
def test_render_parity_other_samplers(renderer, bpt, oracle, strategy):
    a, b = build_both(bpt, oracle, scenes.c1_week3, 160, 90)
    for s in (a, b):
        s.update_settings(sampling_strategy=strategy)
    film, rec, rfilm, rrec = _render_both(renderer, (a, b), 160, 90, 4)
    _check_records(rec, rrec, f"C1 sampler {strategy}")
    _check_film(film, rfilm, f"C1 sampler {strategy}")


def test_render_parity_settings_variants(renderer, bpt, oracle):
    """NEE off / uniform light pick / uniform hemisphere / no MIS / no RR / box filter / lens distortion + DOF"""
    variants = [
        dict(next_event_estimation=0),
        dict(importance_sample_lights=0, importance_sample_diffuse=0),
        dict(use_mis=0, russian_roulette=0, caustics=0),
        dict(lens_distortion=1.0, vignette_strength=0.6),
    ]
    for kw in variants:
        a, b = build_both(bpt, oracle, scenes.c1_week3, 128, 72)
        for s in (a, b):
            s.update_settings(**kw)
        film, rec, rfilm, rrec = _render_both(renderer, (a, b), 128, 72, 2)
        _check_records(rec, rrec, f"C1 {kw}")
        _check_film(film, rfilm, f"C1 {kw}")
    a, b = build_both(bpt, oracle, scenes.c1_week3, 128, 72)
    for s in (a, b):
        s.load_reconstruction_kernel("Box")
    film, rec, rfilm, rrec = _render_both(renderer, (a, b), 128, 72, 2)
    _check_film(film, rfilm, "C1 box filter")
    a, b = build_both(bpt, oracle, scenes.c1_week3, 128, 72)
    for s in (a, b):
        s.load_reconstruction_kernel("Gaussian 3")
    film, rec, rfilm, rrec = _render_both(renderer, (a, b), 128, 72, 2)
    _check_film(film, rfilm, "C1 gaussian-3 filter")


@pytest.mark.parametrize("name", ["Gaussian 12", "Lanczos 3", "Lanczos 4", "Lanczos 6", "Lanczos 12"])
def test_wide_reconstruction_filters_on_device(renderer, bpt, oracle, name):
    """g_filters[] beyond the default (reconstruction_filters.cpp:97-106): splats of radius 3, 4, 6 and 12 pixels -- up to 625 film
    updates per sample, negative lobes included (Lanczos) -- through k_splat_generic against the reference's splat_filter.  The
    radius-12 footprints cross the whole 48-row band that is rendered here, and its borders."""
    w, h = 96, 64
    a, b = build_both(bpt, oracle, scenes.c1_week3, w, h)
    for s in (a, b):
        s.load_reconstruction_kernel(name)
    rect, spp = (8, 8, 88, 56), 2
    renderer.upload_scene(a)
    renderer.film_resize(w, h)
    renderer.render_pass(spp, rect=rect)
    film = renderer.download_film()
    rfilm, _ = b.render_parity(w, h, spp, rect=rect, records=True)
    # weights of a Lanczos kernel change sign: compare them absolutely, scaled by the largest weight sum of the film
    wg, wr = film[..., 3].astype(np.float64), rfilm[..., 3].astype(np.float64)
    assert np.allclose(wg, wr, rtol=0, atol=2e-5 * float(np.abs(wr).max())), f"{name}: filter weight sums differ"
    assert np.count_nonzero(wr[:8]) > 0 and np.count_nonzero(wr[:, :8]) > 0, "the footprint must reach outside the rendered rect"
    e = rel_rmse(film[..., :3].astype(np.float64), rfilm[..., :3].astype(np.float64))
    print(f"C1 {name}: film relRMSE {e:.3g}")
    assert e <= FILM_REL_RMSE, f"{name}: film relRMSE {e}"


def test_subrect_and_row_sharding_is_partition_independent(renderer, c1, bpt, oracle):
    """rendering the image as row bands (the multi-GPU partition) sums to the full-frame film"""
    a, b = build_both(bpt, oracle, scenes.c1_week3, 160, 90)
    renderer.upload_scene(a)
    renderer.film_resize(160, 90)
    renderer.render_pass(4)
    full = renderer.download_film()
    renderer.film_clear()
    for y0, y1 in ((0, 23), (23, 45), (45, 90)):
        renderer.render_pass(4, rect=(0, y0, 160, y1))
    banded = renderer.download_film()
    assert np.allclose(full, banded, rtol=1e-5, atol=1e-6)


def test_errors_are_reported_not_fatal(renderer, bpt):
    r2 = bpt.Renderer(0)
    with pytest.raises(bpt.BptError):
        r2.render_pass(1)                    # no scene
    s = bpt.Scene()
    with pytest.raises(bpt.BptError):
        r2.upload_scene(s)                   # no BVH yet
    scenes.c1_week3(s, 64, 36)
    r2.upload_scene(s)
    with pytest.raises(bpt.BptError):
        r2.render_pass(1)                    # no film
    r2.film_resize(64, 36)
    with pytest.raises(bpt.BptError):
        r2.render_pass(1, rect=(0, 0, 65, 36))
    s.update_settings(integrator="Whitted", max_bounce_count=40)
    r2.update_settings(s)
    with pytest.raises(bpt.BptError):
        r2.render_pass(1)                    # recursion deeper than the library's frame stack is refused
    r2.close()
