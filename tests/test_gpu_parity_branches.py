"""GPU parity tests (-m gpu) for the branches of the default integrator that the BASELINE configs never reach, for the
BASELINE configs at their benchmarked geometry, and for the converged-image criterion of BASELINE.json's north_star.

Tolerances are stated here once:
  * hit ids / t / p / n of ray batches: bit-exact;
  * per-sample radiance: max relative error 2e-4 (as in test_gpu_parity.py), with an outlier budget for samples whose
    path took a different discrete decision after a 1-ulp libm difference;
  * with depth of field the PRIMARY ray itself passes through cos/sin/pow (polygonal bokeh, raytracer.cpp:86-94): the
    device evaluates them in double and rounds once, glibc's float routines are within 1 ulp, so primary rays agree to
    1e-6 absolute with >= 97 % of them bit-exact (without DOF they are bit-exact, asserted elsewhere);
  * converged image (>= 1024 spp): relative RMSE of the resolved image <= 1e-3.
"""
import numpy as np
import pytest

from buas_pathtracer_b200 import capi, scenes
from helpers import build_both, camera_rays, secondary_rays, assert_hits_equal, bits, rel_rmse

pytestmark = pytest.mark.gpu

RADIANCE_REL_TOL = 2e-4
FILM_REL_RMSE = 2e-3
CONVERGED_REL_RMSE = 1e-3


def _render(renderer, pair, w, h, spp, rect=None, salt=0, frame_count=0):
    a, b = pair
    renderer.upload_scene(a)
    renderer.film_resize(w, h)
    x0, y0, x1, y1 = rect if rect else (0, 0, w, h)
    n = (x1 - x0) * (y1 - y0) * spp
    rec = renderer.attach_records(n)
    renderer.render_pass(spp, rect=rect, salt=salt, frame_count=frame_count)
    film = renderer.download_film()
    rec = rec.copy()
    renderer.attach_records(0)
    rfilm, rrec = b.render_parity(w, h, spp, rect=rect, salt=salt, frame_count=frame_count, records=True)
    return film, rec, rfilm, rrec


def _radiance_report(rec, rrec, what, outlier_frac, exact_frac, select=None):
    g, r = rec["radiance"].astype(np.float64), rrec["radiance"].astype(np.float64)
    gb, rb = bits(rec["radiance"]), bits(rrec["radiance"])
    if select is not None:
        g, r, gb, rb = g[select], r[select], gb[select], rb[select]
    n = g.shape[0]
    scale = max(float(np.mean(np.abs(r))), 1e-12)
    err = np.max(np.abs(g - r) / np.maximum(np.abs(r), 1e-3 * scale), axis=1)
    exact = int(np.count_nonzero(np.all(gb == rb, axis=1)))
    outliers = int(np.count_nonzero(err > RADIANCE_REL_TOL))
    print(f"{what}: {n} samples, bit-exact radiance {exact/n:.4%}, outliers(>{RADIANCE_REL_TOL:g}) {outliers} ({outliers/n:.4%})")
    assert outliers <= max(2, outlier_frac * n), f"{what}: {outliers}/{n} samples beyond {RADIANCE_REL_TOL}"
    assert exact >= exact_frac * n, f"{what}: only {exact}/{n} samples bit-exact"


def _film_check(film, rfilm, what, rows=None):
    if rows is not None:
        film, rfilm = film[rows], rfilm[rows]
    wg, wr = film[..., 3].astype(np.float64), rfilm[..., 3].astype(np.float64)
    assert np.allclose(wg, wr, rtol=2e-5, atol=1e-6), f"{what}: filter weight sums differ"
    e = rel_rmse(film[..., :3].astype(np.float64), rfilm[..., :3].astype(np.float64))
    print(f"{what}: film relRMSE {e:.3g}")
    assert e <= FILM_REL_RMSE, f"{what}: film relRMSE {e}"


# ---- the kitchen sink: DOF + polygonal bokeh, rough metal / rough dielectric, three lights, vertex normals ----------------
@pytest.mark.parametrize("isl", [1, 0])
def test_kitchen_sink_render_parity(renderer, bpt, oracle, isl):
    w, h, spp = 160, 90, 6
    pair = build_both(bpt, oracle, scenes.kitchen_sink, w, h, importance_sample_lights=isl)
    a, b = pair
    assert a.get_camera().lens_radius > 0 and a.get_settings().f_factor > 0 and a.counts()["lights"] == 3
    film, rec, rfilm, rrec = _render(renderer, pair, w, h, spp)
    n = rec.shape[0]
    # primary rays go through the bokeh transform's cos/sin/pow: equal to 1e-6, nearly all of them bit-exact
    for f in ("ray_o", "ray_d"):
        assert np.allclose(rec[f], rrec[f], rtol=0, atol=1e-6), f"primary {f} differs beyond libm rounding"
    same_primary = np.all(bits(rec["ray_o"]) == bits(rrec["ray_o"]), axis=1) & np.all(bits(rec["ray_d"]) == bits(rrec["ray_d"]), axis=1)
    print(f"kitchen sink isl={isl}: primary rays bit-exact {np.count_nonzero(same_primary)/n:.4%}")
    assert np.count_nonzero(same_primary) >= 0.97 * n
    assert np.ptp(rec["ray_o"], axis=0).max() > 1e-3, "the lens must actually be sampled (lens_radius > 0)"
    # samples whose primary ray agrees bit for bit follow the usual radiance bar
    _radiance_report(rec, rrec, f"kitchen sink isl={isl} (same primary ray)", 3e-3, 0.90, select=same_primary)
    # all samples: a 1-ulp different primary ray is a different (equally valid) sample; the film bar covers them
    _film_check(film, rfilm, f"kitchen sink isl={isl}")
    same_rays = np.count_nonzero(rec["rays"] == rrec["rays"])
    assert same_rays >= 0.99 * n
    assert rec["rays"].max() >= 8, "paths through the rough dielectric / metal should run several bounces"


def test_kitchen_sink_hit_records_with_vertex_normals(renderer, bpt, oracle):
    """interpolated per-vertex normals through a rotated + scaled instance (intersection.cpp:560-591 and the reference's
    own transform_normal, my_math.h:956-963): normal bits equal on primary, secondary and shadow batches"""
    w, h = 160, 90
    a, b = build_both(bpt, oracle, scenes.kitchen_sink, w, h)
    renderer.upload_scene(a)
    cam = a.get_camera()
    n = 40000
    rays = camera_rays(cam, w, h, n, seed=3)
    g = renderer.trace(rays, capi.TRACE_CLOSEST)
    r = b.trace(rays, capi.TRACE_CLOSEST)
    assert_hits_equal(g, r, capi.TRACE_CLOSEST, "kitchen sink primary")
    mesh_hits = np.count_nonzero(g["triangle"] != 0xFFFFFFFF)
    assert mesh_hits > 200, "the mesh with vertex normals must be in view"
    sec = secondary_rays(g, rays, n, seed=8)
    g2, r2 = renderer.trace(sec, capi.TRACE_CLOSEST), b.trace(sec, capi.TRACE_CLOSEST)
    assert_hits_equal(g2, r2, capi.TRACE_CLOSEST, "kitchen sink secondary")
    lights = a.counts()["primitives"] - 3
    sh = secondary_rays(g, rays, n, seed=9, toward=(6.0, 13.0, -6.0))
    g3, r3 = renderer.trace(sh, capi.TRACE_OCCLUSION, lights), b.trace(sh, capi.TRACE_OCCLUSION, lights)
    assert_hits_equal(g3, r3, capi.TRACE_OCCLUSION, "kitchen sink shadow")


def test_blue_noise_beyond_256_samples(renderer, bpt, oracle):
    """OptimizedBlueNoise falls back to the stratified path once index > 256 (samplers.cpp:27-28): 320 spp on a tiny frame"""
    w, h, spp = 24, 14, 320
    a, b = build_both(bpt, oracle, scenes.c1_week3, w, h)
    for s in (a, b):
        s.update_settings(sampling_strategy=capi.SAMPLING_BLUE_NOISE)
    film, rec, rfilm, rrec = _render(renderer, (a, b), w, h, spp)
    assert np.array_equal(bits(rec["ray_d"]), bits(rrec["ray_d"])), "primary rays (AA sample of every index 0..319) must be bit-exact"
    _radiance_report(rec, rrec, "C1 blue-noise 320 spp", 2e-3, 0.90)
    late = (np.arange(rec.shape[0]) % spp) > 256
    _radiance_report(rec, rrec, "C1 blue-noise, samples with index > 256", 2e-3, 0.90, select=late)
    _film_check(film, rfilm, "C1 blue-noise 320 spp")


# ---- BASELINE configs 3 and 4 at their benchmarked geometry: full rows, all samples -------------------------------------
def _full_rows(renderer, pair, w, h, spp, rect, what, same_rays_frac=0.995):
    film, rec, rfilm, rrec = _render(renderer, pair, w, h, spp, rect=rect)
    n = rec.shape[0]
    assert np.array_equal(bits(rec["ray_o"]), bits(rrec["ray_o"])) and np.array_equal(bits(rec["ray_d"]), bits(rrec["ray_d"]))
    same_rays = np.count_nonzero(rec["rays"] == rrec["rays"])
    print(f"{what}: same ray count {same_rays/n:.4%}, mean rays/sample {rec['rays'].mean():.2f}")
    assert same_rays >= same_rays_frac * n
    _radiance_report(rec, rrec, what, 2e-3, 0.95)
    y0, y1 = rect[1], rect[3]
    _film_check(film, rfilm, what, rows=slice(max(0, y0 - 2), min(h, y1 + 2)))


def test_c3_full_frame_rows_match_reference(renderer, bpt, oracle):
    """config 3 as benchmarked (64 instances of the level-7 icosphere = 20,971,520 triangles, 2048x1024 procedural HDR
    environment, 1920x1080, 256 spp): two full rows through the instance grid, every sample compared"""
    w, h, spp = 1920, 1080, 256
    pair = build_both(bpt, oracle, scenes.c3_instances, w, h)
    _full_rows(renderer, pair, w, h, spp, (0, 600, w, 602), "C3 full-res rows")


def test_c4_full_frame_rows_match_reference(renderer, bpt, oracle):
    """config 4 as benchmarked (nested dielectrics, Russian roulette, max depth 32, 1920x1080, 256 spp): two full rows
    through the water sphere and its marbles"""
    w, h, spp = 1920, 1080, 256
    pair = build_both(bpt, oracle, scenes.c4_nested_dielectrics, w, h)
    a, _ = pair
    assert a.get_settings().max_bounce_count == 32
    # paths through many glass interfaces accumulate more libm (expf in Beer's law) roundings: same bars, they hold
    _full_rows(renderer, pair, w, h, spp, (0, 500, w, 502), "C4 full-res rows", same_rays_frac=0.99)


# ---- north_star: "relative RMSE of the converged image at high spp" ------------------------------------------------------
@pytest.mark.parametrize("recipe,kw,what", [
    (scenes.c1_week3, {}, "C1"),
    (scenes.c2_icosphere, dict(level=5), "icosphere L5"),
])
def test_converged_image_relative_rmse(renderer, bpt, oracle, recipe, kw, what):
    """1024 spp on a 160x90 frame, GPU against the single-threaded reference under the same per-pixel seeding; compared
    as resolved images (sum w*rgb / sum w), the quantity the reference displays (raytracer.cpp:2113-2125)"""
    w, h = 160, 90
    spp = 1024 if recipe is scenes.c1_week3 else 256
    a, b = build_both(bpt, oracle, recipe, w, h, **kw)
    renderer.upload_scene(a)
    renderer.film_resize(w, h)
    passes = 4                                   # progressive: frame_count advances like render_all_tiles (:721-722)
    rfilm = np.zeros((h, w, 4), np.float32)
    for p in range(passes):
        renderer.render_pass(spp // passes, frame_count=p * (spp // passes))
        b.render_parity(w, h, spp // passes, frame_count=p * (spp // passes), film=rfilm)
    film = renderer.download_film()
    gi = film[..., :3].astype(np.float64) / film[..., 3:4]
    ri = rfilm[..., :3].astype(np.float64) / rfilm[..., 3:4]
    e = rel_rmse(gi, ri)
    worst = float(np.max(np.abs(gi - ri) / (np.abs(ri) + 1e-2 * np.mean(ri))))
    print(f"{what} converged {w}x{h}x{spp}: image relRMSE {e:.3g}, worst pixel rel err {worst:.3g}")
    assert e <= CONVERGED_REL_RMSE, f"{what}: converged-image relRMSE {e}"
    assert np.allclose(film[..., 3], rfilm[..., 3], rtol=2e-5, atol=1e-5)
