"""CPU tests that pin the ORACLE: (1) the reference's own known-answer tests (UnitTests/main.cpp:733-787) evaluated
through the reference's real intersect_scene, (2) the committed golden fixtures (tests/golden/golden_v1.npz, produced by
tests/golden/make_golden.py from the same reference code) must be reproduced bit for bit."""
import hashlib
import os
import sys

import numpy as np
import pytest

from buas_pathtracer_b200 import capi, scenes

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")
EPS = 0.001     # UnitTests/main.cpp EPSILON


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _ray(o, d, max_t=np.finfo(np.float32).max):
    r = np.zeros(1, capi.RAY_DTYPE)
    r["o"] = o; r["d"] = d; r["max_t"] = max_t
    return r


def test_reference_sphere_kats(oracle):
    """UnitTests/main.cpp:764-785: sphere r=4 at the origin"""
    s = oracle.RefScene()
    m = s.add_diffuse_material((1, 1, 1), 1.0)
    s.add_sphere(m, 4.0)
    s.create_scene_bvh()
    for ox, t_near in ((0.0, 6.0), (2.0, 6.53589838), (4.0, 10.0)):
        h = s.trace(_ray((ox, 0, 10), (0, 0, -1)))
        if ox == 4.0:
            # tangent ray: discriminant is ~0; the renderer's version may or may not accept, both consistent with the KAT
            assert h["primitive"][0] in (1, capi.HIT_MISS)
            if h["primitive"][0] == 1:
                assert abs(h["t"][0] - t_near) < 0.05
        else:
            assert h["primitive"][0] == 1 and abs(h["t"][0] - t_near) < EPS
    assert s.trace(_ray((6, 0, 10), (0, 0, -1)))["primitive"][0] == capi.HIT_MISS
    # t_far = 14 / 13.4641016: seen from inside the sphere the renderer returns the far root
    assert abs(s.trace(_ray((0, 0, 0), (0, 0, -1)))["t"][0] - 4.0) < EPS
    h = s.trace(_ray((0, 0, 2), (0, 0, -1)))
    assert abs(h["t"][0] - 6.0) < EPS


def test_reference_plane_kats(oracle):
    """UnitTests/main.cpp:738-761"""
    s = oracle.RefScene()
    m = s.add_diffuse_material((1, 1, 1), 1.0)
    s.add_plane(m, (0.0, 0.707106781, 0.707106781), -5.0)
    s.create_scene_bvh()
    for o, t in (((0, 0, 10), 17.0710678), ((0, 2, 10), 19.0710678), ((5, -10, 10), 7.07106781)):
        h = s.trace(_ray(o, (0, 0, -1)))
        assert h["primitive"][0] == capi.HIT_PLANE and abs(h["t"][0] - t) < EPS
    s2 = oracle.RefScene()
    m = s2.add_diffuse_material((1, 1, 1), 1.0)
    s2.add_plane(m, (0, 1, 0), 0.0)
    s2.add_plane(m, (0, 0, 1), 0.0)
    s2.create_scene_bvh()
    assert s2.trace(_ray((0, 0, 0), (0, 0, -1)))["primitive"][0] == capi.HIT_MISS      # parallel / starts on plane
    assert s2.trace(_ray((0, 0.5, -1), (0, 0, -1)))["primitive"][0] == capi.HIT_MISS    # plane behind the ray


def test_oracle_reproduces_golden_samplers(oracle, golden):
    assert np.array_equal(make_golden.sampler_kat(), golden["sampler_kat"])


@pytest.mark.parametrize("name", list(make_golden.GOLDEN_SCENES))
def test_oracle_reproduces_golden_scenes(oracle, bpt, golden, name):
    s, w, h, spp = make_golden.build_ref(name)
    film, rec = s.render_parity(w, h, spp, records=True, salt=0x1234)
    assert rec.tobytes() == golden[f"{name}_records"].tobytes()
    assert film.tobytes() == golden[f"{name}_film"].tobytes()
    hits2 = s.trace(golden[f"{name}_rays2"], capi.TRACE_CLOSEST)
    assert hits2.tobytes() == golden[f"{name}_hits2"].tobytes()
    n, i = s.scene_bvh()
    assert hashlib.sha256(n.tobytes() + i.tobytes()).digest() == golden[f"{name}_tlas_sha"].tobytes()


@pytest.mark.parametrize("level", range(5))
def test_product_blas_matches_golden_hash(bpt, golden, level):
    """the product's host BVH builder against the committed hash (works without the oracle library)"""
    s = bpt.Scene()
    n, i, t = s.mesh_bvh(s.create_mesh(bpt.lib.make_displaced_icosphere(level, 0.08)))
    assert hashlib.sha256(n.tobytes() + i.tobytes() + t.tobytes()).digest() == golden[f"blas_sha_l{level}"].tobytes()


@pytest.mark.parametrize("name", list(make_golden.GOLDEN_SCENES))
def test_product_tlas_matches_golden_hash(bpt, golden, name):
    recipe, w, h, spp, kw = make_golden.GOLDEN_SCENES[name]
    s = bpt.Scene()
    recipe(s, w, h, **kw)
    n, i = s.scene_bvh()
    assert hashlib.sha256(n.tobytes() + i.tobytes()).digest() == golden[f"{name}_tlas_sha"].tobytes()


def test_standalone_inputs_library_matches_the_product(bpt):
    """oracle/_ref/libbpt_inputs.so is the product's procedural_inputs.cpp built standalone (the reference arm of bench.py
    builds its scenes from it without loading libbpt.so): same bytes"""
    from oracle import ref_inputs
    from buas_pathtracer_b200 import lib
    for level in (0, 2, 5):
        a, b = ref_inputs.make_displaced_icosphere(level), lib.make_displaced_icosphere(level)
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    a, b = ref_inputs.make_procedural_skydome(128, 64), lib.make_procedural_skydome(128, 64)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_reference_arm_of_bench_never_loads_the_product_library():
    """bench.py --impl reference must run with the product library absent (BENCH `reference.native_so_loaded`)"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, BPT_LIBRARY="/nonexistent/libbpt.so")
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "c1",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "reference"
    assert line["config"]["spp_per_step_run"] >= 1 and 1.0 < line["rays_per_sample"] < 12.0
