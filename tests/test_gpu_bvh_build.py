"""GPU BVH construction: bpt_build_mesh_bvh_device must reproduce, node for node, the output of the
reference's own create_bvh_for_mesh (compared directly through the oracle) and of the host builder (which
tests/test_host_parity.py pins memcmp-equal to the reference too): same node array (depth-first numbering, bounds, split
axes, leaf ranges) and same leaf-order index array."""
import numpy as np
import pytest

import buas_pathtracer_b200 as B
from buas_pathtracer_b200 import capi, lib, scenes

pytestmark = pytest.mark.gpu


def reference_build(oracle, tris, method=None):
    """the reference's own create_bvh_for_mesh (bvh.cpp:342-426) through the oracle"""
    s = oracle.RefScene()
    m = s.create_mesh(tris, method=method)
    nodes, idx, _ = s.mesh_bvh(m)
    return nodes, idx


def host_build(tris, method=None):
    s = B.Scene()
    m = s.create_mesh(tris, method=method)
    nodes, idx, _ = s.mesh_bvh(m)
    return nodes, idx


def assert_same_bvh(dn, di, hn, hi, what):
    assert dn.shape == hn.shape, f"{what}: node count {dn.shape[0]} vs {hn.shape[0]}"
    if dn.tobytes() != hn.tobytes():
        a = dn.view(np.uint8).reshape(-1, 32); b = hn.view(np.uint8).reshape(-1, 32)
        bad = np.nonzero(np.any(a != b, axis=1))[0]
        raise AssertionError(f"{what}: {bad.size} of {dn.shape[0]} nodes differ, first {bad[0]}: device {dn[bad[0]]} host {hn[bad[0]]}")
    assert np.array_equal(di, hi), f"{what}: leaf-order indices differ at {np.nonzero(di != hi)[0][:5]}"


@pytest.mark.parametrize("level", [0, 1, 3, 5, 7])
def test_device_bvh_equals_host_bvh_icosphere(renderer, oracle, level):
    tris = lib.make_displaced_icosphere(level)
    dn, di, ms = renderer.build_mesh_bvh(tris)
    hn, hi = host_build(tris)
    assert_same_bvh(dn, di, hn, hi, f"icosphere level {level}")
    rn, ri = reference_build(oracle, tris)               # and directly against the reference's builder
    assert_same_bvh(dn, di, rn, ri, f"icosphere level {level} vs the reference")
    print(f"icosphere level {level}: {tris.shape[0]} triangles, {dn.shape[0]} nodes, device build {ms:.2f} ms")


@pytest.mark.parametrize("level", [0, 2, 4, 6, 8])
def test_device_midpoint_bvh_equals_host(renderer, oracle, level):
    """BVH_MidpointSplit (what the reference builds for OBJ files, raytracer.cpp:154) on the device"""
    tris = lib.make_displaced_icosphere(level)
    dn, di, ms = renderer.build_mesh_bvh(tris, capi.BVH_MIDPOINT_SPLIT)
    hn, hi = host_build(tris, capi.BVH_MIDPOINT_SPLIT)
    assert_same_bvh(dn, di, hn, hi, f"midpoint, icosphere level {level}")
    if level <= 6:
        rn, ri = reference_build(oracle, tris, capi.BVH_MIDPOINT_SPLIT)
        assert_same_bvh(dn, di, rn, ri, f"midpoint, icosphere level {level} vs the reference")
    rng = np.random.RandomState(level)
    soup = np.ascontiguousarray((rng.rand(3000, 9) ** 5) * 40, np.float32)         # skewed: lopsided midpoints, forced leaves
    dn, di, _ = renderer.build_mesh_bvh(soup, capi.BVH_MIDPOINT_SPLIT)
    hn, hi = host_build(soup, capi.BVH_MIDPOINT_SPLIT)
    assert_same_bvh(dn, di, hn, hi, "midpoint, skewed soup")
    print(f"midpoint level {level}: {tris.shape[0]} triangles, device build {ms:.2f} ms")


def test_device_bvh_degenerate_inputs(renderer, oracle):
    rng = np.random.RandomState(3)
    cases = {
        "one triangle": rng.rand(1, 9),
        "four triangles": rng.rand(4, 9),
        "five triangles": rng.rand(5, 9),
        "all identical": np.tile(rng.rand(1, 9), (300, 1)),
        "two clusters of duplicates": np.concatenate([np.tile(rng.rand(1, 9), (100, 1)), np.tile(rng.rand(1, 9) + 5, (77, 1))]),
        "collinear centroids": np.stack([np.concatenate([[i, 0, 0], [i + .5, 0, 0], [i, .5, 0]]) for i in range(1000)]),
        "grid with many equal coordinates": np.stack([np.concatenate([[x, y, 0], [x + 1, y, 0], [x, y + 1, 0]])
                                                      for x in range(40) for y in range(40)]),
        "random soup": rng.randn(20000, 9) * 3,
        "huge and tiny": np.concatenate([rng.randn(500, 9) * 1e6, rng.randn(500, 9) * 1e-6]),
    }
    for what, t in cases.items():
        t = np.ascontiguousarray(t, np.float32)
        dn, di, _ = renderer.build_mesh_bvh(t)
        hn, hi = host_build(t)
        assert_same_bvh(dn, di, hn, hi, what)
        rn, ri = reference_build(oracle, t)
        assert_same_bvh(dn, di, rn, ri, what + " vs the reference")


def test_device_bvh_signed_zero_ties_are_the_documented_limit(renderer):
    """The one place the device build may legitimately differ: when +0 and -0 tie for an extreme, the reference's
    ternary min/max keeps the LAST of the tied values, an atomic min keeps -0.  Everything except the sign bit of such
    zeros must still agree, and meshes without mixed-sign zero ties (every other test) agree bit for bit."""
    rng = np.random.RandomState(11)
    t = rng.rand(400, 9).astype(np.float32)
    t[:, 0::3] = np.where(rng.rand(400, 3) < 0.5, np.float32(0.0), np.float32(-0.0))      # all x coordinates are +-0
    dn, di, _ = renderer.build_mesh_bvh(t)
    hn, hi = host_build(t)
    assert dn.shape == hn.shape and np.array_equal(di, hi)
    a = dn.view(np.uint32).reshape(-1, 8).copy(); b = hn.view(np.uint32).reshape(-1, 8).copy()
    diff = a != b
    # only float fields (columns 0..5) may differ, and only between +0 and -0
    assert not diff[:, 6:].any()
    assert np.all((a[diff] & 0x7FFFFFFF) == 0) and np.all((b[diff] & 0x7FFFFFFF) == 0)
    print(f"signed-zero ties: {int(diff.any(axis=1).sum())} of {dn.shape[0]} nodes differ in the sign of a zero")


def test_render_through_device_built_bvh(renderer):
    """a mesh whose BVH came from the device build renders the same film as the host-built one"""
    tris = lib.make_displaced_icosphere(5)
    dn, di, _ = renderer.build_mesh_bvh(tris)
    films = []
    for use_device in (False, True):
        s = B.Scene()
        if use_device:      # same recipe, but the mesh takes the device-built arrays instead of running the host build
            s.create_mesh = lambda positions, normals=None: s.create_mesh_with_bvh(positions, dn, di, normals)
        scenes.c2_icosphere(s, 160, 90, level=5, tris=tris)
        renderer.upload_scene(s)
        renderer.film_resize(160, 90)
        renderer.render_pass(4)
        films.append(renderer.download_film())
    assert np.allclose(films[0], films[1], rtol=1e-5, atol=1e-6)


def test_device_bvh_full_size_c2_mesh(renderer, oracle):
    """BASELINE config 2's mesh: 1,310,720 triangles, bit-identical to the host build and to the reference's own
    create_bvh_for_mesh, and how long each takes"""
    import time
    tris = lib.make_displaced_icosphere(8)
    renderer.build_mesh_bvh(tris[:1000])                  # warm-up (context, allocator)
    dn, di, ms = renderer.build_mesh_bvh(tris)
    t0 = time.perf_counter()
    hn, hi = host_build(tris)
    host_ms = (time.perf_counter() - t0) * 1e3
    assert_same_bvh(dn, di, hn, hi, "icosphere level 8")
    t0 = time.perf_counter()
    rn, ri = reference_build(oracle, tris)
    ref_ms = (time.perf_counter() - t0) * 1e3
    assert_same_bvh(dn, di, rn, ri, "icosphere level 8 vs the reference")
    print(f"icosphere level 8: {tris.shape[0]} triangles, {dn.shape[0]} nodes: device build {ms:.2f} ms, "
          f"host build (incl. scene bookkeeping) {host_ms:.0f} ms, the reference's create_bvh_for_mesh {ref_ms:.0f} ms")
