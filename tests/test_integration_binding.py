"""INTEGRATION.md's binding is real code: oracle/integration/raytracer_bpt.inl is compiled AS WRITTEN against the
reference's own Scene / RenderParameters / g_integrators (oracle/Makefile `integration`, inside the reference's
raytracer.cpp translation unit) and linked with libbpt.so.  Here (-m gpu) one of the reference's own g_scenes[] entries is
loaded by the reference's load_scene, mirrored by bpt_mirror_scene, rendered on the GPU through render_all_tiles_bpt, and the
front buffer it fills is compared with the reference's own CPU render of the same built-in scene (per-pixel seeding, salt =
Scene::total_frame_index as the binding passes it)."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import rel_rmse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libbpt_integration.so")


def test_binding_snippet_is_quoted_in_integration_md():
    """INTEGRATION.md shows the file that is compiled, not a paraphrase"""
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    inl = open(os.path.join(ROOT, "oracle", "integration", "raytracer_bpt.inl")).read()
    body = inl[inl.index('extern "C" {'):]
    assert body.strip() in md, "INTEGRATION.md and oracle/integration/raytracer_bpt.inl have drifted apart"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["Week 3", "Week 5", "Nested Dielectrics"])
def test_binding_renders_a_reference_scene_on_the_gpu(oracle, name):
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libbpt_integration.so not built (make -C oracle integration)")
    L = C.CDLL(LIB)
    L.ref_scene_create.restype = C.c_void_p
    L.integration_render_builtin.restype = C.c_int
    L.integration_render_builtin.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
    w, h, spp = 160, 90, 4
    film = np.zeros((h, w, 4), np.float32)
    s = L.ref_scene_create()
    rc = L.integration_render_builtin(C.c_void_p(s), name.encode(), w, h, spp, 1, 0, film.ctypes.data)
    assert rc == 0
    ref = oracle.RefScene()
    ref.load_builtin(name, w, h)
    rfilm, _ = ref.render_parity(w, h, spp, frame_count=0, salt=0)
    assert np.all(np.isfinite(film)) and float(film[..., 3].sum()) > 0
    assert np.allclose(film[..., 3], rfilm[..., 3], rtol=2e-5, atol=1e-6), "filter weights differ"
    e = rel_rmse(film[..., :3].astype(np.float64), rfilm[..., :3].astype(np.float64))
    print(f"binding, built-in scene {name!r}: film relRMSE {e:.3g}")
    assert e <= 2e-3
