"""SURVEY 8f rank 1 -- resolve/post-process + BMP ("Render to bitmap").  Checked against a numpy port of
raytracer.cpp:2103-2172 (oracle/resolve_port.py; parity unpinned by the reference, tolerance +-1 LSB)."""
import os

import numpy as np
import pytest

from buas_pathtracer_b200 import scenes


def test_bitmap_writer_matches_reference_layout(bpt, tmp_path):
    from oracle import resolve_port
    rng = np.random.RandomState(0)
    px = rng.randint(0, 2 ** 32, size=(7, 13), dtype=np.uint64).astype(np.uint32)
    path = str(tmp_path / "t.bmp")
    rc = bpt.load_library().bpt_write_bitmap(path.encode(), px.ctypes.data, 13, 7)
    assert rc == 0
    assert open(path, "rb").read() == resolve_port.bitmap_bytes(px)


def test_resolve_port_basics():
    from oracle import resolve_port
    film = np.zeros((2, 3, 4), np.float32)
    film[0, 0] = (np.nan, 0, 0, 1)            # NaN -> cyan
    film[0, 1] = (0, 0, 0, 0)                 # no weight -> black
    film[0, 2] = (5, 5, 5, -1)                # negative weight -> magenta
    film[1, 0] = (1e9, 1e9, 1e9, 1)           # saturates to white
    out = resolve_port.resolve_bgra8(film)
    assert out[0, 0] == 0xFF00FFFF and out[0, 1] == 0xFF000000 and out[0, 2] == 0xFFFF00FF and out[1, 0] == 0xFFFFFFFF


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(tonemapping=False, srgb_transform=False), dict(exposure=1.5, contrast=0.4, midpoint=0.45),
                                dict(dither=True)])
def test_gpu_resolve_matches_port(bpt, renderer, kw):
    from oracle import resolve_port
    w, h = 160, 90
    s = bpt.Scene()
    scenes.c1_week3(s, w, h)
    renderer.upload_scene(s)
    renderer.film_resize(w, h)
    renderer.render_pass(8)
    film = renderer.download_film()
    kw = dict(kw)
    if kw.pop("dither", False):
        kw["dither"] = np.random.RandomState(3).randint(0, 256, size=(64, 64, 3)).astype(np.uint8)
    g = renderer.resolve_bgra8(**kw)
    r = resolve_port.resolve_bgra8(film, **kw)
    gc = np.stack([(g >> 16) & 255, (g >> 8) & 255, g & 255], axis=-1).astype(np.int32)
    rc = np.stack([(r >> 16) & 255, (r >> 8) & 255, r & 255], axis=-1).astype(np.int32)
    diff = np.abs(gc - rc)
    assert diff.max() <= 1, f"max channel difference {diff.max()}"
    assert np.count_nonzero(diff) <= 0.002 * diff.size + (0.02 * diff.size if "dither" in kw else 0)
    assert (g >> 24 == 255).all()
    assert gc.max() > 100          # not a black frame
