"""SURVEY 8f rank 1 -- resolve/post-process + BMP ("Render to bitmap"), pinned to the reference itself:

* bpt_write_bitmap against the reference's own write_bitmap (assets.cpp:693-724, compiled in oracle/_ref): same bytes;
* k_resolve (bpt_resolve_bgra8) against the reference's own display-loop resolve -- the loop of raytracer.cpp:2103-2173, cut
  out of the reference's source at oracle build time (oracle/tools/slice_resolve.py) and compiled as is.  Tolerance: the
  reference's remap_tpdf uses the SSE rsqrt approximation (12 bits) and its expf/powf are glibc's, the device uses an
  exact rsqrt and double-evaluated exp/pow, so a channel may differ by 1 LSB where the float value sits on a rounding edge:
  max difference 1, at most 2 % of the channels.
* the numpy port (oracle/resolve_port.py) is kept only for the path the reference does not have (no dither tile) and is
  itself checked against the reference's resolve here."""
import numpy as np
import pytest

from buas_pathtracer_b200 import scenes


def _channels(p):
    return np.stack([(p >> 16) & 255, (p >> 8) & 255, p & 255], axis=-1).astype(np.int32)


def _test_film(seed=1, h=40, w=56):
    rng = np.random.RandomState(seed)
    film = (rng.rand(h, w, 4) ** 3 * 6).astype(np.float32)
    film[..., 3] = rng.rand(h, w).astype(np.float32) * 4 + 0.5
    film[0, 0] = (np.nan, 0, 0, 1)            # NaN -> cyan
    film[0, 1] = (0, 0, 0, 0)                 # no weight -> black
    film[0, 2] = (5, 5, 5, -1)                # negative weight -> magenta
    film[1, 0] = (1e9, 1e9, 1e9, 1)           # saturates to white
    return film


def test_bitmap_writer_matches_the_references_own(bpt, oracle, tmp_path):
    rng = np.random.RandomState(0)
    for h, w in ((7, 13), (1, 1), (64, 64)):
        px = rng.randint(0, 2 ** 32, size=(h, w), dtype=np.uint64).astype(np.uint32)
        ours, theirs = str(tmp_path / "ours.bmp"), str(tmp_path / "theirs.bmp")
        assert bpt.load_library().bpt_write_bitmap(ours.encode(), px.ctypes.data, w, h) == 0
        oracle.write_bitmap(theirs, px)
        assert open(ours, "rb").read() == open(theirs, "rb").read()


def test_resolve_port_agrees_with_the_references_resolve(oracle):
    """the numpy port (used for the no-dither path only) against the reference's own loop, incl. the special pixels"""
    from oracle import resolve_port
    film = _test_film()
    dither = np.random.RandomState(3).randint(0, 256, size=(16, 16, 3)).astype(np.uint8)
    for kw in (dict(), dict(tonemapping=False, srgb_transform=False), dict(exposure=1.5, contrast=0.4, midpoint=0.45)):
        r = oracle.resolve_bgra8(film, dither=dither, **kw)
        p = resolve_port.resolve_bgra8(film, dither=dither, **kw)
        assert r[0, 0] == 0xFF00FFFF and r[0, 1] == 0xFF000000 and r[0, 2] == 0xFFFF00FF and r[1, 0] == 0xFFFFFFFF
        d = np.abs(_channels(r) - _channels(p))
        assert d.max() <= 1 and np.count_nonzero(d) <= 0.01 * d.size


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(tonemapping=False, srgb_transform=False), dict(exposure=1.5, contrast=0.4, midpoint=0.45),
                                dict(exposure=-1.0, contrast=0.8, midpoint=0.6)])
def test_gpu_resolve_matches_the_references_resolve(bpt, renderer, oracle, kw):
    w, h = 160, 90
    s = bpt.Scene()
    scenes.c1_week3(s, w, h)
    renderer.upload_scene(s)
    renderer.film_resize(w, h)
    renderer.render_pass(8)
    film = renderer.download_film()
    dither = np.random.RandomState(3).randint(0, 256, size=(64, 64, 3)).astype(np.uint8)
    g = renderer.resolve_bgra8(dither=dither, **kw)
    r = oracle.resolve_bgra8(film, dither=dither, **kw)
    diff = np.abs(_channels(g) - _channels(r))
    print(f"resolve {kw}: {np.count_nonzero(diff)} of {diff.size} channels differ, max {diff.max()}")
    assert diff.max() <= 1, f"max channel difference {diff.max()}"
    assert np.count_nonzero(diff) <= 0.02 * diff.size
    assert (g >> 24 == 255).all()
    assert _channels(g).max() > 100          # not a black frame


@pytest.mark.gpu
def test_gpu_resolve_without_dither_matches_port(bpt, renderer):
    """dither == NULL is this library's addition (the reference always dithers): checked against the numpy port"""
    from oracle import resolve_port
    w, h = 160, 90
    s = bpt.Scene()
    scenes.c1_week3(s, w, h)
    renderer.upload_scene(s)
    renderer.film_resize(w, h)
    renderer.render_pass(8)
    film = renderer.download_film()
    g = renderer.resolve_bgra8()
    r = resolve_port.resolve_bgra8(film)
    diff = np.abs(_channels(g) - _channels(r))
    assert diff.max() <= 1 and np.count_nonzero(diff) <= 0.002 * diff.size
