"""GPU tests (-m gpu) of the asynchronous host interface: bpt_upload_scene_async (double-buffered scene, the next pass
switches to it) and bpt_download_film_async (front-buffer snapshot + read-back on a copy stream).  Results must be exactly
those of the synchronous calls -- the schedule changes, the work does not."""
import numpy as np
import pytest

from buas_pathtracer_b200 import scenes
from helpers import bits

pytestmark = pytest.mark.gpu


def test_async_upload_and_download_equal_the_synchronous_calls(bpt):
    w, h, spp = 160, 90, 4
    a = bpt.Scene(); scenes.c1_week3(a, w, h)
    b = bpt.Scene(); scenes.c2_icosphere(b, w, h, level=4)
    r = bpt.Renderer(0)
    r.film_resize(w, h)
    # reference results with the synchronous interface
    want = {}
    for name, s in (("a", a), ("b", b)):
        r.upload_scene(s)
        r.film_clear()
        r.render_pass(spp)
        want[name] = r.download_film()
    assert not np.allclose(want["a"], want["b"])
    # a frame loop that never waits: scenes alternate, each film is read back asynchronously into its own pinned array
    outs = [np.zeros((h, w, 4), np.float32) for _ in range(6)]
    for o in outs:
        r.host_register(o)
    order = ["a", "b", "b", "a", "b", "a"]
    for i, name in enumerate(order):
        r.upload_scene_async(a if name == "a" else b)
        r.film_clear()
        r.render_pass(spp)
        r.download_film_async(outs[i])
    r.wait_download()
    r.sync()
    for i, name in enumerate(order):
        # same rays, same per-sample values; only the order of the film's float atomics differs between runs
        assert np.allclose(outs[i], want[name], rtol=1e-5, atol=1e-6), (i, name)
    for o in outs:
        r.host_unregister(o)
    # an upload that is never rendered is simply replaced by the next one
    r.upload_scene_async(a)
    r.upload_scene_async(b)
    r.film_clear()
    r.render_pass(spp)
    assert np.allclose(r.download_film(), want["b"], rtol=1e-5, atol=1e-6)
    # trace after an async upload uses the new scene too
    r.upload_scene_async(a)
    from buas_pathtracer_b200 import capi
    rays = np.zeros(4, capi.RAY_DTYPE)
    rays["o"] = (0, 4, -10); rays["d"] = (0, 0, 1); rays["max_t"] = np.finfo(np.float32).max
    hits = r.trace(rays)
    assert np.all(hits["primitive"] == 1) and np.allclose(hits["t"], 6.0)      # the r = 4 sphere of week_3_scene at (0, 4, 0)
    r.close()


def test_download_async_snapshot_is_not_disturbed_by_the_next_pass(bpt):
    w, h = 128, 72
    s = bpt.Scene(); scenes.c1_week3(s, w, h)
    r = bpt.Renderer(0)
    r.upload_scene(s)
    r.film_resize(w, h)
    r.render_pass(2, frame_count=0)
    first = r.download_film()
    out = np.zeros((h, w, 4), np.float32)
    r.host_register(out)
    r.film_clear()
    r.render_pass(2, frame_count=0)
    r.download_film_async(out)          # snapshot of pass 1 ...
    r.render_pass(2, frame_count=2)     # ... while pass 2 already accumulates into the film
    r.wait_download()
    assert np.allclose(out, first, rtol=1e-5, atol=1e-6)
    both = r.download_film()
    assert float(both[..., 3].sum()) > 1.9 * float(first[..., 3].sum())
    r.host_unregister(out)
    r.close()


def test_back_to_back_passes_equal_synchronised_passes(bpt):
    """Passes enqueued without waiting overlap on the device (the next pass's batches start under the previous pass's
    kernel tails, consecutive one-batch passes alternate between the batch streams).  What touches the film stays ordered:
    the snapshot taken behind pass k holds passes 0..k exactly -- nothing of pass k+1."""
    w, h, spp, passes = 320, 180, 8, 6
    s = bpt.Scene(); scenes.c2_icosphere(s, w, h, level=5)
    r = bpt.Renderer(0)
    r.upload_scene(s)
    r.film_resize(w, h)
    want = []
    for p in range(passes):                       # the synchronous schedule: wait for every pass
        r.render_pass(spp, frame_count=p * spp)
        want.append(r.download_film())
    assert float(want[1][..., 3].sum()) > 1.9 * float(want[0][..., 3].sum())
    outs = [np.zeros((h, w, 4), np.float32) for _ in range(passes)]
    for o in outs:
        r.host_register(o)
    for rep in range(2):                          # the second round runs with the path state of both streams allocated
        r.film_clear()
        for p in range(passes):
            r.render_pass(spp, frame_count=p * spp)
            r.download_film_async(outs[p])
            if p == 2:
                r.wait_download()                 # a host wait in the middle of the loop changes nothing either
        r.wait_download()
        r.sync()
        for p in range(passes):
            assert np.allclose(outs[p], want[p], rtol=1e-5, atol=1e-6), (rep, p)
    for o in outs:
        r.host_unregister(o)
    r.close()
