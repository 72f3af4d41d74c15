"""Render a few passes of one BASELINE config -- the small, fixed command that ncu / compute-sanitizer wrap.
    python tools/profile_pass.py [--config c2] [--passes 2] [--spp N] [--w W --h H]
Prints per-stage CUDA-event times of the last pass (never quote a number printed under a profiler)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import buas_pathtracer_b200 as B  # noqa: E402
from buas_pathtracer_b200 import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--spp", type=int, default=0)
ap.add_argument("--w", type=int, default=0)
ap.add_argument("--h", type=int, default=0)
ap.add_argument("--rows", type=str, default="")      # "y0:y1" sub-rect
ap.add_argument("--stats", action="store_true")
ap.add_argument("--no-detail", action="store_true")
ap.add_argument("--sync-each", action="store_true")   # host synchronisation after every pass (no overlap between passes)
ap.add_argument("--world", type=int, default=1)     # render rank 0's share of an N-rank interleaved row partition
a = ap.parse_args()
cfg = scenes.CONFIGS[a.config]
w, h, spp = a.w or cfg["w"], a.h or cfg["h"], a.spp or cfg["spp"]
s = B.Scene()
cfg["build"](s, w, h)
r = B.Renderer(0)
r.upload_scene(s)
r.film_resize(w, h)
rect = None
if a.rows:
    y0, y1 = [int(v) for v in a.rows.split(":")]
    rect = (0, y0, w, y1)
r.set_detailed_timing(not a.no_detail)
if a.stats:
    r.stats_enable(True)
bands = None
if a.world > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    bands = bench.my_rows(h, 0, a.world)
import time  # noqa: E402


def one_pass():
    if bands:
        r.render_pass_bands(spp, bands, frame_count=0)
    else:
        r.render_pass(spp, rect=rect, frame_count=0)


period_ms = None
if a.no_detail and a.passes > 1 and not a.sync_each:
    # back-to-back passes, one host synchronisation at the end (consecutive passes overlap their kernel tails)
    for i in range(3):       # warm-up in the same mode: back-to-back passes take another batch shape (path state gets allocated for it once)
        one_pass()
    r.sync()
    t0 = time.perf_counter()
    for i in range(a.passes - 1):
        one_pass()
    r.sync()
    period_ms = (time.perf_counter() - t0)*1e3/(a.passes - 1)
    r.get_stats(reset=True)
    one_pass()
    r.sync()
else:
    for i in range(a.passes):
        r.get_stats(reset=True)
        one_pass()
        r.sync()
t = r.pass_timing()
st = r.get_stats().as_dict()
d = {k: round(getattr(t, k), 3) if isinstance(getattr(t, k), float) else getattr(t, k) for k, _ in t._fields_}
if period_ms is not None:
    d["single_pass_ms"] = d["total_ms"]
    d["total_ms"] = round(period_ms, 3)     # pass period of back-to-back passes
    t.total_ms = period_ms
print(d)
print(st)
print("Mrays/s", st["rays"] / t.total_ms / 1e3, "Msamples/s", st["samples"] / t.total_ms / 1e3)
