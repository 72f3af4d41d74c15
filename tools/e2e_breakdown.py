"""Where the end-to-end step goes: the frame loop of bench.py's e2e leg with its parts switched on one by one (wall clock,
N steps, one final wait).  python tools/e2e_breakdown.py [--steps 10] [--config c2]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import buas_pathtracer_b200 as B  # noqa: E402
from buas_pathtracer_b200 import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--config", default="c2")
a = ap.parse_args()
cfg = scenes.CONFIGS[a.config]
w, h, spp = cfg["w"], cfg["h"], cfg["spp"]
s = B.Scene()
cfg["build"](s, w, h)
r = B.Renderer(0)
r.upload_scene(s)
r.film_resize(w, h)
hosts = [np.empty((h, w, 4), np.float32) for _ in range(2)]
for hf in hosts:
    r.host_register(hf)
for _ in range(3):
    r.render_pass(spp); r.sync()


def loop(upload, download, sync_each, clear=True):
    r.sync()
    t0 = time.perf_counter()
    for i in range(a.steps):
        if upload == "async":
            r.upload_scene_async(s)
        elif upload == "sync":
            r.upload_scene(s)
        if clear:
            r.film_clear()
        r.render_pass(spp)
        if download == "async":
            r.download_film_async(hosts[i & 1])
        elif download == "sync":
            r.download_film(hosts[i & 1])
        if sync_each:
            r.sync()
    r.wait_download()
    r.sync()
    return (time.perf_counter() - t0) / a.steps * 1e3


for name, kw in [("render only, sync each step", dict(upload=None, download=None, sync_each=True)),
                 ("render only, one final sync", dict(upload=None, download=None, sync_each=False)),
                 ("+ async upload", dict(upload="async", download=None, sync_each=False)),
                 ("+ async download", dict(upload=None, download="async", sync_each=False)),
                 ("+ async upload + async download (the e2e loop)", dict(upload="async", download="async", sync_each=False)),
                 ("sync upload + sync download (round-1 loop)", dict(upload="sync", download="sync", sync_each=True))]:
    ms = loop(**kw)
    print(f"{name:55s} {ms:8.3f} ms/step")
