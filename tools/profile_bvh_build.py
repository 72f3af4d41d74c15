"""Build the BVH of a displaced icosphere on the device a few times (the small, fixed command that ncu wraps).
    python tools/profile_bvh_build.py [--level 8] [--repeat 3]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import buas_pathtracer_b200 as B  # noqa: E402
from buas_pathtracer_b200 import lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--level", type=int, default=8)
ap.add_argument("--repeat", type=int, default=3)
a = ap.parse_args()
tris = lib.make_displaced_icosphere(a.level)
r = B.Renderer(0)
r.build_mesh_bvh(tris[:1000])
for i in range(a.repeat):
    t0 = time.perf_counter()
    nodes, idx, ms = r.build_mesh_bvh(tris)
    wall = (time.perf_counter() - t0) * 1e3
    print(f"level {a.level}: {tris.shape[0]} triangles -> {nodes.shape[0]} nodes, device {ms:.2f} ms, call {wall:.1f} ms (incl. H2D/D2H, alloc)")
