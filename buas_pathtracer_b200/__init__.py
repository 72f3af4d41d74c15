"""B200-native path-tracing core for BUAS-Pathtracer: Python harness over the C ABI (include/bpt.h).

The product is `libbpt.so` (hand-written CUDA for sm_100a + the host scene model).  This package only loads it
through ctypes; it contains no renderer of its own and no CPU fallback.
"""
from .capi import (HostScene, Material, Camera, Settings, M4x4Inv, FilterCache, Stats, PassTiming,  # noqa: F401
                   RAY_DTYPE, HIT_DTYPE, RECORD_DTYPE, BVH_NODE_DTYPE, HIT_MISS, HIT_PLANE,
                   TRACE_CLOSEST, TRACE_OCCLUSION)
from .lib import load_library, library_path, build_library, Scene, Renderer, BptError, sampler_tables  # noqa: F401
from . import scenes  # noqa: F401
