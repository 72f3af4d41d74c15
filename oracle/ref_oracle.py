"""ORACLE-ONLY (test infrastructure): ctypes loader for oracle/_ref/libbpt_ref.so -- the reference's own
C++ compiled unmodified (oracle/Makefile).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module.  The product package never does."""
import ctypes as C
import os
import subprocess

import numpy as np

from buas_pathtracer_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libbpt_ref.so")
REFERENCE_ROOT = "/root/reference"


def build(force=False):
    """(Re)build the oracle when the reference tree is present; on the GPU box the prebuilt .so is used."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "Raytracer")):
        args = ["make", "-C", HERE, "-j8", "all"] + (["-B"] if force else [])
        subprocess.run(args, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        # INTEGRATION.md's binding compiled against the reference's Scene and linked with the product library (test-only);
        # needs libbpt.so, so it is best-effort here
        if os.path.exists(os.path.join(HERE, "..", "buas_pathtracer_b200", "libbpt.so")):
            subprocess.run(["make", "-C", HERE, "integration"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return os.path.exists(LIB_PATH)


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"oracle library missing: {LIB_PATH} (build with `make -C oracle` where /root/reference exists)")
        L = C.CDLL(LIB_PATH)
        vp, P = C.c_void_p, C.POINTER
        L.ref_trace.restype = C.c_int
        L.ref_trace.argtypes = [vp, C.c_uint32, vp, C.c_int, C.c_uint32, vp]
        L.ref_get_stats.restype = C.c_int
        L.ref_get_stats.argtypes = [P(capi.Stats), C.c_int]
        L.ref_render_parity.restype = C.c_int
        L.ref_render_parity.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_uint32, C.c_uint32, C.c_uint32, vp]
        L.ref_render_threaded.restype = C.c_int
        L.ref_render_threaded.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp, P(C.c_double), P(capi.Stats)]
        L.ref_load_builtin_scene.restype = C.c_int
        L.ref_load_builtin_scene.argtypes = [vp, C.c_char_p, C.c_uint32, C.c_uint32]
        L.ref_sizeof.restype = C.c_int
        L.ref_sizeof.argtypes = [C.c_char_p]
        L.ref_get_sampler_tables.restype = C.c_int
        L.ref_get_sampler_tables.argtypes = [vp, vp, vp, vp]
        L.ref_kat_cosine_hemisphere.argtypes = [capi.c_float3, C.c_float * 2, capi.c_float3]
        L.ref_kat_hemisphere.argtypes = [capi.c_float3, C.c_float * 2, capi.c_float3]
        L.ref_kat_fresnel.restype = C.c_float
        L.ref_kat_fresnel.argtypes = [C.c_float, C.c_float, C.c_float, P(C.c_float)]
        L.ref_kat_random_seed.argtypes = [C.c_uint32, C.c_uint32 * 4]
        L.ref_kat_sample_2d.argtypes = [C.c_uint32 * 4, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_float * 2]
        L.ref_kat_sample_1d.restype = C.c_float
        L.ref_kat_sample_1d.argtypes = [C.c_uint32 * 4, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32]
        L.ref_kat_filter.restype = C.c_float
        L.ref_kat_filter.argtypes = [C.c_char_p, C.c_float]
        _lib = L
    return _lib


class RefScene(capi.HostScene):
    """The reference's Scene behind the same builder calls as the product's HostScene."""

    def __init__(self):
        super().__init__(lib(), "ref_")

    def load_builtin(self, name, w, h):
        rc = self.lib.ref_load_builtin_scene(self.handle, name.encode(), w, h)
        if rc != 0:
            raise RuntimeError(f"unknown built-in scene {name!r}")

    def trace(self, rays, mode=capi.TRACE_CLOSEST, ignored_primitive=0):
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=capi.HIT_DTYPE)
        rc = self.lib.ref_trace(self.handle, rays.shape[0], rays.ctypes.data, mode, ignored_primitive, hits.ctypes.data)
        assert rc == 0
        return hits

    def render_parity(self, w, h, spp, rect=None, frame_count=0, salt=0, film=None, records=False):
        """Single-threaded per-pixel-seeded render with the reference's own render_tile. Returns (film, records)."""
        x0, y0, x1, y1 = rect if rect else (0, 0, w, h)
        if film is None:
            film = np.zeros((h, w, 4), dtype=np.float32)
        rec = None
        if records:
            rec = np.zeros(((y1 - y0) * (x1 - x0) * spp,), dtype=capi.RECORD_DTYPE)
        rc = self.lib.ref_render_parity(self.handle, film.ctypes.data, w, h, x0, y0, x1, y1, frame_count, spp, salt,
                                        rec.ctypes.data if rec is not None else None)
        if rc != 0:
            raise RuntimeError(f"ref_render_parity failed: {rc}")
        return film, rec

    def render_threaded(self, w, h, spp, threads, want_film=True):
        """The reference's verbatim WorkQueue renderer. Returns (film, seconds, stats)."""
        film = np.zeros((h, w, 4), dtype=np.float32) if want_film else None
        sec = C.c_double()
        st = capi.Stats()
        rc = self.lib.ref_render_threaded(self.handle, w, h, spp, threads,
                                          film.ctypes.data if film is not None else None, C.byref(sec), C.byref(st))
        if rc != 0:
            raise RuntimeError(f"ref_render_threaded failed: {rc}")
        return film, sec.value, st


def get_stats(reset=False):
    st = capi.Stats()
    lib().ref_get_stats(C.byref(st), int(reset))
    return st


def sampler_tables():
    perm = np.zeros(256 * 64, np.uint8)
    sobol = np.zeros(256 * 256, np.uint8)
    scr = np.zeros(128 * 128 * 8, np.uint8)
    rank = np.zeros(128 * 128 * 8, np.uint8)
    rc = lib().ref_get_sampler_tables(perm.ctypes.data, sobol.ctypes.data, scr.ctypes.data, rank.ctypes.data)
    assert rc == 0, "table values outside 0..255"
    return perm, sobol, scr, rank


def sizeof(name):
    return lib().ref_sizeof(name.encode())


def resolve_bgra8(film, exposure=0.0, tonemapping=True, srgb_transform=True, midpoint=0.5, contrast=0.0, dither=None):
    """the reference's display-loop resolve (raytracer.cpp:2103-2173, its own text compiled into the oracle): film (h, w, 4)
    float32 -> (h, w) uint32 0xAARRGGBB.  The reference always dithers: `dither` is a power-of-two (dh, dw, 3) uint8 tile."""
    L = lib()
    L.ref_resolve_bgra8.restype = C.c_int
    L.ref_resolve_bgra8.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(capi.PostSettings), C.c_void_p, C.c_uint32,
                                    C.c_uint32, C.c_void_p]
    film = np.ascontiguousarray(film, np.float32)
    h, w, _ = film.shape
    dither = np.ascontiguousarray(dither, np.uint8)
    dh, dw, _ = dither.shape
    post = capi.PostSettings(exposure, int(tonemapping), int(srgb_transform), midpoint, contrast)
    out = np.empty((h, w), np.uint32)
    rc = L.ref_resolve_bgra8(film.ctypes.data, w, h, C.byref(post), dither.ctypes.data, dw, dh, out.ctypes.data)
    assert rc == 0
    return out


def write_bitmap(path, pixels):
    """the reference's write_bitmap (assets.cpp:693-724) on a (h, w) uint32 array"""
    L = lib()
    L.ref_write_bitmap.restype = C.c_int
    L.ref_write_bitmap.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
    px = np.ascontiguousarray(pixels, np.uint32)
    h, w = px.shape
    assert L.ref_write_bitmap(path.encode(), px.ctypes.data, w, h) == 0


def parse_obj(text, winding=1):
    """the reference's parse_obj (assets.cpp:187-400): (positions, normals|None, texcoords|None) as (n, 9) arrays, or None on a parse error"""
    L = lib()
    L.ref_parse_obj.restype = C.c_int
    L.ref_parse_obj.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.c_void_p, C.c_void_p, C.c_void_p]
    raw = text if isinstance(text, bytes) else text.encode()
    n, hn, ht = C.c_uint32(), C.c_int(), C.c_int()
    if L.ref_parse_obj(raw, winding, C.byref(n), C.byref(hn), C.byref(ht), None, None, None) != 0:
        return None
    pos = np.zeros((n.value, 9), np.float32)
    nrm = np.zeros((n.value, 9), np.float32) if hn.value else None
    tex = np.zeros((n.value, 9), np.float32) if ht.value else None
    L.ref_parse_obj(raw, winding, C.byref(n), C.byref(hn), C.byref(ht), pos.ctypes.data,
                    nrm.ctypes.data if nrm is not None else None, tex.ctypes.data if tex is not None else None)
    return pos, nrm, tex


def parse_hdr(data):
    """the reference's parse_hdr (assets.cpp:423-600): (h, w, 3) float32 array in the reference's row order, or None"""
    L = lib()
    L.ref_parse_hdr.restype = C.c_int
    L.ref_parse_hdr.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p]
    w, h = C.c_uint32(), C.c_uint32()
    if L.ref_parse_hdr(data, len(data), C.byref(w), C.byref(h), None) != 0:
        return None
    px = np.zeros((h.value, w.value, 3), np.float32)
    L.ref_parse_hdr(data, len(data), C.byref(w), C.byref(h), px.ctypes.data)
    return px
