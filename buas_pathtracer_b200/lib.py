"""Loader + object wrappers for libbpt.so (the product library)."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# every symbol include/bpt.h declares (checked by tests/test_abi.py against the header itself)
MULTI_GPU_SYMBOLS = ["nccl_get_unique_id", "nccl_comm_init_rank", "nccl_comm_init_all", "nccl_comm_destroy", "reduce_film",
                     "reduced_film_device_ptr", "download_reduced_film"]
DEVICE_SYMBOLS = [
    "create", "destroy", "set_sampler_tables", "upload_scene", "update_settings", "film_resize", "film_clear",
    "film_use_external", "film_device_ptr", "download_film", "render_pass", "render_pass_bands", "sync", "trace", "set_sample_records",
    "stats_enable", "get_stats", "get_pass_timing", "set_detailed_timing", "set_tail_threshold", "get_transfer_bytes", "resolve_bgra8",
    "build_mesh_bvh_device",
]
MISC_SYMBOLS = ["last_error", "version", "make_displaced_icosphere", "make_procedural_skydome",
                "parse_obj", "load_obj", "obj_free", "obj_triangle_count", "obj_positions", "obj_normals", "obj_texcoords",
                "create_mesh_from_obj", "parse_hdr", "load_skydome_hdr"]


class BptError(RuntimeError):
    pass


def library_path():
    return os.environ.get("BPT_LIBRARY") or os.path.join(HERE, "libbpt.so")


def build_library(force=False):
    """Compile libbpt.so in-tree with nvcc for sm_100a (works without a GPU)."""
    args = ["make", "-C", os.path.join(HERE, "csrc")] + (["-B"] if force else [])
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise BptError("building libbpt.so failed:\n" + r.stdout)
    return library_path()


def load_library():
    """dlopen libbpt.so.  Fails loudly when it is missing: there is no fallback implementation."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise BptError(f"{path} not found: build it with `make -C buas_pathtracer_b200/csrc` "
                       "(or __graft_entry__.build()); there is no CPU/Python fallback")
    L = C.CDLL(path)
    vp, P = C.c_void_p, C.POINTER
    L.bpt_last_error.restype = C.c_char_p
    L.bpt_version.restype = C.c_char_p
    L.bpt_make_displaced_icosphere.restype = C.c_uint32
    L.bpt_make_displaced_icosphere.argtypes = [C.c_uint32, C.c_float, vp]
    L.bpt_make_procedural_skydome.restype = C.c_int
    L.bpt_make_procedural_skydome.argtypes = [C.c_uint32, C.c_uint32, vp]
    L.bpt_create.restype = C.c_int
    L.bpt_create.argtypes = [C.c_int, P(vp)]
    L.bpt_destroy.restype = None
    L.bpt_destroy.argtypes = [vp]
    L.bpt_set_sampler_tables.restype = C.c_int
    L.bpt_set_sampler_tables.argtypes = [vp, vp, vp, vp, vp]
    for name in ("upload_scene", "update_settings"):
        f = getattr(L, "bpt_" + name)
        f.restype = C.c_int
        f.argtypes = [vp, vp]
    L.bpt_film_resize.restype = C.c_int
    L.bpt_film_resize.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.bpt_film_clear.restype = C.c_int
    L.bpt_film_clear.argtypes = [vp]
    L.bpt_film_use_external.restype = C.c_int
    L.bpt_film_use_external.argtypes = [vp, vp, C.c_uint32, C.c_uint32]
    L.bpt_film_device_ptr.restype = C.c_int
    L.bpt_film_device_ptr.argtypes = [vp, P(vp)]
    L.bpt_download_film.restype = C.c_int
    L.bpt_download_film.argtypes = [vp, vp]
    L.bpt_render_pass.restype = C.c_int
    L.bpt_render_pass.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.bpt_render_pass_bands.restype = C.c_int
    L.bpt_render_pass_bands.argtypes = [vp, C.c_int32, C.c_int32, C.c_uint32, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.bpt_sync.restype = C.c_int
    L.bpt_sync.argtypes = [vp]
    L.bpt_trace.restype = C.c_int
    L.bpt_trace.argtypes = [vp, C.c_uint32, vp, C.c_int, C.c_uint32, vp]
    L.bpt_set_sample_records.restype = C.c_int
    L.bpt_set_sample_records.argtypes = [vp, vp, C.c_uint64]
    L.bpt_stats_enable.restype = C.c_int
    L.bpt_stats_enable.argtypes = [vp, C.c_int]
    L.bpt_get_stats.restype = C.c_int
    L.bpt_get_stats.argtypes = [vp, P(capi.Stats), C.c_int]
    L.bpt_set_ray_prefilter.restype = C.c_int
    L.bpt_set_ray_prefilter.argtypes = [vp, C.c_int]
    L.bpt_get_ray_prefilter_stats.restype = C.c_int
    L.bpt_get_ray_prefilter_stats.argtypes = [vp, P(C.c_uint64)]
    L.bpt_get_pass_timing.restype = C.c_int
    L.bpt_get_pass_timing.argtypes = [vp, P(capi.PassTiming)]
    L.bpt_set_detailed_timing.restype = C.c_int
    L.bpt_set_detailed_timing.argtypes = [vp, C.c_int]
    L.bpt_set_tail_threshold.restype = C.c_int
    L.bpt_set_tail_threshold.argtypes = [vp, C.c_uint32]
    L.bpt_build_mesh_bvh_device.restype = C.c_int
    L.bpt_build_mesh_bvh_device.argtypes = [vp, C.c_uint32, vp, C.c_int32, vp, C.c_uint32, P(C.c_uint32), vp, P(C.c_float)]
    L.bpt_get_transfer_bytes.restype = C.c_int
    L.bpt_get_transfer_bytes.argtypes = [vp, P(C.c_uint64), P(C.c_uint64), C.c_int]
    L.bpt_resolve_bgra8.restype = C.c_int
    L.bpt_resolve_bgra8.argtypes = [vp, P(capi.PostSettings), vp, C.c_uint32, C.c_uint32, vp]
    L.bpt_write_bitmap.restype = C.c_int
    L.bpt_write_bitmap.argtypes = [C.c_char_p, vp, C.c_uint32, C.c_uint32]
    if hasattr(L, "bpt_upload_scene_async"):
        L.bpt_upload_scene_async.restype = C.c_int
        L.bpt_upload_scene_async.argtypes = [vp, vp]
        L.bpt_download_film_async.restype = C.c_int
        L.bpt_download_film_async.argtypes = [vp, vp, C.c_int]
        L.bpt_wait_download.restype = C.c_int
        L.bpt_wait_download.argtypes = [vp]
        L.bpt_host_register.restype = C.c_int
        L.bpt_host_register.argtypes = [vp, C.c_size_t]
        L.bpt_host_unregister.restype = C.c_int
        L.bpt_host_unregister.argtypes = [vp]
    if hasattr(L, "bpt_reduce_film"):
        L.bpt_nccl_get_unique_id.restype = C.c_int
        L.bpt_nccl_get_unique_id.argtypes = [vp]
        L.bpt_nccl_comm_init_rank.restype = C.c_int
        L.bpt_nccl_comm_init_rank.argtypes = [vp, vp, C.c_int, C.c_int, P(vp)]
        L.bpt_nccl_comm_init_all.restype = C.c_int
        L.bpt_nccl_comm_init_all.argtypes = [C.c_int, vp, P(vp)]
        L.bpt_nccl_comm_destroy.restype = C.c_int
        L.bpt_nccl_comm_destroy.argtypes = [vp]
        L.bpt_reduce_film.restype = C.c_int
        L.bpt_reduce_film.argtypes = [vp, vp, C.c_int]
        L.bpt_reduced_film_device_ptr.restype = C.c_int
        L.bpt_reduced_film_device_ptr.argtypes = [vp, P(vp)]
        L.bpt_download_reduced_film.restype = C.c_int
        L.bpt_download_reduced_film.argtypes = [vp, vp]
    if hasattr(L, "bpt_get_wide_bvh"):        # absent from older builds selected with BPT_LIBRARY for A/B runs
        L.bpt_get_wide_bvh.restype = C.c_int
        L.bpt_get_wide_bvh.argtypes = [vp, C.c_int32, P(vp), P(C.c_uint32), vp, P(vp), P(C.c_uint32), P(C.c_uint32)]
    _LIB = L
    return L


def _check(rc, what):
    if rc != 0:
        msg = load_library().bpt_last_error().decode(errors="replace")
        raise BptError(f"{what} failed ({rc}): {msg}")


def sampler_tables():
    """The four sampler lookup tables as uint8 arrays (data file written by oracle/tools/dump_sampler_tables.py)."""
    blob = np.fromfile(os.path.join(HERE, "data", "sampler_tables.bin"), dtype=np.uint8)
    assert blob.size == 16384 + 65536 + 131072 + 131072
    return blob[:16384], blob[16384:81920], blob[81920:212992], blob[212992:]


class Scene(capi.HostScene):
    """Host scene of the product library (mirrors Raytracer/scene.h:134-149)."""

    def __init__(self):
        super().__init__(load_library(), "bpt_")

    def wide_bvh(self, mesh=-1):
        """The device layout of a BVH (csrc/wide_bvh.h) as (children[2*pairs] structured array, root child, big leaves
        [n,2], depth); mesh < 0 = the TLAS.  Copies."""
        L = self.lib
        pairs, big = C.c_void_p(), C.c_void_p()
        npairs, nbig, depth = C.c_uint32(), C.c_uint32(), C.c_uint32()
        root = np.zeros(1, capi.WIDE_CHILD_DTYPE)
        _check(L.bpt_get_wide_bvh(self.handle, int(mesh), C.byref(pairs), C.byref(npairs), root.ctypes.data, C.byref(big),
                                  C.byref(nbig), C.byref(depth)), "bpt_get_wide_bvh")
        children = np.zeros(2 * npairs.value, capi.WIDE_CHILD_DTYPE)
        if npairs.value:
            C.memmove(children.ctypes.data, pairs.value, children.nbytes)
        bl = np.zeros((nbig.value, 2), np.uint32)
        if nbig.value:
            C.memmove(bl.ctypes.data, big.value, bl.nbytes)
        return children, root[0], bl, depth.value

    def create_mesh_with_bvh(self, positions, nodes, indices, normals=None):
        """bpt_create_mesh with a caller-supplied BVH (e.g. Renderer.build_mesh_bvh) instead of the host build"""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 9)
        nodes = np.ascontiguousarray(nodes, dtype=capi.BVH_NODE_DTYPE)
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        nrm = None if normals is None else np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 9)
        f = self.lib.bpt_create_mesh_with_bvh
        f.restype = C.c_uint32
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        h = f(self.handle, pos.shape[0], pos.ctypes.data, None if nrm is None else nrm.ctypes.data,
              nodes.ctypes.data, nodes.shape[0], idx.ctypes.data)
        if h == 0xFFFFFFFF:
            raise BptError(self.lib.bpt_last_error().decode())
        return h


def make_displaced_icosphere(level, amplitude=0.08):
    L = load_library()
    n = L.bpt_make_displaced_icosphere(level, amplitude, None)
    if n == 0:
        raise BptError(L.bpt_last_error().decode())
    tris = np.zeros((n, 9), np.float32)
    L.bpt_make_displaced_icosphere(level, amplitude, tris.ctypes.data)
    return tris


def parse_obj(text, winding=1, path=None):
    """bpt_parse_obj / bpt_load_obj: (positions, normals|None, texcoords|None) as (n, 9) float32 arrays; raises BptError"""
    L = load_library()
    fp = C.POINTER(C.c_float)
    for name in ("bpt_parse_obj", "bpt_load_obj"):
        getattr(L, name).restype = C.c_void_p
        getattr(L, name).argtypes = [C.c_char_p, C.c_int32]
    L.bpt_obj_free.argtypes = [C.c_void_p]
    L.bpt_obj_triangle_count.restype = C.c_uint32
    L.bpt_obj_triangle_count.argtypes = [C.c_void_p]
    for name in ("bpt_obj_positions", "bpt_obj_normals", "bpt_obj_texcoords"):
        getattr(L, name).restype = fp
        getattr(L, name).argtypes = [C.c_void_p]
    if path is not None:
        h = L.bpt_load_obj(path.encode(), winding)
    else:
        h = L.bpt_parse_obj(text if isinstance(text, bytes) else text.encode(), winding)
    if not h:
        raise BptError(L.bpt_last_error().decode())
    n = L.bpt_obj_triangle_count(h)

    def grab(f):
        p = f(h)
        return np.ctypeslib.as_array(p, shape=(n, 9)).copy() if p else None
    out = (grab(L.bpt_obj_positions) if n else np.zeros((0, 9), np.float32), grab(L.bpt_obj_normals), grab(L.bpt_obj_texcoords))
    L.bpt_obj_free(h)
    return out


def parse_hdr(data):
    """bpt_parse_hdr: (h, w, 3) float32 array in the reference's row order; raises BptError"""
    L = load_library()
    L.bpt_parse_hdr.restype = C.c_int
    L.bpt_parse_hdr.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p]
    w, h = C.c_uint32(), C.c_uint32()
    _check(L.bpt_parse_hdr(data, len(data), C.byref(w), C.byref(h), None), "bpt_parse_hdr")
    px = np.zeros((h.value, w.value, 3), np.float32)
    _check(L.bpt_parse_hdr(data, len(data), C.byref(w), C.byref(h), px.ctypes.data), "bpt_parse_hdr")
    return px


def make_procedural_skydome(w=2048, h=1024):
    L = load_library()
    px = np.zeros((h, w, 3), np.float32)
    _check(L.bpt_make_procedural_skydome(w, h, px.ctypes.data), "bpt_make_procedural_skydome")
    return px


def nccl_comm_init_all(devices):
    """ncclCommInitAll for one process driving several GPUs: a list of ncclComm_t handles, one per device"""
    L = load_library()
    devs = (C.c_int * len(devices))(*devices)
    comms = (C.c_void_p * len(devices))()
    _check(L.bpt_nccl_comm_init_all(len(devices), devs, comms), "bpt_nccl_comm_init_all")
    return [C.c_void_p(c) for c in comms]


class Renderer:
    """One GPU context (one per process/GPU).  Raises BptError when no CUDA device is usable."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib.bpt_create(device, C.byref(h)), "bpt_create")
        self.handle = h
        self.device = device
        self.w = self.h = 0
        self._records = None
        perm, sobol, scr, rank = sampler_tables()
        _check(self.lib.bpt_set_sampler_tables(self.handle, perm.ctypes.data, sobol.ctypes.data, scr.ctypes.data,
                                               rank.ctypes.data), "bpt_set_sampler_tables")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.bpt_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_scene(self, scene):
        _check(self.lib.bpt_upload_scene(self.handle, scene.handle), "bpt_upload_scene")

    def upload_scene_async(self, scene):
        """bpt_upload_scene_async: into the inactive scene buffer on a copy stream; the next pass switches to it"""
        _check(self.lib.bpt_upload_scene_async(self.handle, scene.handle), "bpt_upload_scene_async")

    def download_film_async(self, out, reduced=False):
        """snapshot the film behind the passes enqueued so far and start its copy to `out` (page-locked, see host_register)"""
        _check(self.lib.bpt_download_film_async(self.handle, out.ctypes.data, int(reduced)), "bpt_download_film_async")

    def wait_download(self):
        _check(self.lib.bpt_wait_download(self.handle), "bpt_wait_download")

    def host_register(self, array):
        _check(self.lib.bpt_host_register(array.ctypes.data, array.nbytes), "bpt_host_register")

    def host_unregister(self, array):
        _check(self.lib.bpt_host_unregister(array.ctypes.data), "bpt_host_unregister")

    def update_settings(self, scene):
        _check(self.lib.bpt_update_settings(self.handle, scene.handle), "bpt_update_settings")

    def film_resize(self, w, h):
        _check(self.lib.bpt_film_resize(self.handle, w, h), "bpt_film_resize")
        self.w, self.h = w, h

    def film_use_external(self, device_ptr, w, h):
        _check(self.lib.bpt_film_use_external(self.handle, C.c_void_p(device_ptr), w, h), "bpt_film_use_external")
        self.w, self.h = w, h

    def film_clear(self):
        _check(self.lib.bpt_film_clear(self.handle), "bpt_film_clear")

    def render_pass(self, spp, rect=None, frame_count=0, salt=0):
        x0, y0, x1, y1 = rect if rect else (0, 0, self.w, self.h)
        _check(self.lib.bpt_render_pass(self.handle, x0, y0, x1, y1, frame_count, spp, 0, salt), "bpt_render_pass")

    def render_pass_bands(self, spp, bands, x0=0, x1=None, frame_count=0, salt=0):
        """one pass over a list of row bands [(y0, y1), ...] as a single workload (a rank's share of the frame)"""
        arr = np.ascontiguousarray(np.array(bands, dtype=np.int32).reshape(-1, 2))
        _check(self.lib.bpt_render_pass_bands(self.handle, x0, self.w if x1 is None else x1, arr.shape[0], arr.ctypes.data,
                                              frame_count, spp, 0, salt), "bpt_render_pass_bands")

    def sync(self):
        _check(self.lib.bpt_sync(self.handle), "bpt_sync")

    def download_film(self, out=None):
        if out is None:
            out = np.empty((self.h, self.w, 4), np.float32)
        _check(self.lib.bpt_download_film(self.handle, out.ctypes.data), "bpt_download_film")
        return out

    # -- multi-GPU (include/bpt.h section 3): one Renderer per rank, one NCCL reduce per progressive pass --
    @staticmethod
    def nccl_unique_id():
        """ncclGetUniqueId as 128 bytes (rank 0 creates it and hands it to the other ranks)"""
        buf = np.zeros(128, np.uint8)
        _check(load_library().bpt_nccl_get_unique_id(buf.ctypes.data), "bpt_nccl_get_unique_id")
        return buf

    def nccl_comm_init_rank(self, unique_id, nranks, rank):
        uid = np.ascontiguousarray(unique_id, np.uint8)
        assert uid.size == 128
        comm = C.c_void_p()
        _check(self.lib.bpt_nccl_comm_init_rank(self.handle, uid.ctypes.data, nranks, rank, C.byref(comm)), "bpt_nccl_comm_init_rank")
        return comm

    def nccl_comm_destroy(self, comm):
        _check(self.lib.bpt_nccl_comm_destroy(comm), "bpt_nccl_comm_destroy")

    def reduce_film(self, comm, root=0):
        """root's reduced film := sum of all ranks' films (asynchronous on the context's stream; partial films untouched)"""
        _check(self.lib.bpt_reduce_film(self.handle, comm, root), "bpt_reduce_film")

    def download_reduced_film(self, out=None):
        if out is None:
            out = np.empty((self.h, self.w, 4), np.float32)
        _check(self.lib.bpt_download_reduced_film(self.handle, out.ctypes.data), "bpt_download_reduced_film")
        return out

    def trace(self, rays, mode=capi.TRACE_CLOSEST, ignored_primitive=0):
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=capi.HIT_DTYPE)
        _check(self.lib.bpt_trace(self.handle, rays.shape[0], rays.ctypes.data, mode, ignored_primitive,
                                  hits.ctypes.data), "bpt_trace")
        return hits

    def attach_records(self, count):
        if count:
            self._records = np.zeros(count, dtype=capi.RECORD_DTYPE)
            _check(self.lib.bpt_set_sample_records(self.handle, self._records.ctypes.data, count), "bpt_set_sample_records")
        else:
            self._records = None
            _check(self.lib.bpt_set_sample_records(self.handle, None, 0), "bpt_set_sample_records")
        return self._records

    def stats_enable(self, on=True):
        _check(self.lib.bpt_stats_enable(self.handle, int(on)), "bpt_stats_enable")

    def get_stats(self, reset=False):
        st = capi.Stats()
        _check(self.lib.bpt_get_stats(self.handle, C.byref(st), int(reset)), "bpt_get_stats")
        return st

    def set_ray_prefilter(self, on=True):
        _check(self.lib.bpt_set_ray_prefilter(self.handle, int(on)), "bpt_set_ray_prefilter")

    def ray_prefilter_stats(self):
        """(rays, algorithmic bytes) of the shadow rays the shading kernel settles itself, as counted by the last pass(es) rendered
        with stats_enable(True); read before get_stats(reset=True)"""
        out = (C.c_uint64 * 2)()
        _check(self.lib.bpt_get_ray_prefilter_stats(self.handle, out), "bpt_get_ray_prefilter_stats")
        return int(out[0]), int(out[1])

    def set_detailed_timing(self, on=True):
        _check(self.lib.bpt_set_detailed_timing(self.handle, int(on)), "bpt_set_detailed_timing")

    def build_mesh_bvh(self, positions, method=capi.BVH_SAH_BINNED):
        """create_bvh_for_mesh on the device: (nodes, leaf-order indices, device milliseconds)"""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 9)
        n = pos.shape[0]
        nodes = np.zeros(2 * n + 2, dtype=capi.BVH_NODE_DTYPE)
        idx = np.zeros(n, dtype=np.uint32)
        nc, ms = C.c_uint32(), C.c_float()
        _check(self.lib.bpt_build_mesh_bvh_device(self.handle, n, pos.ctypes.data, int(method), nodes.ctypes.data, nodes.shape[0],
                                                  C.byref(nc), idx.ctypes.data, C.byref(ms)), "bpt_build_mesh_bvh_device")
        return nodes[:nc.value].copy(), idx, ms.value

    def set_tail_threshold(self, paths):
        _check(self.lib.bpt_set_tail_threshold(self.handle, int(paths)), "bpt_set_tail_threshold")

    def transfer_bytes(self, reset=False):
        a, b = C.c_uint64(), C.c_uint64()
        _check(self.lib.bpt_get_transfer_bytes(self.handle, C.byref(a), C.byref(b), int(reset)), "bpt_get_transfer_bytes")
        return a.value, b.value

    def resolve_bgra8(self, exposure=0.0, tonemapping=True, srgb_transform=True, midpoint=0.5, contrast=0.0, dither=None):
        """film -> (h, w) uint32 0xAARRGGBB as the reference's display loop produces it (raytracer.cpp:2103-2172)"""
        post = capi.PostSettings(exposure, int(tonemapping), int(srgb_transform), midpoint, contrast)
        out = np.empty((self.h, self.w), np.uint32)
        dp, dw, dh = None, 0, 0
        if dither is not None:
            dither = np.ascontiguousarray(dither, np.uint8)
            dh, dw, _ = dither.shape
            dp = dither.ctypes.data
        _check(self.lib.bpt_resolve_bgra8(self.handle, C.byref(post), dp, dw, dh, out.ctypes.data), "bpt_resolve_bgra8")
        return out

    def pass_timing(self):
        t = capi.PassTiming()
        _check(self.lib.bpt_get_pass_timing(self.handle, C.byref(t)), "bpt_get_pass_timing")
        return t
