"""ctypes mirror of include/bpt.h.

The host-scene half of the C ABI (``<prefix>add_sphere`` ...) is exported with identical signatures by the
product library (prefix ``bpt_``) and by the test oracle (prefix ``ref_``, oracle/ref_api.h), so one binder
serves both; that is what lets the parity tests replay one scene description into each side.
"""
import ctypes as C

import numpy as np

c_float3 = C.c_float * 3


class M4x4Inv(C.Structure):
    """bpt_m4x4inv == MathLib/math_types.h:42-49 (row-major, translation in e[r][3])."""
    _fields_ = [("forward", (C.c_float * 4) * 4), ("inverse", (C.c_float * 4) * 4)]

    @staticmethod
    def from_numpy(fwd, inv):
        m = M4x4Inv()
        f = np.ascontiguousarray(fwd, dtype=np.float32)
        i = np.ascontiguousarray(inv, dtype=np.float32)
        C.memmove(m.forward, f.ctypes.data, 64)
        C.memmove(m.inverse, i.ctypes.data, 64)
        return m


class Material(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("albedo", c_float3), ("checker_color", c_float3),
                ("emission_color", c_float3), ("ior", C.c_float), ("metallic", C.c_float),
                ("roughness", C.c_float), ("is_participating_medium", C.c_int32), ("absorb", c_float3)]


class Camera(C.Structure):
    _fields_ = [("p", c_float3), ("x", c_float3), ("y", c_float3), ("z", c_float3),
                ("vfov", C.c_float), ("aspect_ratio", C.c_float), ("lens_radius", C.c_float),
                ("focus_distance", C.c_float), ("film_distance", C.c_float),
                ("half_film_w", C.c_float), ("half_film_h", C.c_float)]


class Settings(C.Structure):
    _fields_ = [("next_event_estimation", C.c_int32), ("importance_sample_lights", C.c_int32),
                ("importance_sample_diffuse", C.c_int32), ("use_mis", C.c_int32),
                ("russian_roulette", C.c_int32), ("caustics", C.c_int32),
                ("sampling_strategy", C.c_int32), ("use_path_guide", C.c_int32),
                ("vignette_strength", C.c_float), ("lens_distortion", C.c_float),
                ("f_factor", C.c_float), ("diaphragm_edges", C.c_float), ("phi_shutter_max", C.c_float),
                ("samples_per_pixel", C.c_uint32), ("max_bounce_count", C.c_uint32),
                ("integrator", C.c_int32)]


class BvhNode(C.Structure):
    _fields_ = [("bv_p", c_float3), ("bv_r", c_float3), ("left_first", C.c_uint32),
                ("count", C.c_uint16), ("split_axis", C.c_uint16)]


BVH_MIDPOINT_SPLIT, BVH_SAH_BINNED, BVH_SAH_FULL = 0, 1, 2      # BVHConstructionMethod (bvh.h:7-11)
BVH_NODE_DTYPE = np.dtype([("bv_p", np.float32, 3), ("bv_r", np.float32, 3), ("left_first", np.uint32),
                           ("count", np.uint16), ("split_axis", np.uint16)])
assert BVH_NODE_DTYPE.itemsize == 32 and C.sizeof(BvhNode) == 32
# one child of the device layout (csrc/wide_bvh.h): the node's box verbatim + a packed reference
WIDE_CHILD_DTYPE = np.dtype([("bv_p", np.float32, 3), ("bv_r", np.float32, 3), ("ref", np.uint32), ("aux", np.uint32)])
assert WIDE_CHILD_DTYPE.itemsize == 32
WREF_LEAF, WREF_RECORD_ROOT, WREF_INDEX_MASK = 0x80000000, 0x10000000, 0x0FFFFFFF


class FilterCache(C.Structure):
    _fields_ = [("kernel_size", C.c_uint32), ("cache_size", C.c_uint32), ("cache", C.c_float * 512)]


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3), ("max_t", np.float32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("primitive", np.uint32), ("triangle", np.uint32),
                      ("n", np.float32, 3), ("p", np.float32, 3)])
RECORD_DTYPE = np.dtype([("ray_o", np.float32, 3), ("ray_d", np.float32, 3), ("radiance", np.float32, 3),
                         ("rays", np.uint32)])
assert RAY_DTYPE.itemsize == 28 and HIT_DTYPE.itemsize == 36 and RECORD_DTYPE.itemsize == 40

HIT_MISS = 0xFFFFFFFF
HIT_PLANE = 0x80000000
TRACE_CLOSEST, TRACE_OCCLUSION = 0, 1
SAMPLING_UNIFORM, SAMPLING_BLUE_NOISE, SAMPLING_STRATIFIED = 0, 1, 2


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "rays", "shadow_rays", "tlas_node_pops", "instances_visited", "mesh_intersection_count",
        "mesh_bvh_traversals", "mesh_node_traversals", "mesh_leaf_traversals", "triangles_tested", "samples",
        "shadow_tlas_node_pops", "shadow_instances_visited", "shadow_mesh_intersection_count",
        "shadow_mesh_bvh_traversals", "shadow_mesh_node_traversals", "shadow_mesh_leaf_traversals",
        "shadow_triangles_tested")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class PassTiming(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("raygen_ms", C.c_float), ("trace_ms", C.c_float),
                ("shade_ms", C.c_float), ("shadow_ms", C.c_float), ("splat_ms", C.c_float),
                ("kernel_launches", C.c_uint32), ("trace_launches", C.c_uint32)]


class PostSettings(C.Structure):
    """bpt_post_settings == PostProcessSettings (Raytracer/scene.h:84-90)"""
    _fields_ = [("exposure", C.c_float), ("tonemapping", C.c_int32), ("srgb_transform", C.c_int32),
                ("midpoint", C.c_float), ("contrast", C.c_float)]


def _f3(v):
    return c_float3(*[float(x) for x in v])


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


HOST_SCENE_SYMBOLS = [
    "scene_create", "scene_destroy", "add_material", "add_diffuse_material", "add_translucent_material",
    "add_emissive_material", "add_plane", "add_sphere", "add_box", "create_mesh", "create_mesh_ex", "add_mesh", "set_sky", "set_ambient_light",
    "set_skydome", "get_camera", "set_camera", "aim_camera", "aim_camera_at", "get_settings", "set_settings",
    "find_integrator", "load_reconstruction_kernel", "get_filter_cache", "set_filter_cache",
    "create_scene_bvh", "get_scene_bvh", "get_mesh_bvh", "get_counts",
]


def bind_host_scene(lib, prefix):
    """Declare argtypes/restype for the host-scene functions of `lib` and return them as a namespace."""
    P = C.POINTER
    vp = C.c_void_p
    sig = {
        "scene_create": (vp, []),
        "scene_destroy": (None, [vp]),
        "add_material": (C.c_uint32, [vp, P(Material)]),
        "add_diffuse_material": (C.c_uint32, [vp, c_float3, C.c_float, C.c_float, C.c_int32, c_float3]),
        "add_translucent_material": (C.c_uint32, [vp, c_float3, C.c_float, C.c_float]),
        "add_emissive_material": (C.c_uint32, [vp, c_float3]),
        "add_plane": (C.c_uint32, [vp, C.c_uint32, c_float3, C.c_float]),
        "add_sphere": (C.c_uint32, [vp, C.c_uint32, C.c_float, P(M4x4Inv)]),
        "add_box": (C.c_uint32, [vp, C.c_uint32, c_float3, P(M4x4Inv)]),
        "create_mesh": (C.c_uint32, [vp, C.c_uint32, P(C.c_float), P(C.c_float)]),
        "create_mesh_ex": (C.c_uint32, [vp, C.c_uint32, P(C.c_float), P(C.c_float), C.c_int32]),
        "add_mesh": (C.c_uint32, [vp, C.c_uint32, C.c_uint32, P(M4x4Inv)]),
        "set_sky": (C.c_int, [vp, c_float3, c_float3]),
        "set_ambient_light": (C.c_int, [vp, c_float3]),
        "set_skydome": (C.c_int, [vp, C.c_uint32, C.c_uint32, P(C.c_float)]),
        "get_camera": (C.c_int, [vp, P(Camera)]),
        "set_camera": (C.c_int, [vp, P(Camera)]),
        "aim_camera": (C.c_int, [vp, c_float3]),
        "aim_camera_at": (C.c_int, [vp, c_float3]),
        "get_settings": (C.c_int, [vp, P(Settings)]),
        "set_settings": (C.c_int, [vp, P(Settings)]),
        "find_integrator": (C.c_int, [C.c_char_p]),
        "load_reconstruction_kernel": (C.c_int, [vp, C.c_char_p]),
        "get_filter_cache": (C.c_int, [vp, P(FilterCache)]),
        "set_filter_cache": (C.c_int, [vp, P(FilterCache)]),
        "create_scene_bvh": (C.c_int, [vp]),
        "get_scene_bvh": (C.c_int, [vp, P(P(BvhNode)), P(C.c_uint32), P(P(C.c_uint32)), P(C.c_uint32)]),
        "get_mesh_bvh": (C.c_int, [vp, C.c_uint32, P(P(BvhNode)), P(C.c_uint32), P(P(C.c_uint32)),
                                   P(C.c_uint32), P(P(C.c_float))]),
        "get_counts": (C.c_int, [vp] + [P(C.c_uint32)] * 5),
    }

    class NS:
        pass

    ns = NS()
    for name, (res, args) in sig.items():
        fn = getattr(lib, prefix + name)
        fn.restype = res
        fn.argtypes = args
        setattr(ns, name, fn)
    return ns


class HostScene:
    """Thin object wrapper over the host-scene C ABI of either backend (``bpt_`` or ``ref_``).

    Method names and argument meaning follow Raytracer/scene.h:134-149.
    """

    def __init__(self, lib, prefix):
        self.lib = lib
        self.prefix = prefix
        self.api = bind_host_scene(lib, prefix)
        self.handle = self.api.scene_create()
        if not self.handle:
            raise RuntimeError("scene_create failed")
        self._keep = []

    def close(self):
        if self.handle:
            self.api.scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- materials (scene.cpp:9-62) --
    def add_material(self, **kw):
        m = Material()
        for k, v in kw.items():
            if k in ("albedo", "checker_color", "emission_color", "absorb"):
                setattr(m, k, _f3(v))
            else:
                setattr(m, k, v)
        return self.api.add_material(self.handle, C.byref(m))

    def add_diffuse_material(self, color, ior, roughness=0.0, checkers=False, checker_color=(0.1, 0.1, 0.1)):
        return self.api.add_diffuse_material(self.handle, _f3(color), ior, roughness, int(checkers), _f3(checker_color))

    def add_translucent_material(self, absorb, ior, roughness=0.0):
        return self.api.add_translucent_material(self.handle, _f3(absorb), ior, roughness)

    def add_emissive_material(self, emission):
        return self.api.add_emissive_material(self.handle, _f3(emission))

    # -- primitives (scene.cpp:107-159) --
    @staticmethod
    def _xf(transform):
        if transform is None:
            return None
        if isinstance(transform, M4x4Inv):
            return C.byref(transform)
        fwd, inv = transform
        return C.byref(M4x4Inv.from_numpy(fwd, inv))

    def add_plane(self, material, n, d):
        return self.api.add_plane(self.handle, material, _f3(n), d)

    def add_sphere(self, material, r, transform=None):
        return self.api.add_sphere(self.handle, material, r, self._xf(transform))

    def add_box(self, material, r, transform=None):
        return self.api.add_box(self.handle, material, _f3(r), self._xf(transform))

    def create_mesh(self, positions, normals=None, method=None):
        """method: None = the default build (SAH binned), or BVH_MIDPOINT_SPLIT / BVH_SAH_BINNED / BVH_SAH_FULL (bvh.h:7-11)"""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 9)
        nrm = None
        if normals is not None:
            nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 9)
            assert nrm.shape == pos.shape
        if method is None:
            h = self.api.create_mesh(self.handle, pos.shape[0], _fptr(pos), _fptr(nrm) if nrm is not None else None)
        else:
            h = self.api.create_mesh_ex(self.handle, pos.shape[0], _fptr(pos), _fptr(nrm) if nrm is not None else None, int(method))
        if h == 0xFFFFFFFF:
            raise RuntimeError("create_mesh failed")
        return h

    def add_mesh(self, material, mesh, transform=None):
        return self.api.add_mesh(self.handle, material, mesh, self._xf(transform))

    # -- environment --
    def set_sky(self, top, bot):
        return self.api.set_sky(self.handle, _f3(top), _f3(bot))

    def set_ambient_light(self, rgb):
        return self.api.set_ambient_light(self.handle, _f3(rgb))

    def set_skydome(self, pixels):
        px = np.ascontiguousarray(pixels, dtype=np.float32)
        h, w, _ = px.shape
        return self.api.set_skydome(self.handle, w, h, _fptr(px))

    # -- camera / settings --
    def get_camera(self):
        c = Camera()
        self.api.get_camera(self.handle, C.byref(c))
        return c

    def set_camera(self, cam):
        return self.api.set_camera(self.handle, C.byref(cam))

    def aim_camera(self, d):
        return self.api.aim_camera(self.handle, _f3(d))

    def aim_camera_at(self, at):
        return self.api.aim_camera_at(self.handle, _f3(at))

    def get_settings(self):
        s = Settings()
        self.api.get_settings(self.handle, C.byref(s))
        return s

    def set_settings(self, s):
        return self.api.set_settings(self.handle, C.byref(s))

    def update_settings(self, **kw):
        s = self.get_settings()
        for k, v in kw.items():
            if k == "integrator" and isinstance(v, str):
                v = self.api.find_integrator(v.encode())
            setattr(s, k, v)
        self.set_settings(s)
        return s

    def load_reconstruction_kernel(self, name):
        return self.api.load_reconstruction_kernel(self.handle, name.encode())

    def get_filter_cache(self):
        f = FilterCache()
        self.api.get_filter_cache(self.handle, C.byref(f))
        return f

    # -- BVH --
    def create_scene_bvh(self):
        rc = self.api.create_scene_bvh(self.handle)
        if rc != 0:
            raise RuntimeError(f"create_scene_bvh failed: {rc}")

    def scene_bvh(self):
        """(nodes as structured array, indices) -- copies."""
        nodes = C.POINTER(BvhNode)()
        idx = C.POINTER(C.c_uint32)()
        nc, ic = C.c_uint32(), C.c_uint32()
        rc = self.api.get_scene_bvh(self.handle, C.byref(nodes), C.byref(nc), C.byref(idx), C.byref(ic))
        if rc != 0:
            raise RuntimeError(f"get_scene_bvh failed: {rc}")
        n = np.ctypeslib.as_array(C.cast(nodes, C.POINTER(C.c_uint8)), shape=(nc.value * 32,)).copy().view(BVH_NODE_DTYPE)
        i = np.ctypeslib.as_array(idx, shape=(ic.value,)).copy()
        return n, i

    def mesh_bvh(self, mesh):
        nodes = C.POINTER(BvhNode)()
        idx = C.POINTER(C.c_uint32)()
        tris = C.POINTER(C.c_float)()
        nc, ic = C.c_uint32(), C.c_uint32()
        rc = self.api.get_mesh_bvh(self.handle, mesh, C.byref(nodes), C.byref(nc), C.byref(idx), C.byref(ic), C.byref(tris))
        if rc != 0:
            raise RuntimeError(f"get_mesh_bvh failed: {rc}")
        n = np.ctypeslib.as_array(C.cast(nodes, C.POINTER(C.c_uint8)), shape=(nc.value * 32,)).copy().view(BVH_NODE_DTYPE)
        i = np.ctypeslib.as_array(idx, shape=(ic.value,)).copy()
        t = np.ctypeslib.as_array(tris, shape=(ic.value, 9)).copy()
        return n, i, t

    def counts(self):
        v = [C.c_uint32() for _ in range(5)]
        self.api.get_counts(self.handle, *[C.byref(x) for x in v])
        return dict(zip(("materials", "primitives", "planes", "lights", "meshes"), [x.value for x in v]))
