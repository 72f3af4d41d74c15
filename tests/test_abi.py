"""CPU tests of the drop-in boundary: libbpt.so loads, exports every symbol include/bpt.h declares, keeps torch/C++
types out of its signatures, and fails loudly (no fallback) when no GPU is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bpt.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"BPT_API[^;(]*?\b(bpt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("bpt_create", "bpt_upload_scene", "bpt_render_pass", "bpt_download_film", "bpt_trace",
                 "bpt_create_scene_bvh", "bpt_add_mesh", "bpt_set_sampler_tables", "bpt_get_stats"):
        assert must in syms
    assert len(syms) >= 45


def test_library_exports_every_declared_symbol(bpt):
    lib = bpt.load_library()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/bpt.h but not exported: {missing}"


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: the header must compile as C99 with nothing but stdint/stddef"""
    src = tmp_path / "t.c"
    src.write_text('#include "bpt.h"\nint main(void) { bpt_hit h; bpt_ray r; (void)h; (void)r; return (int)sizeof(bpt_stats) == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                        "-o", str(tmp_path / "t.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_cpu_fallback_without_gpu(bpt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(bpt.BptError) as e:
        bpt.Renderer(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """nothing under the package (or the library's link line) references oracle/"""
    pkg = os.path.join(ROOT, "buas_pathtracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "ref_oracle" not in text and "libbpt_ref" not in text, f"{f} references the oracle"
    out = subprocess.run(["ldd", os.path.join(pkg, "libbpt.so")], capture_output=True, text=True).stdout
    assert "libbpt_ref" not in out


def test_host_scene_error_paths(bpt):
    s = bpt.Scene()
    L = bpt.load_library()
    assert s.api.add_mesh(s.handle, 0, 99, None) == 0xFFFFFFFF
    assert b"unknown mesh" in L.bpt_last_error()
    with pytest.raises(RuntimeError):
        s.scene_bvh()                      # BVH not built yet -> BPT_ERR_STATE
    fc = s.get_filter_cache()
    fc.kernel_size = 40
    assert s.api.set_filter_cache(s.handle, C.byref(fc)) != 0


def test_new_host_entry_points_report_errors(bpt):
    """bpt_create_mesh_ex / bpt_create_mesh_with_bvh / bpt_parse_obj / bpt_parse_hdr: bad input -> error code + message, never a crash"""
    import numpy as np
    from buas_pathtracer_b200 import capi, lib
    s = bpt.Scene()
    tri = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    with pytest.raises(RuntimeError):
        s.create_mesh(tri, method=7)                                   # unknown construction method
    assert "unknown method" in bpt.load_library().bpt_last_error().decode()
    m = s.create_mesh(np.repeat(tri, 9, axis=0), method=capi.BVH_MIDPOINT_SPLIT)
    nodes, idx, _ = s.mesh_bvh(m)
    with pytest.raises(bpt.BptError):
        s.create_mesh_with_bvh(np.repeat(tri, 9, axis=0), nodes, idx + 100)     # index outside the mesh
    assert s.create_mesh_with_bvh(np.repeat(tri, 9, axis=0), nodes, idx) == m + 1
    with pytest.raises(bpt.BptError):
        lib.parse_obj("f 1 2 3\n", 1)                                           # indices into an empty vertex pool
    with pytest.raises(bpt.BptError):
        lib.parse_obj("v 0 0 0\n", 5)                                           # winding
    with pytest.raises(bpt.BptError):
        lib.parse_hdr(b"")                                                      # nothing at all
