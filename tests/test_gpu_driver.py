"""The headless C++ driver (driver/bpt_headless.cpp: what is left of SDL_main once the window is gone) end to end on a
GPU: OBJ mesh + HDR environment from files, device-built BVH, a non-default integrator and filter, "Take picture"."""
import os
import struct
import subprocess

import numpy as np
import pytest

import test_assets

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("extra", [[], ["--gpu-bvh", "--integrator", "Whitted", "--filter", "Gaussian 3"],
                                   ["--integrator", "Normals", "--filter", "Box"]])
def test_driver_renders_obj_with_hdr_sky(tmp_path, extra):
    driver = os.path.join(ROOT, "driver", "bpt_headless")
    if not os.path.exists(driver):
        pytest.skip("driver not built")
    obj = tmp_path / "cube.obj"
    obj.write_text(test_assets.cube_obj())
    hdr = tmp_path / "sky.hdr"
    hdr.write_bytes(test_assets.make_hdr(test_assets.synthetic_rgbe(64, 32, 9), "-Y 32 +X 64"))
    out = tmp_path / "out.bmp"
    cmd = [driver, "--tables", os.path.join(ROOT, "buas_pathtracer_b200", "data", "sampler_tables.bin"), "--obj", str(obj),
           "--hdr", str(hdr), "--w", "160", "--h", "90", "--spp", "8", "--out", str(out)] + extra
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert p.returncode == 0, p.stdout
    data = out.read_bytes()
    assert data[:2] == b"BM"
    w, h = struct.unpack_from("<ii", data, 18)
    assert (w, abs(h)) == (160, 90)
    px = np.frombuffer(data, np.uint8, offset=struct.unpack_from("<I", data, 10)[0])
    assert px.std() > 5, "image is flat"
    if "--gpu-bvh" in extra:
        assert "device BVH build: 12 triangles" in p.stdout
