"""OBJ and Radiance HDR readers (CPU): bpt_parse_obj / bpt_parse_hdr against the reference's own
parse_obj / parse_hdr (Raytracer/assets.cpp) on the same bytes.  The reference ships no asset small enough to commit
(its .obj/.hdr files are LFS blobs missing from the mount), so the inputs are generated here, quirks included."""
import struct

import numpy as np
import pytest

from buas_pathtracer_b200 import lib
from helpers import bits


def cube_obj(newline="\n", indent=""):
    v = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]
    vn = [(0, 0, -1), (0, 0, 1), (0, -1, 0), (0, 1, 0), (-1, 0, 0), (1, 0, 0)]
    vt = [(0, 0), (1, 0), (1, 1), (0, 1)]
    quads = [((1, 2, 3, 4), 1), ((5, 8, 7, 6), 2), ((1, 5, 6, 2), 3), ((4, 3, 7, 8), 4), ((1, 4, 8, 5), 5), ((2, 6, 7, 3), 6)]
    out = ["# a cube with quads, texture coordinates and normals", "o cube", "mtllib nothing.mtl"]
    out += [f"{indent}v {x:.6f} {y:.6f} {z:.6f}" for x, y, z in v]
    out += [f"vt {u:.3f} {w:.3f}" for u, w in vt]
    out += [f"vn {x} {y} {z}" for x, y, z in vn]
    out += ["s off", "usemtl none"]
    for q, n in quads:
        out.append("f " + " ".join(f"{vi}/{k + 1}/{n}" for k, vi in enumerate(q)))
    return newline.join(out) + newline


OBJ_CASES = {
    "cube": cube_obj(),
    "cube crlf indented": cube_obj("\r\n", "  \t"),
    "triangles only": "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nf 1 2 3\nf 1 3 4\nf 2 3 4\n",
    "negative indices": "v 0 0 0\nv 1 0 0\nv 0 1 0\nf -3 -2 -1\nv 0 0 2\nf -1 -2 -3\nf 1 2 -1\n",
    "pentagon fan": "v 1 0 0\nv 0.3 0.95 0\nv -0.8 0.6 0\nv -0.8 -0.6 0\nv 0.3 -0.95 0\nf 1 2 3 4 5\n",
    "normals without texcoords": "v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvn 0 0 -1\nf 1//1 2//1 3//2\n",
    "texcoords without normals": "v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nf 1/1 2/2 3/3\n",
    "vp quirk and unknown commands": "v 0 0 0\nvp 0.5 0.5\nv 1 0 0\nv 0 1 0\ng grp\nl 1 2\nf 1 3 4\nf 1 2 3\n",
    "index zero is the null vertex": "v 5 5 5\nv 1 0 0\nv 0 1 0\nf 0 2 3\n",
    "octal and hex indices": "\n".join(f"v {i} {i * 2} {i * 3}" for i in range(1, 18)) + "\nf 010 0x10 17\nf 1 2 3\n",
    "exponents and signs": "v 1e-3 -2.5E2 +3\nv .5 -.25 1.\nv 1e10 1e-10 0\nf 1 2 3\n",
    "short vertex lines": "v 1 2\nv 3\nv\nv 4 5 6\nf 1 2 3\nf 2 3 4\n",
    "no trailing newline": "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3",
    "empty": "",
    "comments only": "# nothing\n# here\n",
}


@pytest.mark.parametrize("name", list(OBJ_CASES))
@pytest.mark.parametrize("winding", [0, 1])
def test_obj_matches_reference(oracle, name, winding):
    text = OBJ_CASES[name]
    ref = oracle.parse_obj(text, winding)
    assert ref is not None, "the reference rejects this input; move it to the error cases"
    pos, nrm, tex = lib.parse_obj(text, winding)
    assert pos.shape == ref[0].shape, f"{name}: {pos.shape[0]} triangles vs {ref[0].shape[0]}"
    assert np.array_equal(bits(pos), bits(ref[0])), f"{name}: positions differ"
    if ref[0].shape[0]:
        for mine, theirs, what in ((nrm, ref[1], "normals"), (tex, ref[2], "texture coordinates")):
            if theirs is not None and np.any(theirs):      # (reference flags with nothing fanned behind them are not compared)
                assert mine is not None and np.array_equal(bits(mine), bits(theirs)), f"{name}: {what} differ"


@pytest.mark.parametrize("text", [
    "v 0 0 0\nv 1 0 0\nf 1 2\n",                                         # two corners
    "v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\nf 1 2 3\n",  # normals on some faces only
    "v 0 0 0\nv 1 0 0\nv 0 1 0\nf " + " ".join(["1 2 3"] * 11) + "\n",   # 33 corners
])
def test_obj_errors_like_reference(oracle, text):
    assert oracle.parse_obj(text, 1) is None
    with pytest.raises(lib.BptError):
        lib.parse_obj(text, 1)


def test_obj_inputs_the_reference_cannot_survive_are_errors():
    with pytest.raises(lib.BptError):
        lib.parse_obj("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n", 1)           # index past the end (reference: out-of-bounds read)
    with pytest.raises(lib.BptError):
        lib.parse_obj("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 x3\n", 1)          # garbage corner (reference: infinite loop)


def test_obj_file_to_mesh(tmp_path, bpt):
    p = tmp_path / "cube.obj"
    p.write_text(OBJ_CASES["cube"])
    pos, nrm, tex = lib.parse_obj(None, 1, path=str(p))
    assert pos.shape == (12, 9) and nrm.shape == (12, 9) and tex.shape == (12, 9)
    s = bpt.Scene()
    m = s.create_mesh(pos, nrm)
    nodes, idx, tris = s.mesh_bvh(m)
    assert sorted(idx.tolist()) == list(range(12)) and nodes[0]["count"] == 0


# ---- Radiance HDR ------------------------------------------------------------------------------------------------------
def rgbe_from_float(img):
    """float RGB -> RGBE bytes (the usual frexp encoding)"""
    m = np.max(img, axis=-1)
    e = np.zeros(m.shape, np.int32)
    mant = np.zeros(m.shape)
    nz = m > 1e-32
    mant[nz], e[nz] = np.frexp(m[nz])
    scale = np.where(nz, mant * 256.0 / np.where(nz, m, 1), 0)
    out = np.zeros(img.shape[:-1] + (4,), np.uint8)
    out[..., :3] = np.clip(img * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(nz, e + 128, 0).astype(np.uint8)
    return out


def rle_channel(row):
    out = bytearray()
    i, n = 0, len(row)
    while i < n:
        run = 1
        while i + run < n and run < 127 and row[i + run] == row[i]:
            run += 1
        if run >= 4:
            out += bytes([128 + run, row[i]])
            i += run
        else:
            j = i
            while j < n and j - i < 128:
                r = 1
                while j + r < n and r < 4 and row[j + r] == row[j]:
                    r += 1
                if r >= 4:
                    break
                j += 1
            out += bytes([j - i]) + bytes(row[i:j])
            i = j
    return bytes(out)


def make_hdr(rgbe, res_line, header=("#?RADIANCE", "FORMAT=32-bit_rle_rgbe", "EXPOSURE=1.0")):
    h, w = rgbe.shape[:2]
    data = ("\n".join(header) + "\n\n" + res_line + "\n").encode()
    for y in range(h):
        data += struct.pack(">BBH", 2, 2, w)
        for c in range(4):
            data += rle_channel(rgbe[y, :, c].tolist())
    return data


def synthetic_rgbe(w, h, seed):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([0.2 + 0.8 * xx / w, 0.1 + yy / h, 0.5 + 0 * xx], -1) * np.exp(4 * rng.rand(h, w, 1) - 2)
    img[: h // 3] = np.array([1.5, 1.5, 1.5])            # flat area -> long runs
    rgbe = rgbe_from_float(img)
    rgbe[0, : w // 4, 3] = rng.randint(0, 10, size=w // 4)    # exponents <= 9 decode to black
    return rgbe


@pytest.mark.parametrize("w,h,res", [(64, 32, "-Y {h} +X {w}"), (300, 7, "-Y {h} +X {w}"), (96, 10, "+Y {h} +X {w}"),
                                     (40, 9, "-Y {h} -X {w}"), (1024, 3, "+Y {h} -X {w}")])
def test_hdr_matches_reference(oracle, w, h, res):
    data = make_hdr(synthetic_rgbe(w, h, w + h), res.format(w=w, h=h))
    ref = oracle.parse_hdr(data)
    assert ref is not None and ref.shape == (h, w, 3)
    mine = lib.parse_hdr(data)
    assert mine.shape == ref.shape and np.array_equal(bits(mine), bits(ref))
    assert float(mine.max()) > 1.0 and float(mine.min()) == 0.0


def test_hdr_header_variants_and_errors(oracle):
    rgbe = synthetic_rgbe(32, 4, 1)
    ok = make_hdr(rgbe, "-Y 4 +X 32", header=("#?RGBE", "PRIMARIES= 0.64 0.33 0.29 0.6 0.15 0.06 0.333 0.333", "FORMAT=32-bit_rle_xyz"))
    assert np.array_equal(bits(lib.parse_hdr(ok)), bits(oracle.parse_hdr(ok)))
    bad = [
        make_hdr(synthetic_rgbe(200, 2, 2), "-Y 2 +X 200"),                   # width's low byte >= 128: signed-char quirk
        make_hdr(rgbe, "-Y 4 +X 31"),                                         # scanline length != width
        make_hdr(rgbe, "Y 4 X 32"),                                           # resolution string
        ("#?RADIANCE\nFORMAT 32-bit_rle_rgbe\n\n-Y 4 +X 32\n").encode() + b"\x02\x02\x00\x20",   # FORMAT without '='
        ("#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n").encode(),                    # header never ends
        ("#?RADIANCE\n\n-Y 4 +X 32\n").encode() + b"\x01\x01\x00\x20" + b"\x00" * 64,             # old-style / flat scanlines
    ]
    for data in bad:
        assert oracle.parse_hdr(data) is None
        with pytest.raises(lib.BptError):
            lib.parse_hdr(data)


def test_hdr_file_to_skydome(tmp_path, bpt):
    data = make_hdr(synthetic_rgbe(64, 32, 5), "-Y 32 +X 64")
    p = tmp_path / "sky.hdr"
    p.write_bytes(data)
    s = bpt.Scene()
    L = bpt.load_library()
    import ctypes as C
    L.bpt_load_skydome_hdr.restype = C.c_int
    L.bpt_load_skydome_hdr.argtypes = [C.c_void_p, C.c_char_p]
    assert L.bpt_load_skydome_hdr(s.handle, str(p).encode()) == 0
    assert L.bpt_load_skydome_hdr(s.handle, str(tmp_path / "missing.hdr").encode()) != 0
