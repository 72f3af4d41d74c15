"""ORACLE-ONLY (test infrastructure), kind "port": numpy restatement of the reference's display-loop resolve
(Raytracer/raytracer.cpp:2103-2172) and BMP header (assets.cpp:671-724).  Since round 2 the reference's own resolve loop
is compiled into oracle/_ref (oracle/tools/slice_resolve.py) and tests/test_resolve.py compares the device kernel with THAT;
this port is pinned against it there (tests/test_resolve.py::test_resolve_port_agrees_with_the_references_resolve) and is
only still used for the one path the reference does not have: resolving without a dither tile.  exp/pow go through numpy's
float32 routines, and remap_tpdf's SSE rsqrt approximation (my_math.h:50-54) is replaced by an exact 1/sqrt: +-1 LSB."""
import struct

import numpy as np

f32 = np.float32


def sigmoidal_contrast(x, contrast, midpoint):            # raytracer.cpp:69-84
    x = x.astype(f32)
    lo = f32(midpoint) * ((f32(1.0) / f32(midpoint)) * x) ** 2
    y = f32(1.0) / (f32(1.0) - f32(midpoint))
    hi = f32(1.0) - (f32(1.0) - f32(midpoint)) * (y - y * x) ** 2
    curve = np.where(x < f32(midpoint), lo, hi).astype(f32)
    return (x * (f32(1.0) - f32(contrast)) + curve * f32(contrast)).astype(f32)


def remap_tpdf(x):                                         # raytracer.cpp:125-132
    orig = f32(2.0) * x.astype(f32) - f32(1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = orig * (f32(1.0) / np.sqrt(np.abs(orig))).astype(f32)
    v = np.where(f32(-1.0) > v, f32(-1.0), v)            # max(-1, x) with the reference's ternary: NaN -> NaN
    return (v - np.where(v < 0, f32(-1.0), f32(1.0))).astype(f32)


def resolve_bgra8(film, exposure=0.0, tonemapping=True, srgb_transform=True, midpoint=0.5, contrast=0.0, dither=None):
    film = film.astype(f32)
    h, w, _ = film.shape
    rgb, wt = film[..., :3], film[..., 3]
    out = np.zeros((h, w, 3), f32)
    nan = np.isnan(film).any(axis=2)
    ok = (~nan) & (wt > f32(0.001))
    with np.errstate(all="ignore"):
        c = rgb / wt[..., None]
        c = np.where(c > 0, c, f32(0.0)).astype(f32)      # max(c, 0) ternary
        if exposure != 0.0:
            c = c * np.power(f32(2.0), f32(exposure))
        if tonemapping:
            c = (f32(1.0) - np.exp(-c)).astype(f32)
        if srgb_transform:
            c = np.power(c, f32(1.0) / f32(2.23333)).astype(f32)
        if contrast != 0.0:
            c = sigmoidal_contrast(c, contrast, midpoint)
        c = c * f32(255.0)
        if dither is not None:
            dh, dw, _ = dither.shape
            ys, xs = np.mgrid[0:h, 0:w]
            d = dither[ys & (dh - 1), xs & (dw - 1)].astype(f32)
            c = c + (f32(0.5) + remap_tpdf((f32(1.0) / f32(255.0)) * d))
    out[ok] = c[ok]
    neg = (~nan) & (~ok) & (wt < f32(-0.01))
    out[neg] = np.stack([-255.0 * wt[neg], np.zeros(neg.sum()), -255.0 * wt[neg]], axis=1)
    out[nan] = (0.0, 255.0, 255.0)
    q = np.clip(np.nan_to_num(out, nan=0.0), 0.0, 255.0).astype(np.uint8).astype(np.uint32)   # (u8)clamp(): truncation
    return (np.uint32(255) << 24) | (q[..., 0] << 16) | (q[..., 1] << 8) | q[..., 2]


def bitmap_bytes(pixels):                                  # assets.cpp:671-724
    h, w = pixels.shape
    size = 4 * w * h
    hdr = struct.pack("<HIHHIIiiHHIIiiII", 0x4D42, 54 + size, 0, 0, 54, 40, w, -h, 1, 32, 0, size, 4096, 4096, 0, 0)
    return hdr + pixels.astype("<u4").tobytes()
