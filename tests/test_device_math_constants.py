"""CPU check of the double-precision sincos that csrc/device_math.cuh uses for the hemisphere azimuths (sincos_f): the
constants are read out of the CUDA source and the routine is restated here with exact fused multiply-adds (fractions), then
compared with libm over [0, 2 pi] and beyond.  The device result is (float)of this double, and the reference calls glibc's
sinf / cosf; what parity needs from the routine is an error far below half a float ulp -- 1e-15 is asserted (measured 1.7e-16).
A mistyped coefficient would otherwise only show up as a slightly lower bit-exact fraction in the GPU parity tests."""
import math
import os
import re
from fractions import Fraction

import numpy as np

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "buas_pathtracer_b200", "csrc", "device_math.cuh")


def _fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def _constants():
    text = open(SRC).read()
    body = text[text.index("BPT_D void sincos_f(float x, float& s, float& c) {"):]
    body = body[:body.index("\n}\n")]
    num = r"(-?\d+\.\d+(?:e[-+]?\d+)?)"
    two_over_pi = float(re.search(r"rint\(dx\*" + num + r"\)", body).group(1))
    pio2 = [float(m) for m in re.findall(r"fma\(-q, " + num + r",", body)]
    ps0 = float(re.search(r"double ps = " + num + ";", body).group(1))
    ps = [float(m) for m in re.findall(r"ps = fma\(ps, z, " + num + r"\);", body)]
    s1 = float(re.search(r"fma\(z, ps, " + num + r"\)", body).group(1))
    pc0 = float(re.search(r"double pc = " + num + ";", body).group(1))
    pc = [float(m) for m in re.findall(r"pc = fma\(pc, z, " + num + r"\);", body)]
    assert len(pio2) == 2 and len(ps) == 4 and len(pc) == 5, (pio2, ps, pc)
    return two_over_pi, pio2, [ps0] + ps, s1, [pc0] + pc


def _sincos(x, k):
    two_over_pi, (hi, lo), ps, s1, pc = k
    q = float(np.rint(x * two_over_pi))
    r = _fma(-q, hi, x)
    r = _fma(-q, lo, r)
    z = r * r
    p = ps[0]
    for c in ps[1:]:
        p = _fma(p, z, c)
    sn = _fma(z * r, _fma(z, p, s1), r)
    p = pc[0]
    for c in pc[1:]:
        p = _fma(p, z, c)
    cs = _fma(z * z, p, _fma(-0.5, z, 1.0))
    n = int(q)
    a, b = (cs, sn) if n & 1 else (sn, cs)
    return (-a if n & 2 else a), (-b if (n + 1) & 2 else b)


def test_sincos_constants_give_double_accuracy():
    k = _constants()
    assert abs(k[0] - 2 / math.pi) < 1e-16 and abs(k[1][0] - math.pi / 2) < 1e-16 and abs(k[1][0] + k[1][1] - math.pi / 2) < 1e-16
    rng = np.random.RandomState(5)
    xs = np.concatenate([rng.rand(4000).astype(np.float32) * np.float32(6.2831855),
                         (rng.rand(500).astype(np.float32) * 128 - 64).astype(np.float32),
                         np.float32([0.0, 1e-30, 1.5707964, 3.1415927, 4.712389, 6.2831855, -6.2831855, 63.99, -64.0])])
    worst = 0.0
    for x in xs:
        x = float(x)
        s, c = _sincos(x, k)
        worst = max(worst, abs(s - math.sin(x)), abs(c - math.cos(x)))
        assert np.float32(s) == np.float32(math.sin(x)) and np.float32(c) == np.float32(math.cos(x)), x
    assert worst < 1e-15, worst
