"""ORACLE-SIDE (test / reference-arm infrastructure): the procedural inputs of the BASELINE configurations from
oracle/_ref/libbpt_inputs.so -- the same source file the product library compiles
(buas_pathtracer_b200/csrc/procedural_inputs.cpp, built standalone by oracle/Makefile), so the reference arm of bench.py
builds its scene without loading libbpt.so.  tests/test_oracle_golden.py checks both libraries give identical bytes."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libbpt_inputs.so")
_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "inputs"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.inputs_make_displaced_icosphere.restype = C.c_uint32
        L.inputs_make_displaced_icosphere.argtypes = [C.c_uint32, C.c_float, C.c_void_p]
        L.inputs_make_procedural_skydome.restype = C.c_int
        L.inputs_make_procedural_skydome.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


def make_displaced_icosphere(level, amplitude=0.08):
    n = lib().inputs_make_displaced_icosphere(level, amplitude, None)
    out = np.empty((n, 9), np.float32)
    lib().inputs_make_displaced_icosphere(level, amplitude, out.ctypes.data)
    return out


def make_procedural_skydome(w, h):
    out = np.empty((h, w, 3), np.float32)
    rc = lib().inputs_make_procedural_skydome(w, h, out.ctypes.data)
    assert rc == 0
    return out
