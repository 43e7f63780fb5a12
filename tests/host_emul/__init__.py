"""Test-only host build of the device narrowphase header (see emul.cpp)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libxp_emul.so")
_SRC = os.path.join(_HERE, "emul.cpp")
_CSRC = os.path.normpath(os.path.join(_HERE, "..", "..", "xp-game-engine_b200", "csrc"))
_HDRS = [os.path.join(_CSRC, "xp_narrowphase.cuh"), os.path.join(_CSRC, "xp_gridwalk.cuh")]


def lib():
    stale = (not os.path.exists(_SO)) or max([os.path.getmtime(_SRC)] + [os.path.getmtime(h) for h in _HDRS]) > os.path.getmtime(_SO)
    if stale:
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                        "-fno-math-errno", "-o", _SO, _SRC], check=True, capture_output=True)
    L = C.CDLL(_SO)
    vp = C.c_void_p
    L.emul_detect.argtypes = [vp, vp, vp, C.c_int]
    L.emul_detect.restype = C.c_int
    L.emul_respond.argtypes = [vp, vp, vp, vp]
    L.emul_respond.restype = C.c_int
    L.emul_broadphase.argtypes = [vp, vp, C.c_float]
    L.emul_broadphase.restype = C.c_int
    L.emul_detect_many.argtypes = [vp, vp, C.c_long, vp, vp, C.c_int]
    L.emul_detect_many.restype = None
    L.emul_grid_walk.argtypes = [vp, C.c_uint, C.c_uint, C.c_float, C.c_float, C.c_float, C.c_float, vp, C.c_float, vp, C.c_long, vp]
    L.emul_grid_walk.restype = C.c_long
    L.emul_bad_count.argtypes = [C.c_int]
    L.emul_bad_count.restype = C.c_long
    return L
