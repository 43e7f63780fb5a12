"""ctypes binding of the CPU oracle (oracle/libxp_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (xp_physics_cuda) never imports it.
See oracle/xp_oracle.h for what the oracle restates and how it is pinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libxp_oracle.so")

# Layouts of the reference's value types (xp_physics/src/{sphere,triangle,response,collision}.rs)
RESPONSE_DTYPE = np.dtype([("c", "<f4", (3,)), ("r", "<f4"), ("movement", "<f4", (3,))])
COLLISION_DTYPE = np.dtype([("time_to", "<f4"), ("distance_to", "<f4"),
                            ("position", "<f4", (3,)), ("intersection", "<f4", (3,))])
assert RESPONSE_DTYPE.itemsize == 28 and COLLISION_DTYPE.itemsize == 32

ST_RADIUS_NE_1 = 1
ST_ASSERT_NEG_DISTANCE = 2
ST_ITER_CAP = 4


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc, -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("xp_oracle.c", "xp_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "libxp_oracle.so"],
                       check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u64, u32, i32, f32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_float
        L.xpo_get_roots.argtypes = [f32, f32, f32, vp, vp]
        L.xpo_get_roots.restype = i32
        L.xpo_point_in_triangle.argtypes = [vp, vp, vp, vp]
        L.xpo_point_in_triangle.restype = i32
        L.xpo_detect.argtypes = [vp, vp, vp]
        L.xpo_detect.restype = i32
        L.xpo_respond.argtypes = [vp, vp, vp, vp]
        L.xpo_respond.restype = i32
        L.xpo_closest.argtypes = [vp, vp, u64, vp, vp]
        L.xpo_closest.restype = i32
        L.xpo_collision_response.argtypes = [vp, vp, u64, u32, vp, vp, vp, vp]
        L.xpo_collision_response.restype = None
        L.xpo_collision_response_non_trianulated.argtypes = [vp, vp, u64, u32, vp, vp, vp]
        L.xpo_collision_response_non_trianulated.restype = i32
        L.xpo_detect_batch.argtypes = [vp, u64, vp, u64, i32, vp, vp, vp, vp, vp]
        L.xpo_detect_batch.restype = None
        L.xpo_collision_response_batch.argtypes = [vp, u64, vp, u64, u32, i32, vp, vp, vp, vp, vp]
        L.xpo_collision_response_batch.restype = None
        L.xpo_broadphase_accept.argtypes = [vp, vp, f32]
        L.xpo_broadphase_accept.restype = i32
        L.xpo_broadphase_pairs.argtypes = [vp, u64, vp, u64, f32, u64, vp, vp, vp]
        L.xpo_broadphase_pairs.restype = None
        L.xpo_triangle_aabb.argtypes = [vp, vp, vp]
        L.xpo_triangle_aabb.restype = None
        L.xpo_detect_margin_f64.argtypes = [vp, vp, vp, vp]
        L.xpo_detect_margin_f64.restype = C.c_double
        L.xpo_grazing_batch.argtypes = [vp, u64, vp, u64, f32, i32, vp]
        L.xpo_grazing_batch.restype = None
        L.xpo_max_threads.argtypes = []
        L.xpo_max_threads.restype = i32
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def make_responses(c, movement, r=1.0) -> np.ndarray:
    c = np.asarray(c, np.float32).reshape(-1, 3)
    m = np.asarray(movement, np.float32).reshape(-1, 3)
    out = np.zeros(len(c), RESPONSE_DTYPE)
    out["c"], out["movement"], out["r"] = c, m, r
    return out


def as_triangles(t) -> np.ndarray:
    t = np.ascontiguousarray(np.asarray(t, np.float32).reshape(-1, 3, 3))
    return t


def max_threads() -> int:
    return lib().xpo_max_threads()


def get_roots(a, b, c):
    r0, r1 = C.c_float(), C.c_float()
    ok = lib().xpo_get_roots(a, b, c, C.addressof(r0), C.addressof(r1))
    return (r0.value, r1.value) if ok else None


def point_in_triangle(p0, p1, p2, p) -> bool:
    a = [np.asarray(x, np.float32).copy() for x in (p0, p1, p2, p)]
    return bool(lib().xpo_point_in_triangle(*[_p(x) for x in a]))


def detect(resp: np.ndarray, tri):
    """sphere_triangle_detect_collision: returns a COLLISION_DTYPE scalar, None, or raises on r != 1."""
    resp = np.ascontiguousarray(resp).reshape(1)
    tri = as_triangles(tri)
    out = np.zeros(1, COLLISION_DTYPE)
    r = lib().xpo_detect(_p(resp), _p(tri), _p(out))
    if r < 0:
        raise AssertionError("assert_eq!(sphere.r, 1.0)")
    return out[0] if r == 1 else None


def respond(resp: np.ndarray, col: np.ndarray):
    resp = np.ascontiguousarray(resp).reshape(1)
    col = np.ascontiguousarray(col).reshape(1)
    out = np.zeros(1, RESPONSE_DTYPE)
    n = np.zeros(3, np.float32)
    r = lib().xpo_respond(_p(resp), _p(col), _p(out), _p(n))
    return (out[0] if r == 0 else None), n


def collision_response(resp: np.ndarray, tris, max_iter: int = 0):
    resp = np.ascontiguousarray(resp).reshape(1)
    tris = as_triangles(tris)
    out = np.zeros(1, RESPONSE_DTYPE)
    it, st = C.c_uint32(), C.c_uint32()
    n = np.zeros(3, np.float32)
    lib().xpo_collision_response(_p(resp), _p(tris), len(tris), max_iter, _p(out),
                                 C.addressof(it), C.addressof(st), _p(n))
    return out[0], it.value, st.value, n


def detect_batch(resps: np.ndarray, tris, threads: int = 0):
    resps = np.ascontiguousarray(resps)
    tris = as_triangles(tris)
    n = len(resps)
    out = np.zeros(n, COLLISION_DTYPE)
    hit = np.zeros(n, np.uint8)
    ti = np.zeros(n, np.uint32)
    st = np.zeros(n, np.uint32)
    tests = C.c_uint64()
    lib().xpo_detect_batch(_p(resps), n, _p(tris), len(tris), threads, _p(out), _p(hit), _p(ti),
                           _p(st), C.addressof(tests))
    return out, hit, ti, st, tests.value


def collision_response_batch(resps: np.ndarray, tris, max_iter: int = 0, threads: int = 0):
    resps = np.ascontiguousarray(resps)
    tris = as_triangles(tris)
    n = len(resps)
    out = np.zeros(n, RESPONSE_DTYPE)
    it = np.zeros(n, np.uint32)
    st = np.zeros(n, np.uint32)
    ln = np.zeros((n, 3), np.float32)
    tests = C.c_uint64()
    lib().xpo_collision_response_batch(_p(resps), n, _p(tris), len(tris), max_iter, threads,
                                       _p(out), _p(it), _p(st), _p(ln), C.addressof(tests))
    return out, it, st, ln, tests.value


def broadphase_pairs(resps: np.ndarray, tris, horizon: float = np.inf):
    resps = np.ascontiguousarray(resps)
    tris = as_triangles(tris)
    cnt = C.c_uint64()
    cap = max(1, min(len(resps) * len(tris), 1 << 26))
    si = np.zeros(cap, np.uint32)
    ti = np.zeros(cap, np.uint32)
    lib().xpo_broadphase_pairs(_p(resps), len(resps), _p(tris), len(tris), horizon, cap, _p(si),
                               _p(ti), C.addressof(cnt))
    k = cnt.value
    if k > cap:
        raise MemoryError("broadphase pair capacity exceeded")
    return si[:k].copy(), ti[:k].copy()


def triangle_aabbs(tris):
    tris = as_triangles(tris)
    lo = np.zeros((len(tris), 3), np.float32)
    hi = np.zeros((len(tris), 3), np.float32)
    L = lib()
    for j in range(len(tris)):
        L.xpo_triangle_aabb(_p(tris[j:j + 1]), _p(lo[j:j + 1]), _p(hi[j:j + 1]))
    return lo, hi


def grazing_margins(resps: np.ndarray, tris, horizon: float = np.inf, threads: int = 0):
    resps = np.ascontiguousarray(resps)
    tris = as_triangles(tris)
    out = np.zeros(len(resps), np.float64)
    lib().xpo_grazing_batch(_p(resps), len(resps), _p(tris), len(tris), horizon, threads, _p(out))
    return out
