"""Loader for the C-ABI product library (csrc/libxp_physics_cuda.so).

There is no CPU fallback and no alternative backend: if the shared library is missing, or no
CUDA device is usable, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.normpath(os.path.join(_PKG, "..", "csrc"))
SO_PATH = os.environ.get("XP_PHYSICS_CUDA_SO") or os.path.join(CSRC, "libxp_physics_cuda.so")   # override: A/B builds
HEADER = os.path.normpath(os.path.join(_PKG, "..", "..", "include", "xp_physics_cuda.h"))

# constants of include/xp_physics_cuda.h
XP_OK, XP_EINVAL, XP_ECUDA, XP_ENOMEM, XP_ESTATE, XP_ECAP = 0, -1, -2, -3, -4, -5
ST_RADIUS_NE_1, ST_ASSERT_NEG_DISTANCE, ST_ITER_CAP = 1, 2, 4
HORIZON_INF, HORIZON_UNIT = 0, 1
OPT_ARITH, OPT_METHOD, OPT_TIMING, OPT_SORT_SPHERES, OPT_MAX_PAIRS, OPT_PRUNE, OPT_FAR_EXACT, OPT_GRID_WALK = 1, 2, 3, 4, 5, 6, 7, 8
OPT_CONTACT_RECORD = 9
OPT_LEAF_RUNS = 10
ARITH_EXACT, ARITH_FAST, ARITH_EXACT_LITERAL, ARITH_FAST_LITERAL = 0, 1, 2, 3
ARITH = {"exact": 0, "fast": 1, "exact_literal": 2, "fast_literal": 3}
METHODS = {"wide_fused": 0, "lbvh_fused": 1, "lbvh_pairs": 2, "brute": 3, "wide_pairs": 4, "q4_pairs": 5}
STAGE_NAMES = ["build_bounds", "build_morton", "build_sort", "build_pack", "build_tree", "build_refit",
               "ray_setup", "traverse", "narrowphase", "response", "finalize", "sphere_sort"]


class XpStats(C.Structure):
    _fields_ = [("narrowphase_tests", C.c_uint64), ("broadphase_pairs", C.c_uint64),
                ("node_visits", C.c_uint64), ("kernel_launches", C.c_uint32),
                ("iterations_run", C.c_uint32), ("stage_ms", C.c_float * 12),
                ("path_tests", C.c_uint64 * 5)]


class XpError(RuntimeError):
    def __init__(self, code: int, what: str):
        self.code = code
        super().__init__(what)


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc build of libxp_physics_cuda.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return SO_PATH


_lib = None


def lib() -> C.CDLL:
    """The loaded library.  Raises if it has not been built: there is nothing to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise XpError(XP_ECUDA, f"{SO_PATH} is missing: run `make -C {CSRC}` (or __graft_entry__.build()); "
                                "xp_physics_cuda has no CPU fallback")
    L = C.CDLL(SO_PATH)
    vp, u64, u32, i32, i64 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int64
    sig = {
        "xp_device_count": ([vp], i32),
        "xp_scene_create": ([i32, vp], i32),
        "xp_scene_destroy": ([vp], None),
        "xp_scene_set_stream": ([vp, vp], i32),
        "xp_scene_set_option": ([vp, i32, i64], i32),
        "xp_scene_get_stats": ([vp, vp], i32),
        "xp_scene_set_triangles": ([vp, vp, u64], i32),
        "xp_scene_set_triangles_dev": ([vp, vp, u64], i32),
        "xp_scene_set_positions": ([vp, vp, u64], i32),
        "xp_scene_set_heightmap": ([vp, vp, u32, u32, C.c_float, C.c_float, C.c_float], i32),
        "xp_scene_set_clipmap": ([vp, vp, u32, u32, vp, u32, C.c_float], i32),
        "xp_scene_build": ([vp], i32),
        "xp_scene_triangle_count": ([vp], u64),
        "xp_detect_batch": ([vp, vp, u64, u32, vp, vp, vp, vp], i32),
        "xp_detect_batch_dev": ([vp, vp, u64, u32, vp, vp, vp, vp], i32),
        "xp_detect_batch_submit": ([vp, vp, u64, u32, vp, vp, vp, vp, vp], i32),
        "xp_detect_batch_wait": ([vp, i32], i32),
        "xp_detect_batch_compact": ([vp, vp, u64, u32, vp, vp], i32),
        "xp_detect_batch_compact_dev": ([vp, vp, u64, u32, vp, vp], i32),
        "xp_detect_batch_compact_submit": ([vp, vp, u64, u32, vp, vp, vp], i32),
        "xp_collision_response_batch": ([vp, vp, u64, u32, u32, vp, vp, vp, vp], i32),
        "xp_collision_response_batch_dev": ([vp, vp, u64, u32, u32, vp, vp, vp, vp], i32),
        "xp_collide_and_slide_batch": ([vp, vp, vp, u64, u32, vp, vp, vp, vp], i32),
        "xp_collide_and_slide_batch_dev": ([vp, vp, vp, u64, u32, vp, vp, vp, vp], i32),
        "xp_sphere_triangle_detect_collision": ([i32, vp, vp, vp], i32),
        "xp_sphere_triangle_calculate_response": ([i32, vp, vp, vp], i32),
        "xp_broadphase_pairs": ([vp, vp, u64, u32, u64, vp, vp, vp], i32),
        "xp_scene_debug_tree": ([vp, vp, vp, vp, vp, vp], i32),
        "xp_scene_debug_leaf_runs": ([vp, vp], i32),
        "xp_detect_contacts_dev": ([vp, vp, u64, u32, vp], i32),
        "xp_exchange_alloc": ([vp, u64, vp, vp], i32),
        "xp_exchange_free": ([vp, vp], i32),
        "xp_exchange_open": ([vp, vp, vp], i32),
        "xp_exchange_close": ([vp, vp], i32),
        "xp_detect_contacts_exchange_dev": ([vp, vp, u64, u32, vp, i32, u64], i32),
        "xp_detect_contacts_multicast_dev": ([vp, vp, u64, u32, vp, u64], i32),
        "xp_scene_set_exchange_stream": ([vp, vp], i32),
        "xp_detect_contacts_exchange_submit": ([vp, vp, u64, u32, vp, i32, i32, u64, vp], i32),
        "xp_detect_contacts_exchange_wait": ([vp, i32], i32),
        "xp_probe_fp32_peak": ([i32, vp], i32),
        "xp_debug_check_arith": ([i32, u64, u64, vp], i32),
        "xp_strerror": ([i32], C.c_char_p),
        "xp_last_cuda_error": ([], C.c_char_p),
        "xp_version": ([], C.c_char_p),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)          # AttributeError here = the library does not export the ABI
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "xp_device_count", "xp_scene_create", "xp_scene_destroy", "xp_scene_set_stream", "xp_scene_set_option",
    "xp_scene_get_stats", "xp_scene_set_triangles", "xp_scene_set_triangles_dev", "xp_scene_set_positions", "xp_scene_set_heightmap", "xp_scene_set_clipmap",
    "xp_scene_build", "xp_scene_triangle_count", "xp_detect_batch", "xp_detect_batch_dev",
    "xp_detect_batch_submit", "xp_detect_batch_wait", "xp_detect_batch_compact", "xp_detect_batch_compact_dev", "xp_detect_batch_compact_submit",
    "xp_collision_response_batch", "xp_collision_response_batch_dev", "xp_collide_and_slide_batch", "xp_collide_and_slide_batch_dev", "xp_sphere_triangle_detect_collision",
    "xp_sphere_triangle_calculate_response", "xp_broadphase_pairs", "xp_scene_debug_tree", "xp_scene_debug_leaf_runs",
    "xp_detect_contacts_dev", "xp_exchange_alloc", "xp_exchange_free", "xp_exchange_open", "xp_exchange_close",
    "xp_detect_contacts_exchange_dev", "xp_detect_contacts_multicast_dev", "xp_scene_set_exchange_stream",
    "xp_detect_contacts_exchange_submit", "xp_detect_contacts_exchange_wait", "xp_probe_fp32_peak", "xp_debug_check_arith", "xp_strerror", "xp_last_cuda_error", "xp_version",
]


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        L = lib()
        msg = L.xp_strerror(rc).decode()
        if rc == XP_ECUDA or rc == XP_ENOMEM:
            msg += ": " + L.xp_last_cuda_error().decode()
        raise XpError(rc, f"{what or 'xp_physics_cuda'}: {msg}")
    return rc
