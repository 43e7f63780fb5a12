"""The C-ABI library builds, loads and exports every symbol include/xp_physics_cuda.h declares.
No compute call is made here (no GPU in the CPU test tier)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from xp_physics_cuda import _lib


@pytest.fixture(scope="module")
def L():
    _lib.build()
    return _lib.lib()


def _declared_functions():
    src = open(_lib.HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(xp_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_expected_entry_points():
    names = _declared_functions()
    for must in ("xp_scene_create", "xp_scene_set_triangles", "xp_scene_set_positions", "xp_scene_build",
                 "xp_detect_batch", "xp_collision_response_batch", "xp_broadphase_pairs", "xp_scene_destroy",
                 "xp_strerror", "xp_sphere_triangle_detect_collision", "xp_sphere_triangle_calculate_response"):
        assert must in names


def test_library_exports_every_declared_symbol(L):
    raw = C.CDLL(_lib.SO_PATH)
    for name in _declared_functions():
        assert hasattr(raw, name), f"{name} declared in include/xp_physics_cuda.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_functions()


def test_struct_sizes_match_the_reference_layouts():
    from xp_physics_cuda import COLLISION_DTYPE, CONTACT_DTYPE, RESPONSE_DTYPE
    assert RESPONSE_DTYPE.itemsize == 28      # Sphere{Vec3,f32} + Vec3 (response.rs:5-8)
    assert COLLISION_DTYPE.itemsize == 32     # 2 f32 + 2 Vec3 (collision.rs:5-10)
    assert CONTACT_DTYPE.itemsize == 40
    assert C.sizeof(_lib.XpStats) == 8 * 3 + 4 * 2 + 4 * 12 + 8 * 5


def test_strerror_and_version(L):
    assert L.xp_strerror(0) == b"ok"
    assert b"invalid" in L.xp_strerror(_lib.XP_EINVAL)
    assert b"sm_100a" in L.xp_version()


def test_argument_validation_without_a_device(L):
    assert L.xp_scene_create(0, None) == _lib.XP_EINVAL
    assert L.xp_scene_set_option(None, 1, 0) == _lib.XP_EINVAL
    assert L.xp_scene_build(None) == _lib.XP_EINVAL
    assert L.xp_detect_batch(None, None, 0, 0, None, None, None, None) == _lib.XP_EINVAL
    assert L.xp_scene_triangle_count(None) == 0
    L.xp_scene_destroy(None)


def test_new_entry_points_validate_arguments_without_a_device(L):
    t = C.c_int(-1)
    assert L.xp_detect_batch_submit(None, None, 0, 0, None, None, None, None, C.byref(t)) == _lib.XP_EINVAL
    assert L.xp_detect_batch_wait(None, 0) == _lib.XP_EINVAL
    assert L.xp_scene_set_exchange_stream(None, None) == _lib.XP_EINVAL
    assert L.xp_detect_contacts_exchange_submit(None, None, 0, 0, None, 0, -1, 0, C.byref(t)) == _lib.XP_EINVAL
    assert L.xp_detect_contacts_exchange_wait(None, 0) == _lib.XP_EINVAL
    assert L.xp_detect_contacts_multicast_dev(None, None, 0, 0, None, 0) == _lib.XP_EINVAL
    assert L.xp_scene_set_clipmap(None, None, 0, 0, None, 0, 2.0) == _lib.XP_EINVAL


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99, not only as C++."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        import pytest
        pytest.skip("no C compiler")
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", _lib.HEADER],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_cpu_fallback(L):
    """Without a CUDA device the product refuses to work instead of computing on the CPU."""
    n = C.c_int(-1)
    rc = L.xp_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert L.xp_scene_create(0, C.byref(h)) in (_lib.XP_ECUDA, _lib.XP_EINVAL)
    assert not h.value
    import xp_physics_cuda as xp
    with pytest.raises(xp.XpError):
        xp.Scene(0)
    resp = np.zeros(1, xp.RESPONSE_DTYPE)
    tri = np.zeros((1, 3, 3), np.float32)
    out = np.zeros(1, xp.COLLISION_DTYPE)
    rc = L.xp_sphere_triangle_detect_collision(0, resp.ctypes.data_as(C.c_void_p), tri.ctypes.data_as(C.c_void_p),
                                               out.ctypes.data_as(C.c_void_p))
    assert rc == _lib.XP_ECUDA


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may import, link or name it."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(_lib.HEADER)))
    pkg = os.path.join(root, "xp-game-engine_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".rs", "Makefile")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "xp_oracle" not in text and "xpo_" not in text and "libxp_oracle" not in text, os.path.join(dp, f)


def test_python_constants_match_the_header():
    """the ctypes mirror restates the header's numeric constants: options, arithmetic modes, methods, status bits, errors"""
    src = open(_lib.HEADER).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(XP_[A-Z0-9_]+)\s+(-?\d+)u?\b", src)}
    for name, value in defs.items():
        if name.startswith("XP_OPT_"):
            assert getattr(_lib, name[3:]) == value, name
    assert len([k for k in defs if k.startswith("XP_OPT_")]) == 10
    assert {k: defs["XP_ARITH_" + k.upper()] for k in _lib.ARITH} == _lib.ARITH
    assert {k: defs["XP_METHOD_" + k.upper()] for k in _lib.METHODS} == _lib.METHODS
    assert (defs["XP_ST_RADIUS_NE_1"], defs["XP_ST_ASSERT_NEG_DISTANCE"], defs["XP_ST_ITER_CAP"]) == (
        _lib.ST_RADIUS_NE_1, _lib.ST_ASSERT_NEG_DISTANCE, _lib.ST_ITER_CAP)
    for e in ("XP_OK", "XP_EINVAL", "XP_ECUDA", "XP_ENOMEM", "XP_ESTATE", "XP_ECAP"):
        assert defs[e] == getattr(_lib, e), e
    assert (defs["XP_HORIZON_INF"], defs["XP_HORIZON_UNIT"]) == (_lib.HORIZON_INF, _lib.HORIZON_UNIT)
