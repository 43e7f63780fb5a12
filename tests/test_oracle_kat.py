"""Pins the CPU oracle against every known-answer test the reference holds for the path
 plus hand-derived known answers for the unpinned parts."""
import numpy as np

R = None


def _R(o, c, m, r=1.0):
    return o.make_responses([c], [m], r)


# --- the reference's own tests -------------------------------------------------------------
def test_detect_where_collision_inside_triangle(oracle):
    # xp_physics/src/collision_detect.rs:222-232
    c = oracle.detect(_R(oracle, [0, 4, 0], [0, -2, 0]), [[-2, 2, -2], [-2, 2, 2], [2, 2, 0]])
    assert c is not None and c["time_to"] == np.float32(0.5)


def test_detect_where_collision_against_vertex(oracle):
    # collision_detect.rs:235-246
    c = oracle.detect(_R(oracle, [0, 4, 0], [0, -8, 0]), [[0, 0, 0], [-2, -1, 0], [2, -1, 0]])
    assert c is not None and c["time_to"] == np.float32(0.375)


def test_detect_where_collision_against_edge(oracle):
    # collision_detect.rs:249-260
    c = oracle.detect(_R(oracle, [0, 4, 0], [0, -8, 0]), [[0, -2, 0], [-2, -1, 0], [2, -1, 0]])
    assert c is not None and c["time_to"] == np.float32(0.5)


def test_triangle_normal_orientation(oracle):
    # low_poly_nice_graphics/src/mesh/mesh.rs:342-359 (`check_understanding_normal_calculation_0/1`): the same
    # cross(p1 - p0, p2 - p0) orientation as glm::triangle_normal (triangle.rs:44-46); counter-clockwise -> +y
    import ctypes as C

    class V3(C.Structure):
        _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]
    L = oracle.lib()
    L.xpo_triangle_normal.restype = V3
    L.xpo_triangle_normal.argtypes = [C.c_void_p]
    for tri in ([[0, 0, 0], [0, 0, 1], [1, 0, 0]], [[-1, 0, -1], [0, 0, 0], [0, 0, -1]]):
        t = np.asarray(tri, np.float32)
        n = L.xpo_triangle_normal(t.ctypes.data_as(C.c_void_p))
        assert (n.x, n.y, n.z) == (0.0, 1.0, 0.0)


def test_point_in_triangle(oracle):
    # xp_physics/src/triangle.rs:62-76
    t = ([-1, 0, 0], [0, 1, 1], [1, 0, 0])
    assert oracle.point_in_triangle(*t, [0, 0.5, 0.5])
    assert oracle.point_in_triangle(*t, [0, 0.6, 0.6])
    assert oracle.point_in_triangle(*t, [0, 0.99, 0.99])
    assert not oracle.point_in_triangle(*t, [0, 1.001, 1.001])
    assert oracle.point_in_triangle(*t, [0, 0, 0])
    assert oracle.point_in_triangle(*t, [0, 1, 1])          # u + v == 1 inclusive


# --- derived known answers ------------------------------------------
def test_collision_fields(oracle):
    c = oracle.detect(_R(oracle, [0, 4, 0], [0, -2, 0]), [[-2, 2, -2], [-2, 2, 2], [2, 2, 0]])
    assert c["distance_to"] == 1.0
    assert list(c["position"]) == [0, 3, 0]
    assert list(c["intersection"]) == [0, 2, 0]              # position + m_hat * r  (collision_detect.rs:193)


def test_hit_beyond_the_step_is_reported(oracle):
    # quirk B2: only t > 0 is checked on the vertex/edge path (collision_detect.rs:201-202)
    c = oracle.detect(_R(oracle, [0, 40, 0], [0, -8, 0]), [[0, 0, 0], [-2, -1, 0], [2, -1, 0]])
    assert c is not None and c["time_to"] == np.float32(4.875)


def test_back_face_is_culled(oracle):
    # collision_detect.rs:166: the face triangle seen from below
    assert oracle.detect(_R(oracle, [0, 0, 0], [0, 2, 0]), [[-2, 2, -2], [-2, 2, 2], [2, 2, 0]]) is None


def test_radius_assert(oracle):
    import pytest
    with pytest.raises(AssertionError):
        oracle.detect(_R(oracle, [0, 4, 0], [0, -2, 0], r=0.5), [[-2, 2, -2], [-2, 2, 2], [2, 2, 0]])


def test_get_roots(oracle):
    # xp_math/src/lib.rs:3-12: x^2 - 3x + 2 -> (1, 2); det < 0 -> None; a == 0 is not guarded
    assert oracle.get_roots(1.0, -3.0, 2.0) == (1.0, 2.0)
    assert oracle.get_roots(1.0, 0.0, 1.0) is None
    r0, r1 = oracle.get_roots(0.0, 1.0, 1.0)
    assert np.isinf(r0) and np.isnan(r1)


def test_response_assert_status(oracle):
    # quirk B6: d = 1 - |m|(1-t) < 0 -> the reference panics at collision_response.rs:22
    tri = [[-2, 2, -2], [-2, 2, 2], [2, 2, 0]]
    out, it, st, n = oracle.collision_response(_R(oracle, [0, 4, 0], [0, -4, 0]), tri)
    assert st == oracle.ST_ASSERT_NEG_DISTANCE and it == 0
    assert list(out["c"]) == [0, 4, 0] and list(out["movement"]) == [0, -4, 0]


def test_response_marches_along_the_ray(oracle):
    # quirk B5: slide normal == -m_hat, new centre one unit past the contact position, new movement == m_hat
    tri = [[-2, 2, -2], [-2, 2, 2], [2, 2, 0]]
    out, it, st, n = oracle.collision_response(_R(oracle, [0.1, 3.2, 0.05], [0.3, -0.5, 0.1]), tri)
    assert it == 1 and st == 0
    np.testing.assert_allclose(n, [-0.5070925, 0.8451543, -0.1690309], atol=2e-7)
    np.testing.assert_allclose(out["c"], [0.7270926, 2.1548457, 0.2590309], atol=3e-7)
    np.testing.assert_allclose(out["movement"], -n, atol=3e-7)


def test_min_by_ties_go_to_the_later_triangle(oracle):
    # lib.rs:36-42: the comparator never returns Equal -> the later of two equal times wins
    tri = np.array([[[-2, 2, -2], [-2, 2, 2], [2, 2, 0]]] * 3, np.float32)
    col, hit, ti, st, tests = oracle.detect_batch(_R(oracle, [0, 4, 0], [0, -2, 0]), tri)
    assert hit[0] == 1 and ti[0] == 2 and tests == 3


def test_non_trianulated_length_check(oracle):
    import ctypes as C
    soup = np.zeros((4, 3), np.float32)
    r = _R(oracle, [0, 4, 0], [0, -2, 0])
    out = np.zeros(1, oracle.RESPONSE_DTYPE)
    rc = oracle.lib().xpo_collision_response_non_trianulated(
        r.ctypes.data_as(C.c_void_p), soup.ctypes.data_as(C.c_void_p), 4, 0, out.ctypes.data_as(C.c_void_p), None, None)
    assert rc == -1


def test_zero_movement_is_a_miss(oracle):
    # quirk B11: m == 0 -> NaN direction, roots +-inf/NaN, min seeded with NaN -> None
    tri = [[-2, 2, -2], [-2, 2, 2], [2, 2, 0]]
    assert oracle.detect(_R(oracle, [0, 2.5, 0], [0, 0, 0]), tri) is None


def test_degenerate_triangle_can_still_hit(oracle):
    # quirk B10: zero-area triangle -> NaN normal falls through every gate; vertex test still hits
    tri = [[0, 0, 0], [0, 0, 0], [0, 0, 0]]
    c = oracle.detect(_R(oracle, [0, 4, 0], [0, -8, 0]), tri)
    assert c is not None and c["time_to"] == np.float32(0.375)
