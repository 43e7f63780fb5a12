"""Row f4 (beyond the reference): collide-and-slide with the true contact point, unit horizon, ellipsoid space.
No reference exists to compare bits with; the loop is checked by geometric invariants, with the CPU oracle as the
injected detection backend (the GPU library plugs into the same loop through Scene.detect_batch)."""
import numpy as np
import pytest

from xp_physics_cuda import fauerby, scenes


def oracle_detect(oracle, tris):
    return lambda resps: oracle.detect_batch(resps, tris)[:4]


def dist_to_soup(p, tris):
    """distance of every point p[i] to the nearest triangle of the soup"""
    best = np.full(len(p), np.inf)
    for t in tris:
        cp = fauerby.closest_point_on_triangle(p, np.broadcast_to(t, (len(p), 3, 3)))
        best = np.minimum(best, np.linalg.norm(p - cp, axis=1))
    return best


def test_closest_point_regions_and_optimality():
    tri = np.array([[0, 0, 0], [4, 0, 0], [0, 3, 0]], np.float64)
    pts = np.array([[1, 1, 2], [-1, -1, 0], [5, -1, 1], [-1, 5, 0], [2, -2, 0], [-2, 1, 0], [3, 3, 0]], np.float64)
    want = np.array([[1, 1, 0], [0, 0, 0], [4, 0, 0], [0, 3, 0], [2, 0, 0], [0, 1, 0], [1.92, 1.56, 0]])
    got = fauerby.closest_point_on_triangle(pts, np.broadcast_to(tri, (len(pts), 3, 3)))
    np.testing.assert_allclose(got, want, atol=1e-9)
    rng = np.random.default_rng(0)
    tris = rng.normal(0, 2, (300, 3, 3))
    p = rng.normal(0, 3, (300, 3))
    cp = fauerby.closest_point_on_triangle(p, tris)
    d = np.linalg.norm(p - cp, axis=1)
    u, v = rng.random((2, 300, 64))
    flip = u + v > 1
    u[flip], v[flip] = 1 - u[flip], 1 - v[flip]
    samples = tris[:, None, 0] + u[..., None] * (tris[:, None, 1] - tris[:, None, 0]) + v[..., None] * (tris[:, None, 2] - tris[:, None, 0])
    assert (d[:, None] <= np.linalg.norm(p[:, None] - samples, axis=2) + 1e-9).all()        # no sampled point is closer
    # the result lies in the triangle's plane and inside it (barycentric coordinates in [0, 1])
    e1, e2, w = tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0], cp - tris[:, 0]
    nrm = np.cross(e1, e2)
    assert np.abs((w * nrm).sum(1) / np.linalg.norm(nrm, axis=1)).max() < 1e-9


def floor(n=12, y=0.0):
    h = np.full((n + 1, n + 1), y, np.float32)
    return scenes.heightmap_from_heights(h, x0=-n / 2, z0=-n / 2, spacing=1.0)


def test_slides_along_a_floor_instead_of_stopping(oracle):
    tris = floor()
    pos = np.array([[0.0, 1.5, 0.0]], np.float32)
    vel = np.array([[2.0, -1.0, 1.0]], np.float32)                   # would end at y = 0.5: half inside the floor
    p, v, handled = fauerby.collide_and_slide(None, pos, vel, triangles=tris, detect=oracle_detect(oracle, tris))
    assert handled[0] >= 1
    assert abs(p[0, 1] - 1.0) < 3 * fauerby.VERY_CLOSE               # resting on the floor
    assert p[0, 0] > 1.9 and p[0, 2] > 0.9                           # the tangential part of the step was kept
    # the reference's provisional response (slide along -m_hat) loses it
    r = oracle.collision_response_batch(oracle.make_responses(pos, vel), tris, max_iter=8)
    assert r[0]["c"][0, 0] < p[0, 0] - 0.5


def test_wall_and_corner(oracle):
    wall = np.array([[[3, -5, -5], [3, 5, -5], [3, 5, 5]], [[3, -5, -5], [3, 5, 5], [3, -5, 5]]], np.float32)   # plane x = 3
    tris = np.concatenate([floor(), wall, wall[:, ::-1]])            # both windings: a two-sided wall
    pos = np.array([[0.0, 1.0 + 0.01, 0.0]], np.float32)
    vel = np.array([[4.0, 0.0, 2.0]], np.float32)
    p, v, handled = fauerby.collide_and_slide(None, pos, vel, triangles=tris, detect=oracle_detect(oracle, tris))
    assert abs(p[0, 0] - 2.0) < 3 * fauerby.VERY_CLOSE               # one radius in front of the wall
    assert p[0, 2] > 1.9                                             # and it slid along it
    assert dist_to_soup(p.astype(np.float64), tris).min() > 1.0 - 1e-3


def test_crowd_on_bumpy_terrain_never_penetrates(oracle):
    tris, h = scenes.sine_terrain(20, 20, seed=2)
    resps = scenes.terrain_spheres(h, 300, seed=3)
    pos, vel = resps["c"].copy(), resps["movement"] * 12.0           # big steps: every sphere reaches the ground
    start = dist_to_soup(pos.astype(np.float64), tris)
    ok = start > 1.0 + 1e-3
    p, v, handled = fauerby.collide_and_slide(None, pos[ok], vel[ok], triangles=tris, detect=oracle_detect(oracle, tris), max_iter=6)
    assert (handled > 0).sum() > 100
    assert dist_to_soup(p.astype(np.float64), tris).min() > 1.0 - 2e-3
    assert (np.linalg.norm(v, axis=1)[handled < 6] == 0).all()       # whoever finished has no velocity left


def test_ellipsoid_space(oracle):
    space = fauerby.EllipsoidSpace((0.4, 0.9, 0.4))
    np.testing.assert_allclose(space.back(space.to([[1, 2, 3]])), [[1, 2, 3]], rtol=1e-6)
    with pytest.raises(ValueError):
        fauerby.EllipsoidSpace((1, 0, 1))
    tris = floor()
    e_tris = space.to(tris)
    p, v, handled = fauerby.collide_and_slide(None, space.to([[0.0, 2.0, 0.0]]), space.to([[0.5, -3.0, 0.0]]),
                                              triangles=e_tris, detect=oracle_detect(oracle, e_tris))
    world = space.back(p)
    assert abs(world[0, 1] - 0.9) < 0.02                             # the ellipsoid stands on the floor: centre one y-radius up
    assert world[0, 0] > 0.4
