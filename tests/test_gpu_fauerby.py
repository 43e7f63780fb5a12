"""Row f4 on the DEVICE (xp_collide_and_slide_batch: unit horizon, sliding plane through the true contact point,
ellipsoid space): the geometric invariants of tests/test_fauerby.py with the GPU library as the backend, and
agreement with the float64 host form of the same loop (fauerby.collide_and_slide) running on the library's sweep."""
import numpy as np
import pytest

import xp_physics_cuda as xp
from tests.test_fauerby import dist_to_soup, floor
from xp_physics_cuda import fauerby, scenes

pytestmark = pytest.mark.gpu


def test_device_slide_along_a_floor_and_a_wall():
    wall = np.array([[[3, -5, -5], [3, 5, -5], [3, 5, 5]], [[3, -5, -5], [3, 5, 5], [3, -5, 5]]], np.float32)   # plane x = 3
    tris = np.concatenate([floor(), wall, wall[:, ::-1]])
    pos = np.array([[0.0, 1.5, 0.0], [0.0, 1.01, 0.0], [0.0, 5.0, 0.0], [1.0, 3.0, 1.0]], np.float32)
    vel = np.array([[2.0, -1.0, 1.0], [4.0, 0.0, 2.0], [0.5, 0.25, 0.0], [0.0, 0.001, 0.0]], np.float32)
    with xp.Scene(0, tris) as s:
        p, v, handled = s.collide_and_slide_batch(pos, vel)
    assert handled[0] >= 1 and abs(p[0, 1] - 1.0) < 3 * fauerby.VERY_CLOSE        # resting on the floor ...
    assert p[0, 0] > 1.9 and p[0, 2] > 0.9                                        # ... having kept the tangential part of the step
    assert abs(p[1, 0] - 2.0) < 3 * fauerby.VERY_CLOSE and p[1, 2] > 1.9          # one radius in front of the wall, slid along it
    assert handled[2] == 0 and np.allclose(p[2], pos[2] + vel[2]) and not v[2].any()   # nothing in the way: the whole step
    assert handled[3] == 0 and np.array_equal(p[3], pos[3]) and np.array_equal(v[3], vel[3])   # (almost) no velocity: untouched
    assert dist_to_soup(p[:2].astype(np.float64), tris).min() > 1.0 - 1e-3


def test_device_crowd_matches_the_host_form_and_never_penetrates():
    tris, h = scenes.sine_terrain(20, 20, seed=2)
    resps = scenes.terrain_spheres(h, 4000, seed=3)
    pos, vel = resps["c"].copy(), resps["movement"] * 12.0
    with xp.Scene(0) as s:
        s.set_heightmap(h)                                                        # grid walk under the slide loop
        s.build()
        p, v, handled = s.collide_and_slide_batch(pos, vel, max_iter=6)
        hp, hv, hh = fauerby.collide_and_slide(s, pos, vel, triangles=tris, max_iter=6)   # float64 host loop on the library's sweep
    with xp.Scene(0, tris) as s2:                                                 # LBVH walk under the slide loop
        p2, v2, handled2 = s2.collide_and_slide_batch(pos, vel, max_iter=6)
    assert np.array_equal(p, p2) and np.array_equal(v, v2) and np.array_equal(handled, handled2)
    assert (handled > 0).sum() > 2000
    same = handled == hh                                                          # fp32 vs float64 may flip a contact at a threshold
    assert same.mean() > 0.98
    assert np.abs(p[same] - hp[same]).max() < 2e-3 and np.abs(v[same] - hv[same]).max() < 2e-3
    sample = np.arange(0, len(p), 13)
    start_ok = dist_to_soup(pos[sample].astype(np.float64), tris) > 1.0 + 1e-3
    assert dist_to_soup(p[sample][start_ok].astype(np.float64), tris).min() > 1.0 - 2e-3
    assert (np.linalg.norm(v, axis=1)[handled < 6] == 0).all()                    # whoever finished has no velocity left


def test_device_ellipsoid_space():
    radii = (0.4, 0.9, 0.4)
    space = fauerby.EllipsoidSpace(radii)
    tris = floor()
    with xp.Scene(0, space.to(tris)) as s:                                        # the scene lives in ellipsoid space
        world, vel, handled = s.collide_and_slide_batch([[0.0, 2.0, 0.0]], [[0.5, -3.0, 0.0]], radii=radii)
    assert handled[0] >= 1
    assert abs(world[0, 1] - 0.9) < 0.02                                          # standing on the floor: centre one y-radius up
    assert world[0, 0] > 0.4
    with xp.Scene(0, tris) as s, pytest.raises(xp.XpError):
        s.collide_and_slide_batch([[0, 2, 0]], [[0, -1, 0]], radii=(1, 0, 1))
