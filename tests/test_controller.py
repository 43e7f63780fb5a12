"""Row f1: the game's simulation step (simulation.rs:30-52) driven through the library.
CPU part: the host glue alone; GPU part: whole trajectories, bit-compared with the same loop on the oracle."""
import os

import numpy as np
import pytest

from xp_physics_cuda import controller as ctl
from xp_physics_cuda import scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_move_along_local_axis_matches_the_reference_convention():
    ident = np.array([1, 0, 0, 0], np.float32)
    np.testing.assert_allclose(ctl.move_along_local_axis(ident, 2.0, 0.0)[0], [0, 0, -2])      # forward is -z
    np.testing.assert_allclose(ctl.move_along_local_axis(ident, 0.0, 3.0)[0], [3, 0, 0])       # right is +x
    yaw90 = np.array([np.cos(np.pi / 4), 0, np.sin(np.pi / 4), 0], np.float32)                 # 90 deg about +y
    np.testing.assert_allclose(ctl.move_along_local_axis(yaw90, 1.0, 0.0)[0], [-1, 0, 0], atol=1e-6)


def test_client_rejects_non_monotonic_frames():
    c = ctl.Client(None, sweep=lambda r: r)
    p = ctl.Player(ctl.Pose(np.zeros(3, np.float32)))
    c.handle(5, ctl.InputState(None), p, 1 / 60)
    with pytest.raises(AssertionError):                      # simulation.rs:37
        c.handle(5, ctl.InputState(None), p, 1 / 60)


def _oracle_sweep(oracle, tris, max_iter):
    return lambda resp: oracle.collision_response_batch(resp, tris, max_iter=max_iter)[0]


def test_oracle_backed_walk_changes_height(oracle):
    """sanity of the glue with the CPU oracle as backend: a player pushed into rising terrain is moved"""
    red = np.load(os.path.join(GOLDEN, "heightmap_crop_red.npy"))[60:101, 60:101]
    tris = scenes.heightmap_from_pixels(red) * np.float32(2.0)             # unit-sphere space of an r = 0.5 player
    start = np.array([0.0, red[20, 20] / 10.0 + 0.6, 0.0], np.float32)
    c = ctl.Client(None, radius=0.5, gravity=(0, -9.81 / 60, 0), max_iter=4, sweep=_oracle_sweep(oracle, tris, 4))
    p = ctl.Player(ctl.Pose(start.copy()))
    for f in range(10):
        c.handle(f, ctl.InputState(ctl.Movement(forward=1.0, right=0.3)), p, 1 / 60)
    assert np.isfinite(p.pose.position).all() and not np.array_equal(p.pose.position, start)


@pytest.mark.gpu
def test_crowd_trajectories_match_the_oracle_bitwise(oracle):
    import xp_physics_cuda as xp
    red = np.load(os.path.join(GOLDEN, "heightmap_crop_red.npy"))
    tris = scenes.heightmap_from_pixels(red)[:: 1] * np.float32(2.0)
    rng = np.random.default_rng(3)
    n = 48
    xz = rng.uniform(-40, 40, (n, 2))
    h = red[(xz[:, 0] + 100).astype(int), (xz[:, 1] + 100).astype(int)] / 10.0
    pos = np.stack([xz[:, 0], h + 0.8, xz[:, 1]], 1).astype(np.float32)
    yaw = rng.uniform(0, 2 * np.pi, n)
    ori = np.stack([np.cos(yaw / 2), 0 * yaw, np.sin(yaw / 2), 0 * yaw], 1).astype(np.float32)
    sub = tris                                                           # 80 000 triangles
    with xp.Scene(0, sub) as s:
        gpu = ctl.Crowd(s, pos, ori, gravity=(0, -0.5, 0), max_iter=3)
        cpu = ctl.Crowd(None, pos, ori, gravity=(0, -0.5, 0), max_iter=3, sweep=_oracle_sweep(oracle, sub, 3))
        for f in range(6):
            fwd, rgt = rng.uniform(0.2, 1.0, n), rng.uniform(-0.5, 0.5, n)
            a = gpu.step(fwd, rgt, 1 / 60)
            b = cpu.step(fwd, rgt, 1 / 60)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"frame {f}"
    assert not np.array_equal(a, pos)
