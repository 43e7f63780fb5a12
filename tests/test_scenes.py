"""Host-side scene generators (equivalents of the reference's triangle sources)."""
import os

import numpy as np

from xp_physics_cuda import scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _normals(t):
    return np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])


def test_heightmap_crop_is_the_reference_mesh():
    red = np.load(os.path.join(GOLDEN, "heightmap_crop_red.npy"))
    tris = scenes.heightmap_from_pixels(red)
    assert tris.shape == (80000, 3, 3)                 # 200 x 200 quads x 2 (heightmap_loader.rs:14-15,19-20)
    assert tris[..., 0].min() == -100 and tris[..., 0].max() == 100
    assert tris[..., 1].max() == np.float32(red.max()) / np.float32(10.0)
    # first quad (x=0,z=0): p00,p01,p11 then p00,p11,p10 (heightmap_loader.rs:56-63)
    p00 = [-100, red[0, 0] / 10.0, -100]
    p01 = [-100, red[0, 1] / 10.0, -99]
    p11 = [-99, red[1, 1] / 10.0, -99]
    p10 = [-99, red[1, 0] / 10.0, -100]
    np.testing.assert_allclose(tris[0], [p00, p01, p11])
    np.testing.assert_allclose(tris[1], [p00, p11, p10])
    assert (_normals(tris)[:, 1] > 0).all()            # collision normals face +y


def test_sine_terrain_sizes():
    tris, h = scenes.sine_terrain(708, 707, seed=1)
    assert len(tris) == 1001112                        # config 2: T = 1 001 112
    assert h.shape == (709, 708)
    assert (_normals(tris[:5000])[:, 1] > 0).all()
    s = scenes.terrain_spheres(h, 1000, seed=1)
    assert (s["r"] == 1).all()
    assert np.linalg.norm(s["movement"], axis=1).max() < 0.74


def test_cube_and_icosphere():
    c = scenes.cube(2.0)
    assert c.shape == (12, 3, 3)
    n = _normals(c)
    centre = c.mean(axis=1)
    assert (np.einsum("ij,ij->i", n, centre) > 0).all()          # outward winding
    for sub, cnt in ((0, 20), (1, 80), (2, 320)):
        s = scenes.icosphere(1.5, sub)
        assert s.shape == (cnt, 3, 3)
        np.testing.assert_allclose(np.linalg.norm(s.reshape(-1, 3), axis=1), 1.5, rtol=1e-6)
        assert (np.einsum("ij,ij->i", _normals(s), s.mean(axis=1)) > 0).all()


def test_mesh_field():
    t, lo, hi = scenes.mesh_field(5000, seed=2)
    assert len(t) == 5000
    t2, _, _ = scenes.mesh_field(5000, seed=2, frame=3)
    assert not np.array_equal(t, t2) and t.shape == t2.shape
    s = scenes.field_spheres(lo, hi, 100)
    assert np.linalg.norm(s["movement"], axis=1).max() < 1.0


def test_clipmap_soup():
    c = scenes.clipmap_triangles((37.0, -21.0), scenes.sine_height)
    # level 0: 126^2 quads; each ring level: 12 m x m (31^2) + 2 p x m + 2 m x p (2*31) + trims 2 * 64 quads
    per_ring = 12 * 31 * 31 + 4 * 2 * 31 + 2 * 64
    assert len(c) == 2 * (126 * 126 + 4 * per_ring)
    n = _normals(c)
    assert (n[:, 1] > 0).all()                              # no zero-area stitch triangles, all facing up
    # level l vertices sit on the 2 * 2^l lattice and rings nest around the centre
    assert np.all(np.mod(c[: 2 * 126 * 126, :, [0, 2]], 2.0) == 0)
    ext = np.abs(c[:, :, [0, 2]] - np.float32([37.0, -21.0])).max()
    assert 126 * 16 - 64 < ext <= 128 * 16 + 64


def test_greedy_voxel_mesh_is_closed_and_outward():
    solid = scenes.procedural_chunk((0, 0, 0))
    g = scenes.greedy_voxel_mesh(solid)
    vol = np.einsum("ij,ij->i", g[:, 0], np.cross(g[:, 1], g[:, 2])).sum() / 6.0
    assert abs(vol - solid.sum() * 1e-3) < 1e-3             # divergence theorem: closed + outward wound
    one = np.zeros((4, 4, 4), bool)
    one[1:3, 1:3, 1:3] = True
    assert len(scenes.greedy_voxel_mesh(one)) == 12         # a 2x2x2 block merges into 6 quads
    assert len(scenes.greedy_voxel_mesh(np.zeros((3, 3, 3), bool))) == 0


def test_config4_scene():
    t = scenes.config4_scene()
    assert t.shape[1:] == (3, 3) and len(t) > 127000
    s = scenes.scene_spheres(t, 100)
    assert (s["r"] == 1).all()
