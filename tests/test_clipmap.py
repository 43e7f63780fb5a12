"""Geometry clipmap as a collision scene (row f2): host mirror of game/src/graphics/clipmap.rs."""
import numpy as np

from xp_physics_cuda import clipmap as cm
from xp_physics_cuda import scenes


def test_calculate_update_range_1d_kat():
    f = cm.calculate_update_range_1d                      # clipmap.rs:1021-1029
    assert f(0, 1, 4) == range(4, 5)
    assert f(0, 6, 4) == range(6, 10)
    assert f(1, 0, 4) == range(0, 1)
    assert f(0, -1, 4) == range(-1, 0)
    assert f(0, -10, 4) == range(-10, -6)
    assert f(0, 3, 1) == range(3, 4)


def test_calculate_copy_ranges_1d_kat():
    g = cm.calculate_copy_ranges_1d                       # clipmap.rs:1031-1041
    assert g(range(3, 5), 4) == [range(0, 1), range(3, 4)]
    assert g(range(0, 4), 4) == [range(0, 4)]
    assert g(range(1, 5), 4) == [range(0, 4)]
    assert g(range(2, 5), 4) == [range(0, 1), range(2, 4)]
    assert g(range(-6, -1), 4) == [range(0, 4)]
    assert g(range(-3, -1), 4) == [range(1, 3)]
    assert g(range(-2, 1), 4) == [range(0, 1), range(2, 4)]
    assert g(range(-2, 2), 4) == [range(0, 4)]


def test_constants_and_snapping():
    assert (cm.CM_M, cm.CM_P, cm.CM_TEXTURE_SIZE, cm.BASE_OFFSET, cm.CM_INTERIOR_SIZE) == (32, 3, 128, 62, 65)
    assert cm.unit_size_for_level(3) == 16.0
    assert cm.snap_down(37.0, 0) == 36.0 and cm.snap_down(-21.0, 1) == -24.0
    assert cm.snap_down_to_index(37.0, 0) == 18 and cm.snap_down_to_index(-21.0, 0) == -12
    assert len(cm.OFFSETS_MXM) == 12


def test_store_emission_equals_direct_generation():
    c = cm.Clipmap(5)
    c.update_heightmap((37.0, -21.0), scenes.sine_height)
    assert c.texels_generated == 5 * 128 * 128
    a = c.triangles()
    b = scenes.clipmap_triangles((37.0, -21.0), scenes.sine_height)
    assert a.shape == b.shape == (127016, 3, 3)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    table = c.part_table()
    assert len(table) == 1 + 4 * 18 and table[:, 6].sum() == len(a) and table[0, 6] == 2 * 126 * 126
    assert (np.diff(table[:, 5]) == table[:-1, 6]).all()


def test_incremental_toroidal_update_equals_regeneration():
    rng = np.random.default_rng(1)
    c = cm.Clipmap(5)
    pos = np.array([37.0, -21.0])
    c.update_heightmap(tuple(pos), scenes.sine_height)
    small_moves = 0
    for step in range(14):
        pos = pos + rng.uniform(-9, 9, 2) * (1 if step % 5 else 40)          # mostly small moves, some jumps
        prev = c.center
        c.update_heightmap(tuple(pos), scenes.sine_height)
        fresh = cm.Clipmap(5)
        fresh.update_heightmap(tuple(pos), scenes.sine_height)
        assert np.array_equal(c.triangles().view(np.uint32), fresh.triangles().view(np.uint32)), step
        if step % 5:
            assert c.texels_generated < 5 * 128 * 128 // 8                    # only the rows that scrolled in
            small_moves += 1
        # the regions an incremental upload would copy cover every texel that changed
        _, regions = c.get_copy_descriptions(prev)
        covered = np.zeros_like(c.data, bool)
        for r in regions:
            covered[r.level, r.y:r.y + r.ylen, r.x:r.x + r.xlen] = True
        old = cm.Clipmap(5)
        old.update_heightmap(prev, scenes.sine_height)
        changed = old.data != c.data
        # `old` regenerated from scratch holds, outside the scrolled rows, exactly what c kept
        assert not (changed & ~covered).any(), step
    assert small_moves >= 8
    c2 = cm.Clipmap(5)
    c2.update_heightmap((5.0, 5.0), scenes.sine_height)
    c2.update_heightmap((5.5, 5.5), scenes.sine_height)                       # same base index on every level
    assert c2.texels_generated == 0


def test_range_helpers_against_a_set_model():
    """The two 1-D helpers against what they mean: the rows of the new window that the old window did not hold, and the
    same rows as texture coordinates (row mod size), for every small move and window size."""
    from xp_physics_cuda.clipmap import calculate_copy_ranges_1d, calculate_update_range_1d
    for size in (1, 2, 5, 8, 128):
        for first in range(-2 * size - 3, 2 * size + 4, max(1, size // 5)):
            for second in range(first - 2 * size - 2, first + 2 * size + 3):
                rows = calculate_update_range_1d(first, second, size)
                new, old = set(range(second, second + size)), set(range(first, first + size))
                assert set(rows) == new - old, (first, second, size)
                if first >= 0 and second >= 0:                                  # (u32 cast: non-negative rows only)
                    pieces = calculate_copy_ranges_1d(rows, size)
                    assert len(pieces) <= 2
                    texels = [t for p in pieces for t in p]
                    assert len(texels) == len(set(texels)) == len(rows)
                    assert set(texels) == {r % size for r in rows}
                    assert all(0 <= p.start < p.stop <= size for p in pieces)
