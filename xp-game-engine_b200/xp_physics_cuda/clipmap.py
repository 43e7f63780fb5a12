"""Geometry-clipmap terrain as a collision scene (clipmap half).

Mirror of `game/src/graphics/clipmap.rs` restricted to what collision needs:

* the per-level TOROIDAL height store (`Clipmap`, clipmap.rs:840-986): `levels` textures of
  (CM_N + 1)^2 heights addressed modulo the texture size, refreshed incrementally -- when the centre
  moves only the rows and columns that scroll in are regenerated (`calculate_update_range_1d`,
  `calculate_copy_ranges_1d`, `get_copy_descriptions`; their known answers are the reference's own
  tests, clipmap.rs:1021-1041);
* the part layout the renderer draws (clipmap.rs:14-68, 566-625): level 0 as the full N x N grid, every
  coarser level as a ring of twelve m x m blocks, four fix-ups and two interior trims;
* the vertex placement of `shader_clipmap.vert:43-65`: vertex = (ix * unit, heights[level][iz % S][ix % S],
  iz * unit).

`Clipmap.triangles()` emits the soup on the host from the store; `Scene.set_clipmap(clipmap)` uploads the
327 KB store and emits the same triangles on the device (`xp_scene_set_clipmap`), bit for bit.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

CM_N = 127                                   # clipmap.rs:15
CM_UNIT_SIZE_SMALLEST = np.float32(2.0)      # clipmap.rs:18
CM_M = (CM_N + 1) // 4                       # :19
CM_P = 3                                     # :20
CM_M_SIZE, CM_P_SIZE = CM_M - 1, CM_P - 1
CM_INTERIOR_SIZE = CM_M * 2 + 1              # :23
CM_TEXTURE_SIZE = CM_N + 1                   # :24
CM_1M = CM_M_SIZE
CM_2M = 2 * CM_M_SIZE
CM_2M1P = 2 * CM_M_SIZE + CM_P_SIZE
CM_3M1P = 3 * CM_M_SIZE + CM_P_SIZE
BASE_OFFSET = (CM_N - 3) // 2                # :70

OFFSETS_MXM = [[0, 0], [CM_1M, 0], [CM_2M1P, 0], [CM_3M1P, 0], [0, CM_1M], [CM_3M1P, CM_1M], [0, CM_2M1P],
               [CM_3M1P, CM_2M1P], [0, CM_3M1P], [CM_1M, CM_3M1P], [CM_2M1P, CM_3M1P], [CM_3M1P, CM_3M1P]]
OFFSETS_PXM = [[CM_2M, 0], [CM_2M, CM_3M1P]]
OFFSETS_MXP = [[0, CM_2M], [CM_3M1P, CM_2M]]


def level_factor(level: int) -> int:
    return 2 ** level                        # :788-790


def unit_size_for_level(level: int) -> np.float32:
    return np.float32(level_factor(level)) * CM_UNIT_SIZE_SMALLEST          # :792-794


def snap_down(val, level: int) -> np.float32:
    s = unit_size_for_level(level + 1)
    return np.float32(np.floor(np.float32(val) / s) * s)                    # :800-803


def snap_diff(val, level_a: int, level_b: int) -> np.float32:
    return np.float32(snap_down(val, level_a) - snap_down(val, level_b))    # :796-798


def snap_down_to_index(val, level: int) -> int:
    s = unit_size_for_level(level + 1)
    return int(np.float32(np.floor(np.float32(val) / s)) * np.float32(2.0))  # :805-808


def calculate_update_range_1d(first: int, second: int, size: int) -> range:
    """The rows that scroll in when a window of `size` rows moves from base `first` to base `second`: the new window
    [second, second + size) minus the old one [first, first + size), which is one interval because both have the same
    length (what clipmap.rs:988-996 computes; its known answers are checked in tests/test_clipmap.py)."""
    moved_up = second > first
    lo = max(second, first + size) if moved_up else second
    hi = second + size if moved_up else min(first, second + size)
    return range(lo, max(lo, hi))


def calculate_copy_ranges_1d(rng: range, size: int) -> list[range]:
    """The same rows as texture coordinates: rows live at `row mod size` (the reference casts to u32 first,
    clipmap.rs:998-1019), so an interval becomes the whole texture, one piece, or -- when it crosses the seam -- the piece
    after the seam followed by the piece before it (the reference's order)."""
    count = rng.stop - rng.start
    assert count >= 0
    if count == 0:
        return []
    if count >= size:
        return [range(0, size)]
    start = (rng.start & 0xFFFFFFFF) % size
    over = start + count - size                      # rows that wrap past the seam
    if over <= 0:
        return [range(start, start + count)] if over < 0 else [range(start, size)]
    return [range(0, over), range(start, size)]


@dataclass
class CopyDescription:                       # clipmap.rs:813-820
    offset: int
    x: int
    y: int
    xlen: int
    ylen: int
    level: int


@dataclass
class Part:
    """One drawn grid: integer offset of its first vertex from the level's base index, vertices per side."""
    level: int
    off_x: int
    off_z: int
    size_x: int
    size_z: int

    @property
    def triangles(self) -> int:
        return 2 * (self.size_x - 1) * (self.size_z - 1)


class Clipmap:
    """clipmap.rs:840-986.  `generator(x_pos, z_pos)` takes float32 arrays and returns heights."""

    def __init__(self, levels: int = 5, size: int = CM_TEXTURE_SIZE):
        self.levels, self.size = int(levels), int(size)
        self.data = np.zeros((self.levels, self.size, self.size), np.float32)      # [level][z][x]
        self.center = None
        self.texels_generated = 0            # bookkeeping for tests / callers: work done by the last update

    def get_height(self, x: int, z: int, level: int) -> np.float32:
        return self.data[level, z, x]

    def _update(self, xs: range, zs: range, level: int, generator) -> None:
        if len(xs) == 0 or len(zs) == 0:
            return
        unit = unit_size_for_level(level)
        x = np.arange(xs.start, xs.stop, dtype=np.int64)
        z = np.arange(zs.start, zs.stop, dtype=np.int64)
        x_pos = x.astype(np.float32) * unit
        z_pos = z.astype(np.float32) * unit
        Z, X = np.meshgrid(z_pos, x_pos, indexing="ij")
        h = np.asarray(generator(X, Z), np.float32)
        self.data[level][np.ix_(z % self.size, x % self.size)] = h          # `as u32 % CM_TEXTURE_SIZE`, a power of two
        self.texels_generated += h.size

    def update_heightmap(self, center, generator) -> None:
        """clipmap.rs:957-985.  First call: every level in full; later calls: only the rows (`update_xrows`,
        :861-879) and columns (`update_zrows`, :881-899) that scrolled in."""
        self.texels_generated = 0
        for level in range(self.levels):
            bx = snap_down_to_index(center[0], level) - BASE_OFFSET
            bz = snap_down_to_index(center[1], level) - BASE_OFFSET
            if self.center is not None:
                px = snap_down_to_index(self.center[0], level) - BASE_OFFSET
                pz = snap_down_to_index(self.center[1], level) - BASE_OFFSET
                if (px, pz) != (bx, bz):
                    xrows = calculate_update_range_1d(px, bx, self.size)
                    zrows = calculate_update_range_1d(pz, bz, self.size)
                    self._update(xrows, range(bz, bz + self.size), level, generator)
                    self._update(range(bx, bx + self.size), zrows, level, generator)
            else:
                self._update(range(bx, bx + self.size), range(bz, bz + self.size), level, generator)
        self.center = (np.float32(center[0]), np.float32(center[1]))

    def get_copy_descriptions(self, previous):
        """clipmap.rs:901-955: the texture regions an incremental upload has to refresh."""
        out: list[CopyDescription] = []
        for level in range(self.levels):
            if self.center is not None and previous is not None:
                px = snap_down_to_index(previous[0], level) - BASE_OFFSET
                pz = snap_down_to_index(previous[1], level) - BASE_OFFSET
                bx = snap_down_to_index(self.center[0], level) - BASE_OFFSET
                bz = snap_down_to_index(self.center[1], level) - BASE_OFFSET
                if (px, pz) != (bx, bz):
                    for r in calculate_copy_ranges_1d(calculate_update_range_1d(px, bx, self.size), self.size):
                        out.append(CopyDescription(0, r.start, 0, r.stop - r.start, self.size, level))
                    for r in calculate_copy_ranges_1d(calculate_update_range_1d(pz, bz, self.size), self.size):
                        out.append(CopyDescription(0, 0, r.start, self.size, r.stop - r.start, level))
            else:
                out.append(CopyDescription(0, 0, 0, self.size, self.size, level))
        return self.center, out

    # -- geometry ------------------------------------------------------------------------------
    def parts(self) -> list[Part]:
        """The grids drawn around the current centre (clipmap.rs:566-625), without the zero-area
        stitching strips (`create_degenerates_*`, render-only)."""
        if self.center is None:
            raise RuntimeError("update_heightmap first")
        cx, cz = self.center
        out: list[Part] = []
        for level in range(self.levels):
            if level == 0:
                out.append(Part(0, 0, 0, CM_N, CM_N))
                continue
            out += [Part(level, o[0], o[1], CM_M, CM_M) for o in OFFSETS_MXM]
            out += [Part(level, o[0], o[1], CM_P, CM_M) for o in OFFSETS_PXM]
            out += [Part(level, o[0], o[1], CM_M, CM_P) for o in OFFSETS_MXP]
            dz = snap_diff(cz, level - 1, level)
            dx = snap_diff(cx, level - 1, level)
            out.append(Part(level, CM_1M, CM_1M if dz > 1e-7 else CM_3M1P - 1, CM_INTERIOR_SIZE, 2))     # h_top / h_bottom
            out.append(Part(level, CM_1M if dx > 1e-7 else CM_3M1P - 1, CM_1M, 2, CM_INTERIOR_SIZE))     # v_left / v_right
        return out

    def base_index(self, level: int):
        return (snap_down_to_index(self.center[0], level) - BASE_OFFSET,
                snap_down_to_index(self.center[1], level) - BASE_OFFSET)

    def part_table(self) -> np.ndarray:
        """int32 [n_parts, 8]: level, first vertex index x / z (absolute), vertices per side x / z, first
        triangle, triangles, 0 -- what `xp_scene_set_clipmap` takes."""
        rows, first = [], 0
        for p in self.parts():
            bx, bz = self.base_index(p.level)
            rows.append([p.level, bx + p.off_x, bz + p.off_z, p.size_x, p.size_z, first, p.triangles, 0])
            first += p.triangles
        return np.asarray(rows, np.int32)

    def triangles(self) -> np.ndarray:
        """Host emission from the toroidal store: per part all [i0, i2, i1] triangles of `create_grid`
        (clipmap.rs:752-786, quads with x fastest), then all [i3, i1, i2]."""
        out = []
        for level, x0, z0, sx, sz, _, _, _ in self.part_table():
            unit = unit_size_for_level(int(level))
            x, z = np.meshgrid(np.arange(sx - 1), np.arange(sz - 1), indexing="xy")
            x, z = x.reshape(-1) + x0, z.reshape(-1) + z0
            i0, i1, i2, i3 = (x, z), (x + 1, z), (x, z + 1), (x + 1, z + 1)

            def vert(i):
                ix, iz = i
                h = self.data[level, iz % self.size, ix % self.size]
                return np.stack([ix.astype(np.float32) * unit, h, iz.astype(np.float32) * unit], axis=-1)
            out.append(np.stack([vert(i0), vert(i2), vert(i1)], axis=1))
            out.append(np.stack([vert(i3), vert(i1), vert(i2)], axis=1))
        return np.ascontiguousarray(np.concatenate(out), dtype=np.float32)
