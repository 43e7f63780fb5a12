"""Triangle-soup and sphere-batch generators for the benchmark configs.

These produce soups EQUIVALENT to the reference's geometry sources (same vertex placement and
winding), written from scratch in numpy -- host-side input generation, not part of the hot path:
    heightmap_from_pixels   low_poly_nice_graphics/src/world/heightmap_loader.rs:12-63
    heightmap_grid          low_poly_nice_graphics/src/mesh/mesh.rs:45-130 (Plane) winding
    cube / icosphere        low_poly_nice_graphics/src/mesh/mesh.rs:142-209 / :225-336
    mesh_field              config 3: a jittered lattice of cubes and icospheres
Everything is in unit-sphere space (the swept sphere has r == 1, collision_detect.rs:161).
"""
from __future__ import annotations

import numpy as np

from .api import make_responses


# ---------------------------------------------------------------------------------------------
# terrain
# ---------------------------------------------------------------------------------------------
def _quads_to_tris(p00, p01, p11, p10) -> np.ndarray:
    """Two triangles per quad in the reference's order (p00,p01,p11),(p00,p11,p10); normals face +y."""
    t0 = np.stack([p00, p01, p11], axis=-2)
    t1 = np.stack([p00, p11, p10], axis=-2)
    tris = np.stack([t0, t1], axis=-3)                 # (..., 2, 3, 3)
    return np.ascontiguousarray(tris.reshape(-1, 3, 3), dtype=np.float32)


def heightmap_from_heights(h: np.ndarray, x0: float = 0.0, z0: float = 0.0, spacing: float = 1.0) -> np.ndarray:
    """h[(nx+1),(nz+1)] vertex heights -> 2*nx*nz triangles, x outer / z inner like the reference loops."""
    h = np.asarray(h, np.float32)
    nx, nz = h.shape[0] - 1, h.shape[1] - 1
    xs = (x0 + spacing * np.arange(nx + 1, dtype=np.float32)).astype(np.float32)
    zs = (z0 + spacing * np.arange(nz + 1, dtype=np.float32)).astype(np.float32)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    P = np.stack([X, h, Z], axis=-1)                   # (nx+1, nz+1, 3)
    return _quads_to_tris(P[:-1, :-1], P[:-1, 1:], P[1:, 1:], P[1:, :-1])


def heightmap_from_pixels(red: np.ndarray, height_divisor: float = 10.0) -> np.ndarray:
    """heightmap_loader.rs:12-63 on an already-cropped red channel red[x, z] of shape (w+1, h+1):
    vertex (x - w/2, red/10, z - w/2), 200x200 quads -> 80 000 triangles for the reference crop."""
    red = np.asarray(red)
    w = red.shape[0] - 1
    h = red.astype(np.float32) / np.float32(height_divisor)
    return heightmap_from_heights(h, x0=-(w // 2), z0=-(w // 2), spacing=1.0)


def _hash_noise(ix: np.ndarray, iz: np.ndarray, seed: int) -> np.ndarray:
    """Deterministic per-vertex noise in [0,1) (integer hash; no RNG state, any grid size)."""
    with np.errstate(over="ignore"):
        a = (ix.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ (iz.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F))
        a ^= np.uint64((seed * 0x165667B19E3779F9) & 0xFFFFFFFFFFFFFFFF)
        a ^= a >> np.uint64(29)
        a *= np.uint64(0xBF58476D1CE4E5B9)
        a ^= a >> np.uint64(32)
    return ((a >> np.uint64(40)).astype(np.float64) / float(1 << 24)).astype(np.float32)


def sine_terrain_heights(nx: int, nz: int, seed: int = 1, noise: float = 0.5) -> np.ndarray:
    """Config-2 height field: 2 sin(x/4) + 2 sin(z/4) + noise*hash (cf. game/src/terrain/sine.rs:6-8)."""
    ix, iz = np.meshgrid(np.arange(nx + 1), np.arange(nz + 1), indexing="ij")
    h = 2.0 * np.sin(ix / 4.0) + 2.0 * np.sin(iz / 4.0)
    h = h.astype(np.float32) + np.float32(noise) * _hash_noise(ix, iz, seed)
    return h.astype(np.float32)


def sine_terrain(nx: int, nz: int, seed: int = 1, noise: float = 0.5):
    """Returns (triangles, heights); grid spans [0,nx] x [0,nz], spacing 1, 2*nx*nz triangles."""
    h = sine_terrain_heights(nx, nz, seed, noise)
    return heightmap_from_heights(h), h


def terrain_spheres(heights: np.ndarray, n: int, seed: int = 1) -> np.ndarray:
    """Config-2 sphere batch: c.xz ~ U(grid), c.y = h(c.xz) + U(1.05, 6),
    m = (U(-.3,.3), U(-.6,-.05), U(-.3,.3)); |m| <= 0.73 keeps collision_response.rs:22 quiet."""
    rng = np.random.default_rng(seed)
    nx, nz = heights.shape[0] - 1, heights.shape[1] - 1
    x = rng.uniform(0.0, nx, n).astype(np.float32)
    z = rng.uniform(0.0, nz, n).astype(np.float32)
    hx = np.minimum(x.astype(np.int64), nx - 1)
    hz = np.minimum(z.astype(np.int64), nz - 1)
    hmax = np.maximum.reduce([heights[hx, hz], heights[hx + 1, hz], heights[hx, hz + 1], heights[hx + 1, hz + 1]])
    y = (hmax + rng.uniform(1.05, 6.0, n)).astype(np.float32)
    m = np.stack([rng.uniform(-0.3, 0.3, n), rng.uniform(-0.6, -0.05, n), rng.uniform(-0.3, 0.3, n)], axis=1)
    return make_responses(np.stack([x, y, z], axis=1), m.astype(np.float32))


# ---------------------------------------------------------------------------------------------
# procedural meshes
# ---------------------------------------------------------------------------------------------
def cube(size: float = 1.0) -> np.ndarray:
    """12 outward-wound triangles, vertex order of mesh.rs:155-203."""
    a, b = -size / 2.0, size / 2.0
    v = [
        [a, b, a], [a, b, b], [b, b, a], [b, b, a], [a, b, b], [b, b, b],      # top
        [a, a, a], [b, a, a], [a, a, b], [b, a, a], [b, a, b], [a, a, b],      # bottom
        [b, a, a], [b, b, a], [b, a, b], [b, b, a], [b, b, b], [b, a, b],      # right
        [a, a, b], [a, b, b], [a, a, a], [a, b, b], [a, b, a], [a, a, a],      # left
        [a, a, b], [b, a, b], [a, b, b], [b, a, b], [b, b, b], [a, b, b],      # front
        [a, b, a], [b, b, a], [a, a, a], [b, b, a], [b, a, a], [a, a, a],      # back
    ]
    return np.asarray(v, np.float32).reshape(12, 3, 3)


def icosphere(radius: float = 1.0, subdivisions: int = 1) -> np.ndarray:
    """20 * 4^subdivisions triangles (mesh.rs:261-336): icosahedron, midpoint subdivision on the unit sphere."""
    X, Z, N = np.float32(0.525731112119133606), np.float32(0.850650808352039932), np.float32(0.0)
    pts = [[-X, N, Z], [X, N, Z], [-X, N, -Z], [X, N, -Z], [N, Z, X], [N, Z, -X], [N, -Z, X], [N, -Z, -X],
           [Z, X, N], [-Z, X, N], [Z, -X, N], [-Z, -X, N]]
    pts = [np.asarray(p, np.float32) for p in pts]
    tris = [[0, 1, 4], [0, 4, 9], [9, 4, 5], [4, 8, 5], [4, 1, 8], [8, 1, 10], [8, 10, 3], [5, 8, 3], [5, 3, 2],
            [2, 3, 7], [7, 3, 10], [7, 10, 6], [7, 6, 11], [11, 6, 0], [0, 6, 1], [6, 10, 1], [9, 11, 0],
            [9, 2, 11], [9, 5, 2], [7, 11, 2]]
    for _ in range(subdivisions):
        lookup = {}

        def mid(i, j):
            key = (i, j) if i < j else (j, i)
            if key not in lookup:
                s = pts[i] + pts[j]
                lookup[key] = len(pts)
                pts.append((s / np.float32(np.sqrt(np.float32(np.dot(s, s))))).astype(np.float32))
            return lookup[key]

        nxt = []
        for a, b, c in tris:
            m0, m1, m2 = mid(a, b), mid(b, c), mid(c, a)
            nxt += [[a, m0, m2], [b, m1, m0], [c, m2, m1], [m0, m1, m2]]
        tris = nxt
    P = np.stack(pts) * np.float32(radius)
    return np.ascontiguousarray(P[np.asarray(tris)], dtype=np.float32)


def mesh_field(n_triangles: int, seed: int = 2, spacing: float = 6.0, frame: int = 0):
    """Config 3: cubes (12 tris) and subdivision-2 icospheres (320 tris) on a jittered 3-D lattice until
    n_triangles is reached.  `frame` rigidly translates every mesh (forces an LBVH rebuild per frame).
    Returns (triangles, aabb_lo, aabb_hi)."""
    rng = np.random.default_rng(seed)
    cub, ico = cube(2.0), icosphere(1.5, 2)
    per_pair = len(cub) + len(ico)
    n_pairs = max(1, int(np.ceil(n_triangles / per_pair)))
    n_mesh = 2 * n_pairs
    side = int(np.ceil(n_mesh ** (1.0 / 3.0)))
    idx = np.arange(n_mesh)
    cell = np.stack([idx // (side * side), (idx // side) % side, idx % side], axis=1).astype(np.float32)
    centres = cell * np.float32(spacing) + rng.uniform(-1.0, 1.0, (n_mesh, 3)).astype(np.float32)
    drift = rng.uniform(-0.05, 0.05, (n_mesh, 3)).astype(np.float32)
    centres = (centres + np.float32(frame) * drift).astype(np.float32)
    c_cub = centres[0::2][:, None, None, :]
    c_ico = centres[1::2][:, None, None, :]
    tris = np.concatenate([(cub[None] + c_cub).reshape(-1, 3, 3), (ico[None] + c_ico).reshape(-1, 3, 3)])
    tris = np.ascontiguousarray(tris[:n_triangles] if len(tris) > n_triangles else tris, dtype=np.float32)
    lo = tris.reshape(-1, 3).min(axis=0) - 2.0
    hi = tris.reshape(-1, 3).max(axis=0) + 2.0
    return tris, lo.astype(np.float32), hi.astype(np.float32)


def field_spheres(lo, hi, n: int, seed: int = 2) -> np.ndarray:
    """Config-3 sphere batch: centres uniform in the field's AABB, movement uniform in the unit ball."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    m = d * (rng.uniform(0.0, 1.0, (n, 1)) ** (1.0 / 3.0)) * 0.999
    return make_responses(c, m.astype(np.float32))


# ---------------------------------------------------------------------------------------------
# clipmap terrain (config 4): the triangles the game's clipmap vertex shader places
# ---------------------------------------------------------------------------------------------
CM_N, CM_M, CM_P = 127, 32, 3                    # game/src/graphics/clipmap.rs:14-21
CM_UNIT = 2.0                                    # CM_UNIT_SIZE_SMALLEST
CM_BASE_OFFSET = (CM_N - 3) // 2                 # clipmap.rs:70


def _cm_unit(level: int) -> float:
    return float(2 ** level) * CM_UNIT           # clipmap.rs:793-795


def _cm_snap_down(val: float, level: int) -> float:
    s = _cm_unit(level + 1)
    return float(np.floor(val / s) * s)          # clipmap.rs:801-804


def _cm_grid_quads(size_x: int, size_z: int):
    """create_grid (clipmap.rs:752-786): per quad vertices i0=(x,z) i1=(x+1,z) i2=(x,z+1) i3=(x+1,z+1),
    triangles [i0,i2,i1] and [i3,i1,i2]."""
    x, z = np.meshgrid(np.arange(size_x - 1), np.arange(size_z - 1), indexing="xy")
    x, z = x.reshape(-1), z.reshape(-1)
    i0, i1 = np.stack([x, z], 1), np.stack([x + 1, z], 1)
    i2, i3 = np.stack([x, z + 1], 1), np.stack([x + 1, z + 1], 1)
    return np.concatenate([np.stack([i0, i2, i1], 1), np.stack([i3, i1, i2], 1)])       # (2q, 3, 2) integer offsets


def clipmap_triangles(center_xz, height_fn, levels: int = 5) -> np.ndarray:
    """Collision soup of the clipmap around `center_xz`: level 0 as the full N x N grid, levels 1..levels-1 as
    rings of 12 m x m blocks, 2 p x m, 2 m x p fix-ups and the two interior trims chosen like
    clipmap.rs:566-625; vertex = (ix * unit, h, iz * unit) as in shader_clipmap.vert:43-65.  The zero-area
    stitching strips (create_degenerates_*, clipmap.rs:664-750) are render-only and left out."""
    m1 = CM_M - 1
    off_mxm = [[0, 0], [m1, 0], [2 * m1 + 2, 0], [3 * m1 + 2, 0], [0, m1], [3 * m1 + 2, m1], [0, 2 * m1 + 2],
               [3 * m1 + 2, 2 * m1 + 2], [0, 3 * m1 + 2], [m1, 3 * m1 + 2], [2 * m1 + 2, 3 * m1 + 2], [3 * m1 + 2, 3 * m1 + 2]]
    off_pxm = [[2 * m1, 0], [2 * m1, 3 * m1 + 2]]
    off_mxp = [[0, 2 * m1], [3 * m1 + 2, 2 * m1]]
    g_mxm, g_pxm, g_mxp = _cm_grid_quads(CM_M, CM_M), _cm_grid_quads(CM_P, CM_M), _cm_grid_quads(CM_M, CM_P)
    g_nxn = _cm_grid_quads(CM_N, CM_N)
    g_ih, g_iv = _cm_grid_quads(2 * CM_M + 1, 2), _cm_grid_quads(2, 2 * CM_M + 1)
    cx, cz = float(center_xz[0]), float(center_xz[1])
    out = []
    for level in range(levels):
        unit = _cm_unit(level)
        snap = unit * 2.0
        centre = np.array([int(np.floor(cx / snap) * 2.0), int(np.floor(cz / snap) * 2.0)]) - CM_BASE_OFFSET
        parts = []
        if level == 0:
            parts.append(g_nxn)
        else:
            parts += [g_mxm + np.array(o) for o in off_mxm]
            parts += [g_pxm + np.array(o) for o in off_pxm]
            parts += [g_mxp + np.array(o) for o in off_mxp]
            dz = _cm_snap_down(cz, level - 1) - _cm_snap_down(cz, level)
            dx = _cm_snap_down(cx, level - 1) - _cm_snap_down(cx, level)
            parts.append(g_ih + np.array([m1, m1] if dz > 1e-7 else [m1, 3 * m1 + 1]))        # h_top / h_bottom
            parts.append(g_iv + np.array([m1, m1] if dx > 1e-7 else [3 * m1 + 1, m1]))        # v_left / v_right
        idx = np.concatenate(parts) + centre                                                   # (t, 3, 2)
        x = idx[..., 0].astype(np.float32) * np.float32(unit)
        z = idx[..., 1].astype(np.float32) * np.float32(unit)
        y = height_fn(x, z).astype(np.float32)
        out.append(np.stack([x, y, z], axis=-1))
    return np.ascontiguousarray(np.concatenate(out), dtype=np.float32)


def sine_height(x, z):
    """game/src/terrain/sine.rs:6-8."""
    return np.sin(x / 4.0) + np.sin(z / 4.0)


# ---------------------------------------------------------------------------------------------
# voxel chunks (config 4): greedy-meshed 32^3 chunks, voxel 0.1 (low_poly_nice_graphics world.rs)
# ---------------------------------------------------------------------------------------------
def procedural_chunk(chunk, size: int = 32, voxel: float = 0.1) -> np.ndarray:
    """bool[size,size,size] fill of world.rs:146-157: y_w > -5 and sin(x_w) * sin(z_w) > y_w."""
    i = np.arange(size, dtype=np.float32)
    xw = np.float32(chunk[0] * size * voxel) + i * np.float32(voxel)
    yw = np.float32(chunk[1] * size * voxel) + i * np.float32(voxel)
    zw = np.float32(chunk[2] * size * voxel) + i * np.float32(voxel)
    X, Y, Z = np.meshgrid(xw, yw, zw, indexing="ij")
    return (Y > -5.0) & (np.sin(X) * np.sin(Z) > Y)


def greedy_voxel_mesh(solid: np.ndarray, origin=(0.0, 0.0, 0.0), voxel: float = 0.1) -> np.ndarray:
    """Greedy quads over the boundary faces of a boolean voxel block, 2 outward-wound triangles per quad
    (an equivalent of world.rs:234-365: same surface, same merge rule -- widest run first, then extend rows)."""
    solid = np.asarray(solid, bool)
    tris = []
    org = np.asarray(origin, np.float32)
    for u in range(3):                               # axis normal to the face
        v, w = (u + 1) % 3, (u + 2) % 3
        for sign in (1, -1):
            shifted = np.zeros_like(solid)
            src = [slice(None)] * 3
            dst = [slice(None)] * 3
            if sign == 1:
                src[u], dst[u] = slice(1, None), slice(0, -1)
            else:
                src[u], dst[u] = slice(0, -1), slice(1, None)
            shifted[tuple(dst)] = solid[tuple(src)]
            exposed = solid & ~shifted               # faces looking along +u (sign=1) or -u
            for s in range(solid.shape[u]):
                sl = [slice(None)] * 3
                sl[u] = s
                mask = exposed[tuple(sl)]
                if u == 1:                           # keep (v, w) = (z, x) orientation consistent
                    mask = mask.T
                mask = mask.copy()
                nv, nw = mask.shape
                for b in range(nw):
                    a = 0
                    while a < nv:
                        if not mask[a, b]:
                            a += 1
                            continue
                        wdt = 1
                        while a + wdt < nv and mask[a + wdt, b]:
                            wdt += 1
                        hgt = 1
                        while b + hgt < nw and mask[a:a + wdt, b + hgt].all():
                            hgt += 1
                        mask[a:a + wdt, b:b + hgt] = False
                        p = np.zeros((4, 3), np.float32)
                        plane = s + (1 if sign == 1 else 0)
                        for k, (da, db) in enumerate(((0, 0), (wdt, 0), (wdt, hgt), (0, hgt))):
                            p[k, u] = plane
                            p[k, v] = a + da
                            p[k, w] = b + db
                        p = org + p * np.float32(voxel)
                        quad = (p[0], p[1], p[2], p[3]) if sign == 1 else (p[0], p[3], p[2], p[1])
                        tris.append([quad[0], quad[1], quad[2]])
                        tris.append([quad[0], quad[2], quad[3]])
                        a += wdt
    if not tris:
        return np.zeros((0, 3, 3), np.float32)
    return np.ascontiguousarray(np.asarray(tris, np.float32))


def config4_scene(center_xz=(37.0, -21.0), chunks=((0, -1, 0), (1, -1, 0), (0, -1, 1), (0, 0, 0)), radius: float = 0.5):
    """Clipmap terrain + greedy-meshed procedural voxel chunks, scaled by 1/radius into unit-sphere space
    (the player sphere has r = 0.5 in low_poly_nice_graphics/src/main.rs:122)."""
    parts = [clipmap_triangles(center_xz, sine_height)]
    for c in chunks:
        origin = (c[0] * 3.2 + center_xz[0], c[1] * 3.2 + 4.0, c[2] * 3.2 + center_xz[1])
        parts.append(greedy_voxel_mesh(procedural_chunk(c), origin=origin))
    tris = np.concatenate(parts) / np.float32(radius)
    return np.ascontiguousarray(tris, dtype=np.float32)


def scene_spheres(tris: np.ndarray, n: int, seed: int = 3, extent: float = 60.0) -> np.ndarray:
    """Spheres dropped from above the central part of a scene (config 4 batch)."""
    rng = np.random.default_rng(seed)
    mid = 0.5 * (tris.reshape(-1, 3).min(axis=0) + tris.reshape(-1, 3).max(axis=0))
    top = tris.reshape(-1, 3)[:, 1].max()
    c = np.stack([mid[0] + rng.uniform(-extent, extent, n), top + rng.uniform(1.05, 4.0, n),
                  mid[2] + rng.uniform(-extent, extent, n)], axis=1).astype(np.float32)
    m = np.stack([rng.uniform(-0.3, 0.3, n), rng.uniform(-0.6, -0.05, n), rng.uniform(-0.3, 0.3, n)], axis=1)
    return make_responses(c, m.astype(np.float32))
