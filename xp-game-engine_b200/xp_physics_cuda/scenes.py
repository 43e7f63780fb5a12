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
