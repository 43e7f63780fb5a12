"""Row f4 of SURVEY.md section 8 -- beyond the reference: the collide-and-slide response its TODO points at.

`xp_physics::sphere_triangle_calculate_response` slides along the plane through `position + m_hat` with normal
`-m_hat` (collision_response.rs:9-16, with the author's note that this is provisional), accepts hits at any
`t > 0` and only handles unit spheres.  Fauerby's algorithm ("Improved Collision detection and Response", 2003),
which the crate follows everywhere else, does three things differently; this module provides them ON TOP of the
library's sweep (the detection stays the bit-exact GPU path, the response arithmetic here is host-side numpy and
has no reference to be bit-compared with -- it is checked by geometric invariants instead):

* only contacts inside the step count (`t <= 1`; the library's `horizon="unit"` query volume serves exactly those);
* the sliding plane passes through the TRUE contact point on the triangle (face point, edge point or vertex --
  the closest point of the hit triangle to the sphere centre at the time of impact) with the normal pointing from
  it to that centre, so a sphere slides along walls and over edges instead of along its own direction of travel;
* ellipsoids: positions and velocities are divided by the radii (`EllipsoidSpace`), in which space the ellipsoid is
  the unit sphere the sweep requires (collision_detect.rs:161).

    space = EllipsoidSpace((0.4, 0.9, 0.4))
    scene = Scene(0, space.to(triangles))                     # the scene lives in ellipsoid space
    pos, vel = collide_and_slide(scene, space.to(p), space.to(v))
    p_new = space.back(pos)
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from .api import make_responses

VERY_CLOSE = np.float32(0.005)              # Fauerby's veryCloseDistance, in unit-sphere space


class EllipsoidSpace:
    """R3 <-> ellipsoid space of an ellipsoid with the given radii (a scalar = a sphere of that radius)."""

    def __init__(self, radii):
        r = np.asarray(radii, np.float32).reshape(-1)
        self.radii = np.broadcast_to(r, (3,)).astype(np.float32) if r.size in (1, 3) else None
        if self.radii is None or not (self.radii > 0).all():
            raise ValueError("radii: one or three positive numbers")

    def to(self, v):
        return (np.asarray(v, np.float32) / self.radii).astype(np.float32)

    def back(self, v):
        return (np.asarray(v, np.float32) * self.radii).astype(np.float32)


def closest_point_on_triangle(p: np.ndarray, tri: np.ndarray) -> np.ndarray:
    """Closest point of triangle tri[i] = (a, b, c) to p[i] (Ericson, Real-Time Collision Detection 5.1.5),
    vectorised: p [n,3], tri [n,3,3] -> [n,3] float64."""
    p = np.asarray(p, np.float64).reshape(-1, 3)
    t = np.asarray(tri, np.float64).reshape(-1, 3, 3)
    a, b, c = t[:, 0], t[:, 1], t[:, 2]
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = (ab * ap).sum(1), (ac * ap).sum(1)
    bp = p - b
    d3, d4 = (ab * bp).sum(1), (ac * bp).sum(1)
    cp = p - c
    d5, d6 = (ab * cp).sum(1), (ac * cp).sum(1)
    vc = d1 * d4 - d3 * d2
    vb = d5 * d2 - d1 * d6
    va = d3 * d6 - d5 * d4
    out = np.empty_like(p)
    done = np.zeros(len(p), bool)

    def put(mask, value):
        m = mask & ~done
        out[m] = value[m]
        done[m] = True

    with np.errstate(divide="ignore", invalid="ignore"):
        put((d1 <= 0) & (d2 <= 0), a)                                            # vertex a
        put((d3 >= 0) & (d4 <= d3), b)                                           # vertex b
        put((vc <= 0) & (d1 >= 0) & (d3 <= 0), a + (d1 / (d1 - d3))[:, None] * ab)          # edge ab
        put((d6 >= 0) & (d5 <= d6), c)                                           # vertex c
        put((vb <= 0) & (d2 >= 0) & (d6 <= 0), a + (d2 / (d2 - d6))[:, None] * ac)          # edge ac
        w = (d4 - d3) / ((d4 - d3) + (d5 - d6))
        put((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), b + w[:, None] * (c - b))      # edge bc
        denom = 1.0 / (va + vb + vc)
        inside = a + (vb * denom)[:, None] * ab + (vc * denom)[:, None] * ac
    bad = ~np.isfinite(inside).all(axis=1)                                       # degenerate triangle: fall back to a vertex
    inside[bad] = a[bad]
    put(np.ones(len(p), bool), inside)
    return out


def true_contacts(triangles: np.ndarray, position: np.ndarray, tri_index: np.ndarray):
    """Contact point on the hit triangle and unit contact normal (from that point to the sphere centre) for
    centres `position` [n,3] at their time of impact; `tri_index` [n] indexes `triangles` [T,3,3]."""
    tri = np.asarray(triangles, np.float32)[np.asarray(tri_index, np.int64)]
    cp = closest_point_on_triangle(position, tri)
    d = np.asarray(position, np.float64) - cp
    ln = np.linalg.norm(d, axis=1, keepdims=True)
    n = np.divide(d, ln, out=np.zeros_like(d), where=ln > 0)
    return cp, n


def collide_and_slide(scene, position, velocity, triangles: Optional[np.ndarray] = None, max_iter: int = 5,
                      detect: Optional[Callable] = None):
    """Fauerby's collideWithWorld for a batch of unit spheres (ellipsoid space): move every sphere by its
    velocity, stopping VERY_CLOSE short of the first contact inside the step and continuing with the velocity
    projected onto the sliding plane through the true contact point, up to `max_iter` times.

    scene:     a built `Scene` (its triangles in the same space); `triangles` defaults to what was uploaded.
    detect:    injectable closest-hit function responses -> (collisions, hit, tri_index, status); default
               `scene.detect_batch(responses, horizon="unit")`.  Tests run the same loop on the CPU oracle.
    Returns (final position [n,3] float32, remaining velocity [n,3] float32, contacts handled per sphere)."""
    if triangles is None:
        triangles = getattr(scene, "_keep", None)
        if triangles is None:
            raise ValueError("pass the triangles the scene was built from")
    triangles = np.asarray(triangles, np.float32).reshape(-1, 3, 3)
    if detect is None:
        def detect(resps):
            return scene.detect_batch(resps, horizon="unit")
    pos = np.array(position, np.float64).reshape(-1, 3)
    vel = np.array(velocity, np.float64).reshape(-1, 3)
    n = len(pos)
    handled = np.zeros(n, np.int32)
    active = np.linalg.norm(vel, axis=1) > VERY_CLOSE
    for _ in range(max_iter):
        idx = np.flatnonzero(active)
        if not len(idx):
            break
        resps = make_responses(pos[idx].astype(np.float32), vel[idx].astype(np.float32))
        col, hit, tri_index, status = detect(resps)
        p32, v32 = resps["c"].astype(np.float64), resps["movement"].astype(np.float64)     # what the sweep saw
        t = col["time_to"].astype(np.float64)
        hits = hit.astype(bool) & (t <= 1.0) & (status == 0)
        free = idx[~hits]
        pos[free] = p32[~hits] + v32[~hits]                                      # nothing in the way: take the step
        vel[free] = 0.0
        active[free] = False
        if not hits.any():
            continue
        h = idx[hits]
        p, v, th = p32[hits], v32[hits], t[hits]
        vlen = np.linalg.norm(v, axis=1)
        dist = vlen * th
        centre = p + v * th[:, None]                                             # sphere centre at the time of impact
        cp, _ = true_contacts(triangles, centre, tri_index[hits])
        # stop VERY_CLOSE short of the contact and pull the contact point back by the same amount
        vhat = v / vlen[:, None]
        short = np.maximum(dist - VERY_CLOSE, 0.0)
        moved = dist >= VERY_CLOSE
        new_base = np.where(moved[:, None], p + vhat * short[:, None], p)
        cp = np.where(moved[:, None], cp - vhat * VERY_CLOSE, cp)
        # sliding plane: through the contact point, normal towards the (new) centre
        nrm = new_base - cp
        nl = np.linalg.norm(nrm, axis=1, keepdims=True)
        nrm = np.divide(nrm, nl, out=np.zeros_like(nrm), where=nl > 0)
        dest = p + v
        sd = ((dest - cp) * nrm).sum(1)
        new_dest = dest - sd[:, None] * nrm
        new_vel = new_dest - cp
        pos[h] = new_base
        vel[h] = new_vel
        handled[h] += 1
        stop = np.linalg.norm(new_vel, axis=1) < VERY_CLOSE
        vel[h[stop]] = 0.0
        active[h[stop]] = False
    return pos.astype(np.float32), vel.astype(np.float32), handled
