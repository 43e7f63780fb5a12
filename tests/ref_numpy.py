"""TEST-ONLY: a second, independent restatement of the reference's algorithm, in numpy float32, written from
the Rust sources again (not from oracle/xp_oracle.c) and vectorised over spheres.  It exists to cross-check the
C oracle -- which is the definition of parity wherever the reference has no test of its own -- against a
reading of the same sources in a different language and a different shape (masks instead of early returns).
Every arithmetic step is a float32 numpy operation (IEEE, unfused), in the reference's order.

Sources: xp_physics/src/collision_detect.rs:1-213, collision_response.rs:6-35, triangle.rs:10-54, lib.rs:29-62,
collision.rs:3, xp_math/src/lib.rs:3-12; nalgebra-glm 0.8: dot = (a0 b0 + a1 b1) + a2 b2, length = sqrt(dot),
normalize = component-wise divide, triangle_normal(p1, p2, p3) = normalize(cross(p2 - p1, p3 - p1)).
"""
import numpy as np

F = np.float32
DISTANCE_EPSILON = F(0.001)


def dot(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def length(a):
    return np.sqrt(dot(a, a))


def normalize(a):
    return a / length(a)[..., None]


def cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def plane_constant(p, n):                                        # triangle.rs:34-38
    return -((n[..., 0] * p[..., 0] + n[..., 1] * p[..., 1]) + n[..., 2] * p[..., 2])


def get_roots(a, b, c):                                          # xp_math/src/lib.rs:3-12
    det = b * b - F(4.0) * a * c
    some = ~(det < 0)
    sq = np.sqrt(det)
    two_a = F(2.0) * a
    return some, (-b - sq) / two_a, (-b + sq) / two_a


def point_in_triangle(p0, p1, p2, p):                            # triangle.rs:11-32
    v0, v1, v2 = p2 - p0, p1 - p0, p - p0
    d00, d01, d02, d11, d12 = dot(v0, v0), dot(v0, v1), dot(v0, v2), dot(v1, v1), dot(v1, v2)
    inv = F(1.0) / (d00 * d11 - d01 * d01)
    u = (d11 * d02 - d01 * d12) * inv
    v = (d00 * d12 - d01 * d02) * inv
    return (u >= 0) & (v >= 0) & (u + v <= F(1.0))


class RunningMin:
    """`min(vals: &[f32])` (collision_detect.rs:10-21) over a Vec that is only ever appended to: seeded with the
    first element, replaced on strict `<`."""

    def __init__(self, n):
        self.have = np.zeros(n, bool)
        self.cur = np.zeros(n, np.float32)

    def push(self, x, on):
        first = on & ~self.have
        self.cur = np.where(first, x, self.cur)
        self.have = self.have | on
        with np.errstate(invalid="ignore"):
            better = on & ~first & (x < self.cur)
        self.cur = np.where(better, x, self.cur)


def detect(c, m, tri):
    """sphere_triangle_detect_collision for N (sphere, triangle) pairs: c, m [N,3], tri [N,3,3] ->
    (some [N] bool, t [N] float32).  r == 1 is the caller's business."""
    c, m, tri = np.asarray(c, np.float32), np.asarray(m, np.float32), np.asarray(tri, np.float32)
    n = len(c)
    v0, v1, v2 = tri[:, 0], tri[:, 1], tri[:, 2]
    with np.errstate(all="ignore"):
        tnormal = normalize(cross(v1 - v0, v2 - v0))             # Triangle::normal (triangle.rs:44-46)
        normal = normalize(tnormal)                              # :162
        nm = normalize(m)                                        # :163
        alive = ~(dot(normal, nm) > 0)                           # :166-168
        pndm = dot(normal, m)                                    # :170
        sd = dot(normal, c) + plane_constant(v0, normalize(tnormal))     # :171, triangle.rs:47-50
        asd = np.abs(sd)
        alive &= ~((pndm == 0) & (asd >= F(1.0)))                # :175-177
        mlen = length(m)                                         # :179
        msl = mlen * mlen                                        # :180
        # face (:184-196 -> :23-42)
        try_face = alive & (pndm != 0) & (asd >= F(1.0))
        t0 = (F(1.0) - sd) / pndm
        t1 = (F(-1.0) - sd) / pndm
        lo = np.where(t0 < t1, t0, t1)
        hi = np.where(t0 < t1, t1, t0)
        in_range = ~((lo > F(1.0)) | (hi < 0))
        lo = np.where(lo > 0, lo, F(0.0))
        p = c + m * lo[:, None]
        face = try_face & in_range & point_in_triangle(v0, v1, v2, p)
        # vertices (:44-81) and edges (:83-152)
        rm = RunningMin(n)
        for v in (v0, v1, v2):
            b = F(2.0) * dot(m, c - v)
            cc = length(v - c)
            cc = cc * cc - F(1.0)
            some, r0, r1 = get_roots(msl, b, cc)
            rm.push(r0, some)
            rm.push(r1, some)
        for a_, b_ in ((v0, v1), (v1, v2), (v2, v0)):
            edge = b_ - a_
            btv = a_ - c
            btv_len = length(btv)
            el = length(edge)
            els = el * el
            edm = dot(edge, m)
            edb = dot(edge, btv)
            qa = els * -msl + edm * edm
            qb = els * (F(2.0) * dot(m, btv)) - F(2.0) * edm * edb
            qc = els * (F(1.0) - btv_len * btv_len) + edb * edb
            some, r0, r1 = get_roots(qa, qb, qc)
            t = np.where(r1 < r0, r1, r0)                        # min(&[r0, r1]): seeded with r0, replaced if r1 < r0
            f = (edm * t - edb) / els
            rm.push(t, some & (f >= 0) & (f <= F(1.0)))
        ve = alive & ~face & rm.have & (rm.cur > 0)              # :201-202
    some = face | ve
    return some, np.where(face, lo, rm.cur).astype(np.float32)


def collision_fields(c, m, t):
    """Collision { time_to, distance_to, position, intersection } (:188-194, :203-209)."""
    with np.errstate(all="ignore"):
        pos = c + m * t[:, None]
        return t, length(m) * t, pos, pos + normalize(m) * F(1.0)


def closest(c, m, tris):
    """lib.rs:33-42 for S spheres against the same T triangles: filter_map + min_by, whose comparator never
    returns Equal, so among equal times the LATER triangle wins (and a NaN time always loses to what follows)."""
    s = len(c)
    have = np.zeros(s, bool)
    best_t = np.zeros(s, np.float32)
    best_j = np.full(s, 0xFFFFFFFF, np.uint32)
    for j, tri in enumerate(np.asarray(tris, np.float32)):
        some, t = detect(c, m, np.broadcast_to(tri, (s, 3, 3)))
        with np.errstate(invalid="ignore"):
            keep_old = have & (best_t < t)                       # Ordering::Less -> keep the accumulator
        take = some & ~keep_old
        best_t = np.where(take, t, best_t)
        best_j = np.where(take, np.uint32(j), best_j)
        have |= some
    return have, best_t, best_j


def respond(c, m, position, intersection):
    """sphere_triangle_calculate_response (collision_response.rs:6-35) -> (ok, new c, new m)."""
    with np.errstate(all="ignore"):
        n = normalize(position - intersection)
        dest = c + m
        d = dot(dest, n) + plane_constant(intersection, n)
        ok = d >= 0
        new_dest = dest - d[:, None] * n
        return ok, new_dest, new_dest - position


def collision_response(c, m, tris, max_iter):
    """lib.rs:29-62 for S spheres with the iteration cap of the batch API (status bits like the C ABI:
    2 = the reference's assert fired, 4 = cap reached).  Returns (c, m, iterations, status)."""
    c, m = np.array(c, np.float32), np.array(m, np.float32)
    s = len(c)
    it = np.zeros(s, np.uint32)
    status = np.zeros(s, np.uint32)
    active = np.ones(s, bool)
    rounds = 0
    while active.any():
        if max_iter and rounds >= max_iter:
            status[active] |= 4
            break
        idx = np.flatnonzero(active)
        have, t, _ = closest(c[idx], m[idx], tris)
        active[idx[~have]] = False                               # None -> break response
        h = idx[have]
        if len(h):
            _, _, pos, inter = collision_fields(c[h], m[h], t[have])
            ok, nc, nm = respond(c[h], m[h], pos, inter)
            status[h[~ok]] |= 2
            active[h[~ok]] = False
            g = h[ok]
            c[g], m[g] = nc[ok], nm[ok]
            it[g] += 1
            with np.errstate(invalid="ignore"):
                done = length(m[g]) < DISTANCE_EPSILON
            active[g[done]] = False
        rounds += 1
    return c, m, it, status
