"""Host-side mirror of the xp_physics public API on top of the C ABI (ctypes).

Names, argument meaning and error behaviour follow /root/reference/xp_physics/src/lib.rs:9-15:
    Sphere, Triangle, Response, Collision                       (sphere.rs, triangle.rs, response.rs, collision.rs)
    collision_response(response, triangles) -> Response         (lib.rs:29)
    collision_response_non_trianulated(response, vec3s)         (lib.rs:17)
    sphere_triangle_detect_collision(response, triangle)        (collision_detect.rs:155) -> Collision | None
    sphere_triangle_calculate_response(response, collision)     (collision_response.rs:6)
Where the reference panics these raise AssertionError (assert_eq!/assert!) or IndexError
(the chunks(3) out-of-bounds of lib.rs:19-24).  The batch forms live on `Scene`, which is the
"upload the triangle soup once, sweep many spheres" object the GPU needs.

All arithmetic happens on the GPU inside libxp_physics_cuda.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check

RESPONSE_DTYPE = np.dtype([("c", "<f4", (3,)), ("r", "<f4"), ("movement", "<f4", (3,))])
COLLISION_DTYPE = np.dtype([("time_to", "<f4"), ("distance_to", "<f4"),
                            ("position", "<f4", (3,)), ("intersection", "<f4", (3,))])
HIT_DTYPE = np.dtype([("time_to", "<f4"), ("tri_index", "<u4")])       # xp_hit: the compact result, 8 B
CONTACT_DTYPE = np.dtype([("time_to", "<f4"), ("distance_to", "<f4"), ("position", "<f4", (3,)),
                          ("intersection", "<f4", (3,)), ("tri_index", "<u4"), ("flags", "<u4")])
# xp_contact16 (contact_record="compact"): what a peer can use without the sphere's own Response
CONTACT16_DTYPE = np.dtype([("position", "<f4", (3,)), ("tri_index", "<u4")])
assert RESPONSE_DTYPE.itemsize == 28 and COLLISION_DTYPE.itemsize == 32 and CONTACT_DTYPE.itemsize == 40
assert CONTACT16_DTYPE.itemsize == 16


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- value types (reference structs) ------------------------------------------------------
@dataclass
class Sphere:                       # sphere.rs:4-13
    c: Sequence[float]
    r: float = 1.0


@dataclass
class Triangle:                     # triangle.rs:4-8,41-43
    v0: Sequence[float]
    v1: Sequence[float]
    v2: Sequence[float]

    def as_array(self) -> np.ndarray:
        return np.asarray([self.v0, self.v1, self.v2], np.float32)


@dataclass
class Response:                     # response.rs:5-8
    sphere: Sphere
    movement: Sequence[float]

    def as_record(self) -> np.ndarray:
        return make_responses([self.sphere.c], [self.movement], self.sphere.r)

    @staticmethod
    def from_record(rec) -> "Response":
        return Response(Sphere([float(x) for x in rec["c"]], float(rec["r"])),
                        [float(x) for x in rec["movement"]])


@dataclass
class Collision:                    # collision.rs:5-10
    time_to: float
    distance_to: float
    position: Sequence[float]
    intersection: Sequence[float]

    def as_record(self) -> np.ndarray:
        out = np.zeros(1, COLLISION_DTYPE)
        out["time_to"], out["distance_to"] = self.time_to, self.distance_to
        out["position"], out["intersection"] = self.position, self.intersection
        return out

    @staticmethod
    def from_record(rec) -> "Collision":
        return Collision(float(rec["time_to"]), float(rec["distance_to"]),
                         [float(x) for x in rec["position"]], [float(x) for x in rec["intersection"]])


def make_responses(c, movement, r=1.0) -> np.ndarray:
    c = np.asarray(c, np.float32).reshape(-1, 3)
    m = np.asarray(movement, np.float32).reshape(-1, 3)
    out = np.zeros(len(c), RESPONSE_DTYPE)
    out["c"], out["movement"], out["r"] = c, m, r
    return out


def as_triangles(t) -> np.ndarray:
    if isinstance(t, (list, tuple)) and len(t) and isinstance(t[0], Triangle):
        t = np.stack([x.as_array() for x in t])
    return np.ascontiguousarray(np.asarray(t, np.float32).reshape(-1, 3, 3))


# ---- the scene: triangle soup uploaded once + device LBVH ------------------------------------
class Scene:
    """Owns an xp_scene (device triangle records + LBVH) on one GPU."""

    def __init__(self, device: int = 0, triangles=None, *, arith: str = "exact", method: str = "lbvh_pairs",
                 prune="auto", sort_spheres: bool = True, timing: bool = False, far_exact: bool = False):
        self._L = _lib.lib()
        h = C.c_void_p()
        check(self._L.xp_scene_create(device, C.byref(h)), "xp_scene_create")
        self._h = h
        self.device = device
        self.set_options(arith=arith, method=method, prune=prune, sort_spheres=sort_spheres, timing=timing,
                         far_exact=far_exact)
        if triangles is not None:
            self.set_triangles(triangles)
            self.build()

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._L.xp_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- options
    def set_options(self, *, arith=None, method=None, prune=None, sort_spheres=None, timing=None, max_pairs=None,
                    far_exact=None, grid_walk=None, contact_record=None, leaf_runs=None):
        L, h = self._L, self._h
        if contact_record is not None:           # "full" (xp_contact, 40 B) or "compact" (xp_contact16, 16 B)
            check(L.xp_scene_set_option(h, _lib.OPT_CONTACT_RECORD, {"full": 0, "compact": 1}[contact_record]))
        if arith is not None:
            check(L.xp_scene_set_option(h, _lib.OPT_ARITH, _lib.ARITH[arith]))
        if method is not None:
            check(L.xp_scene_set_option(h, _lib.OPT_METHOD, _lib.METHODS[method]))
        if prune is not None:
            check(L.xp_scene_set_option(h, _lib.OPT_PRUNE, {False: 0, True: 1, "auto": 2}.get(prune, prune)))
        if sort_spheres is not None:
            check(L.xp_scene_set_option(h, _lib.OPT_SORT_SPHERES, int(bool(sort_spheres))))
        if timing is not None:
            check(L.xp_scene_set_option(h, _lib.OPT_TIMING, int(bool(timing))))
        if max_pairs is not None:
            check(L.xp_scene_set_option(h, _lib.OPT_MAX_PAIRS, int(max_pairs)))
        if far_exact is not None:
            check(L.xp_scene_set_option(h, _lib.OPT_FAR_EXACT, int(bool(far_exact))))
        if grid_walk is not None:
            check(L.xp_scene_set_option(h, _lib.OPT_GRID_WALK, int(bool(grid_walk))))
        if leaf_runs is not None:
            check(L.xp_scene_set_option(h, _lib.OPT_LEAF_RUNS, int(bool(leaf_runs))))

    def set_stream(self, cuda_stream_ptr: int):
        check(self._L.xp_scene_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    # -- geometry
    def set_triangles(self, triangles):
        t = as_triangles(triangles)
        self._keep = t
        check(self._L.xp_scene_set_triangles(self._h, _p(t), len(t)), "xp_scene_set_triangles")

    def set_triangles_dev(self, dev_ptr: int, n: int):
        check(self._L.xp_scene_set_triangles_dev(self._h, C.c_void_p(dev_ptr), n), "xp_scene_set_triangles_dev")

    def set_positions(self, vec3s):
        v = np.ascontiguousarray(np.asarray(vec3s, np.float32).reshape(-1, 3))
        rc = self._L.xp_scene_set_positions(self._h, _p(v), len(v))
        if rc == _lib.XP_EINVAL:
            raise IndexError("index out of bounds: chunks(3) of a slice whose length is not a multiple of 3 "
                             "(xp_physics/src/lib.rs:19-24)")
        check(rc, "xp_scene_set_positions")

    def set_heightmap(self, heights, x0: float = 0.0, z0: float = 0.0, spacing: float = 1.0):
        """heights[(nx+1),(nz+1)] -> 2*nx*nz triangles emitted on the device (same soup and order as
        scenes.heightmap_from_heights)."""
        h = np.ascontiguousarray(heights, np.float32)
        check(self._L.xp_scene_set_heightmap(self._h, _p(h), h.shape[0] - 1, h.shape[1] - 1, x0, z0, spacing),
              "xp_scene_set_heightmap")

    def set_clipmap(self, clipmap):
        """The collision soup of a `clipmap.Clipmap` (toroidal height store + part layout), emitted on the
        device; same triangles, same order as `clipmap.triangles()`."""
        from .clipmap import CM_UNIT_SIZE_SMALLEST
        h = np.ascontiguousarray(clipmap.data, np.float32)
        parts = np.ascontiguousarray(clipmap.part_table(), np.int32)
        check(self._L.xp_scene_set_clipmap(self._h, _p(h), clipmap.levels, clipmap.size, _p(parts), len(parts),
                                           float(CM_UNIT_SIZE_SMALLEST)), "xp_scene_set_clipmap")

    def build(self):
        check(self._L.xp_scene_build(self._h), "xp_scene_build")

    @property
    def triangle_count(self) -> int:
        return int(self._L.xp_scene_triangle_count(self._h))

    def stats(self) -> dict:
        st = _lib.XpStats()
        check(self._L.xp_scene_get_stats(self._h, C.byref(st)))
        return {"narrowphase_tests": int(st.narrowphase_tests), "broadphase_pairs": int(st.broadphase_pairs),
                "node_visits": int(st.node_visits), "kernel_launches": int(st.kernel_launches),
                "iterations_run": int(st.iterations_run),
                "stage_ms": {n: float(st.stage_ms[i]) for i, n in enumerate(_lib.STAGE_NAMES)},
                "path_tests": dict(zip(("cull", "far", "face_hit", "full", "full_pit"),
                                       (int(x) for x in st.path_tests)))}

    # -- sweeps, host buffers
    def detect_batch(self, responses: np.ndarray, horizon: str = "inf"):
        """Closest collision per sphere (a11 + the min_by of lib.rs:33-42).
        Returns (collisions, hit, tri_index, status)."""
        r = np.ascontiguousarray(responses, RESPONSE_DTYPE)
        n = len(r)
        col = np.zeros(n, COLLISION_DTYPE)
        hit = np.zeros(n, np.uint8)
        tri = np.full(n, 0xFFFFFFFF, np.uint32)
        st = np.zeros(n, np.uint32)
        check(self._L.xp_detect_batch(self._h, _p(r), n, _horizon(horizon), _p(col), _p(hit), _p(tri), _p(st)),
              "xp_detect_batch")
        return col, hit, tri, st

    def detect_batch_submit(self, h_in: int, n: int, h_col: int = 0, h_hit: int = 0, h_tri: int = 0,
                            h_status: int = 0, horizon: str = "inf") -> int:
        """Queue one batch of a stream of batches (xp_detect_batch_submit): host addresses (ints) of
        page-locked buffers that stay untouched until `detect_batch_wait(ticket)`.  Up to
        XP_ASYNC_SLOTS = 2 batches may be in flight; one batch's copies overlap another's kernels."""
        vp = C.c_void_p
        ticket = C.c_int(-1)
        check(self._L.xp_detect_batch_submit(self._h, vp(h_in), n, _horizon(horizon), vp(h_col or None),
                                             vp(h_hit or None), vp(h_tri or None), vp(h_status or None),
                                             C.byref(ticket)), "xp_detect_batch_submit")
        return ticket.value

    def detect_batch_compact(self, responses: np.ndarray, horizon: str = "inf"):
        """Closest collision per sphere in the compact form (xp_hit: time_to, tri_index; 0xFFFFFFFF = none).
        Returns (hits, status); `collision_from_hit` rebuilds the reference's Collision."""
        r = np.ascontiguousarray(responses, RESPONSE_DTYPE)
        out = np.zeros(len(r), HIT_DTYPE)
        st = np.zeros(len(r), np.uint32)
        check(self._L.xp_detect_batch_compact(self._h, _p(r), len(r), _horizon(horizon), _p(out), _p(st)),
              "xp_detect_batch_compact")
        return out, st

    def detect_batch_compact_dev(self, d_in: int, n: int, d_out: int, d_status: int = 0, horizon: str = "inf"):
        vp = C.c_void_p
        check(self._L.xp_detect_batch_compact_dev(self._h, vp(d_in), n, _horizon(horizon), vp(d_out), vp(d_status or None)),
              "xp_detect_batch_compact_dev")

    def detect_batch_compact_submit(self, h_in: int, n: int, h_out: int, h_status: int = 0, horizon: str = "inf") -> int:
        """Queue one batch with compact results (8 B per sphere back over the host link); same tickets and
        `detect_batch_wait` as `detect_batch_submit`."""
        vp = C.c_void_p
        ticket = C.c_int(-1)
        check(self._L.xp_detect_batch_compact_submit(self._h, vp(h_in), n, _horizon(horizon), vp(h_out), vp(h_status or None),
                                                     C.byref(ticket)), "xp_detect_batch_compact_submit")
        return ticket.value

    def detect_batch_wait(self, ticket: int) -> None:
        """Block until the batch behind `ticket` is in the caller's buffers; `stats()` then describes it."""
        check(self._L.xp_detect_batch_wait(self._h, ticket), "xp_detect_batch_wait")

    def collision_response_batch(self, responses: np.ndarray, max_iter: int = 0, horizon: str = "inf"):
        """collision_response (lib.rs:29-62) for every sphere.
        Returns (responses_out, iterations, status, last_slide_normal)."""
        r = np.ascontiguousarray(responses, RESPONSE_DTYPE)
        n = len(r)
        out = np.zeros(n, RESPONSE_DTYPE)
        it = np.zeros(n, np.uint32)
        st = np.zeros(n, np.uint32)
        ln = np.zeros((n, 3), np.float32)
        check(self._L.xp_collision_response_batch(self._h, _p(r), n, max_iter, _horizon(horizon), _p(out),
                                                  _p(it), _p(st), _p(ln)), "xp_collision_response_batch")
        return out, it, st, ln

    def collide_and_slide_batch(self, position, velocity, max_iter: int = 5, radii=None):
        """Row f4 on the device (xp_collide_and_slide_batch): Fauerby's collide-and-slide with the true contact point,
        unit horizon and optional ellipsoid radii (the scene must be in that ellipsoid's space).
        Returns (position, remaining velocity, contacts handled), like fauerby.collide_and_slide."""
        p = np.ascontiguousarray(np.asarray(position, np.float32).reshape(-1, 3))
        v = np.ascontiguousarray(np.asarray(velocity, np.float32).reshape(-1, 3))
        assert p.shape == v.shape
        op, ov = np.zeros_like(p), np.zeros_like(v)
        handled = np.zeros(len(p), np.uint32)
        rad = None if radii is None else np.ascontiguousarray(np.broadcast_to(np.asarray(radii, np.float32).reshape(-1), (3,)))
        check(self._L.xp_collide_and_slide_batch(self._h, _p(p), _p(v), len(p), max_iter, _p(rad), _p(op), _p(ov), _p(handled)),
              "xp_collide_and_slide_batch")
        return op, ov, handled.astype(np.int32)

    def broadphase_pairs(self, responses: np.ndarray, horizon: str = "inf", cap: Optional[int] = None):
        r = np.ascontiguousarray(responses, RESPONSE_DTYPE)
        cap = int(cap if cap is not None else max(1024, min(len(r) * max(self.triangle_count, 1), 1 << 26)))
        si = np.zeros(cap, np.uint32)
        ti = np.zeros(cap, np.uint32)
        cnt = C.c_uint64()
        rc = self._L.xp_broadphase_pairs(self._h, _p(r), len(r), _horizon(horizon), cap, _p(si), _p(ti), C.byref(cnt))
        if rc == _lib.XP_ECAP:
            raise MemoryError(f"broadphase pair capacity {cap} < {cnt.value}")
        check(rc, "xp_broadphase_pairs")
        return si[:cnt.value].copy(), ti[:cnt.value].copy()

    def debug_tree(self):
        n = self.triangle_count
        ni = max(n - 1, 1)
        sorted_tri = np.zeros(n, np.uint32)
        morton = np.zeros(n, np.uint32)
        child = np.zeros((ni, 2), np.int32)
        lo = np.zeros((ni, 3), np.float32)
        hi = np.zeros((ni, 3), np.float32)
        check(self._L.xp_scene_debug_tree(self._h, _p(sorted_tri), _p(morton), _p(child), _p(lo), _p(hi)))
        return sorted_tri, morton, child, lo, hi

    def debug_leaf_runs(self):
        run = np.zeros((max(self.triangle_count - 1, 1), 2), np.int32)
        check(self._L.xp_scene_debug_leaf_runs(self._h, _p(run)))
        return run

    # -- sweeps, device pointers (ints, e.g. torch.Tensor.data_ptr())
    def detect_batch_dev(self, d_in: int, n: int, d_col: int = 0, d_hit: int = 0, d_tri: int = 0,
                         d_status: int = 0, horizon: str = "inf"):
        vp = C.c_void_p
        check(self._L.xp_detect_batch_dev(self._h, vp(d_in), n, _horizon(horizon), vp(d_col or None), vp(d_hit or None),
                                          vp(d_tri or None), vp(d_status or None)), "xp_detect_batch_dev")

    def detect_contacts_dev(self, d_in: int, n: int, d_contacts: int, horizon: str = "inf"):
        vp = C.c_void_p
        check(self._L.xp_detect_contacts_dev(self._h, vp(d_in), n, _horizon(horizon), vp(d_contacts)),
              "xp_detect_contacts_dev")

    # -- fused contact exchange over peer memory (multi-GPU)
    def exchange_alloc(self, nbytes: int):
        """Allocate a gather buffer on this scene's GPU.  Returns (device pointer, 64-byte IPC handle)."""
        ptr = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        check(self._L.xp_exchange_alloc(self._h, nbytes, C.byref(ptr), handle), "xp_exchange_alloc")
        return int(ptr.value), bytes(handle)

    def exchange_free(self, ptr: int):
        check(self._L.xp_exchange_free(self._h, C.c_void_p(ptr)), "xp_exchange_free")

    def exchange_open(self, handle: bytes) -> int:
        """Map a peer rank's gather buffer (IPC handle from its exchange_alloc) into this process."""
        ptr = C.c_void_p()
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        check(self._L.xp_exchange_open(self._h, buf, C.byref(ptr)), "xp_exchange_open")
        return int(ptr.value)

    def exchange_close(self, ptr: int):
        check(self._L.xp_exchange_close(self._h, C.c_void_p(ptr)), "xp_exchange_close")

    def detect_contacts_exchange_dev(self, d_in: int, n: int, bufs, slot_first: int, horizon: str = "inf"):
        """Sweep and write every contact record into each buffer of `bufs` (own + mapped peers) at record
        index slot_first + i: the record-building kernel performs the all-gather with P2P stores."""
        arr = (C.c_void_p * len(bufs))(*[C.c_void_p(b) for b in bufs])
        check(self._L.xp_detect_contacts_exchange_dev(self._h, C.c_void_p(d_in), n, _horizon(horizon), arr, len(bufs),
                                                      slot_first), "xp_detect_contacts_exchange_dev")

    def set_exchange_stream(self, cuda_stream: int):
        """Second stream for `detect_contacts_exchange_submit` (record build + exchange of a step)."""
        check(self._L.xp_scene_set_exchange_stream(self._h, C.c_void_p(cuda_stream)), "xp_scene_set_exchange_stream")

    def detect_contacts_exchange_submit(self, d_in: int, n: int, bufs, slot_first: int, horizon: str = "inf",
                                        self_index: int = -1) -> int:
        """Queue one step of a loop (sweep on the scene's stream, exchange on the exchange stream) and
        return its ticket; `detect_contacts_exchange_wait(ticket)` completes it.  Two steps may be in flight."""
        arr = (C.c_void_p * len(bufs))(*[C.c_void_p(b) for b in bufs])
        ticket = C.c_int(-1)
        check(self._L.xp_detect_contacts_exchange_submit(self._h, C.c_void_p(d_in), n, _horizon(horizon), arr, len(bufs),
                                                         self_index, slot_first, C.byref(ticket)),
              "xp_detect_contacts_exchange_submit")
        return ticket.value

    def detect_contacts_exchange_wait(self, ticket: int) -> bool:
        """Block until the step's records have left this GPU.  False = the candidate buffer overflowed and
        the step must be redone with the synchronous call (rare)."""
        rc = self._L.xp_detect_contacts_exchange_wait(self._h, ticket)
        if rc == _lib.XP_ECAP:
            return False
        check(rc, "xp_detect_contacts_exchange_wait")
        return True

    def detect_contacts_multicast_dev(self, d_in: int, n: int, multicast_buf: int, slot_first: int, horizon: str = "inf"):
        """Sweep and write every contact record ONCE, to an NVSwitch multicast address bound to all ranks'
        gather buffers (record index slot_first + i): the switch performs the replication."""
        check(self._L.xp_detect_contacts_multicast_dev(self._h, C.c_void_p(d_in), n, _horizon(horizon),
                                                       C.c_void_p(multicast_buf), slot_first),
              "xp_detect_contacts_multicast_dev")

    def collision_response_batch_dev(self, d_in: int, n: int, d_out: int, max_iter: int = 0, d_iter: int = 0,
                                     d_status: int = 0, d_normal: int = 0, horizon: str = "inf"):
        vp = C.c_void_p
        check(self._L.xp_collision_response_batch_dev(self._h, vp(d_in), n, max_iter, _horizon(horizon), vp(d_out),
                                                      vp(d_iter or None), vp(d_status or None), vp(d_normal or None)),
              "xp_collision_response_batch_dev")


def collision_from_hit(responses: np.ndarray, hits: np.ndarray):
    """The reference's Collision from the compact result (collision_detect.rs:188-194 / :203-209): position = c + m*t,
    distance_to = |m|*t, intersection = position + m/|m| * r -- float32, unfused, in the reference's operation
    order, so the fields equal the ones xp_detect_batch returns bit for bit.  Returns (collisions, hit)."""
    r = np.ascontiguousarray(responses, RESPONSE_DTYPE)
    h = np.ascontiguousarray(hits, HIT_DTYPE)
    hit = h["tri_index"] != 0xFFFFFFFF
    col = np.zeros(len(r), COLLISION_DTYPE)
    c, m, t = r["c"][hit], r["movement"][hit], h["time_to"][hit][:, None]
    with np.errstate(all="ignore"):
        ml = np.sqrt(((m[:, 0] * m[:, 0] + m[:, 1] * m[:, 1]) + m[:, 2] * m[:, 2]).astype(np.float32))[:, None]   # nalgebra dot order
        pos = (c + (m * t).astype(np.float32)).astype(np.float32)
        col["time_to"][hit] = t[:, 0]
        col["distance_to"][hit] = (ml * t)[:, 0]
        col["position"][hit] = pos
        col["intersection"][hit] = pos + ((m / ml).astype(np.float32) * r["r"][hit][:, None]).astype(np.float32)
    return col, hit.astype(np.uint8)


def _horizon(h) -> int:
    return {"inf": _lib.HORIZON_INF, "unit": _lib.HORIZON_UNIT, 0: 0, 1: 1}[h]


# ---- the reference's free functions ----------------------------------------------------------
def sphere_triangle_detect_collision(response: Response, triangle: Triangle, device: int = 0) -> Optional[Collision]:
    """collision_detect.rs:155-213 (one sphere, one triangle), evaluated on the GPU."""
    L = _lib.lib()
    r = response.as_record()
    t = np.ascontiguousarray(triangle.as_array())
    out = np.zeros(1, COLLISION_DTYPE)
    rc = L.xp_sphere_triangle_detect_collision(device, _p(r), _p(t), _p(out))
    if rc == _lib.XP_EINVAL:
        raise AssertionError("assertion failed: `(left == right)` sphere.r == 1.0 (collision_detect.rs:161)")
    check(rc, "xp_sphere_triangle_detect_collision")
    return Collision.from_record(out[0]) if rc == 1 else None


def sphere_triangle_calculate_response(response: Response, collision: Collision, device: int = 0) -> Response:
    """collision_response.rs:6-35, evaluated on the GPU."""
    L = _lib.lib()
    r = response.as_record()
    c = collision.as_record()
    out = np.zeros(1, RESPONSE_DTYPE)
    rc = L.xp_sphere_triangle_calculate_response(device, _p(r), _p(c), _p(out))
    check(rc, "xp_sphere_triangle_calculate_response")
    if rc == 1:
        raise AssertionError("assertion failed: original_destination_to_plane_distance >= 0.0 "
                             "(collision_response.rs:22)")
    return Response.from_record(out[0])


def collision_response(response: Response, triangles, device: int = 0, *, max_iter: int = 0, **scene_opts) -> Response:
    """lib.rs:29-62 for one sphere.  Uploads the triangles, builds the LBVH, sweeps."""
    with Scene(device, triangles, **scene_opts) as s:
        out, _, st, _ = s.collision_response_batch(response.as_record(), max_iter=max_iter)
    _raise_status(int(st[0]))
    return Response.from_record(out[0])


def collision_response_non_trianulated(response: Response, vec3s, device: int = 0, *, max_iter: int = 0,
                                       **scene_opts) -> Response:
    """lib.rs:17-27: every three Vec3 are one triangle."""
    with Scene(device, **scene_opts) as s:
        s.set_positions(vec3s)
        s.build()
        out, _, st, _ = s.collision_response_batch(response.as_record(), max_iter=max_iter)
    _raise_status(int(st[0]))
    return Response.from_record(out[0])


def _raise_status(st: int):
    if st & _lib.ST_RADIUS_NE_1:
        raise AssertionError("assertion failed: `(left == right)` sphere.r == 1.0 (collision_detect.rs:161)")
    if st & _lib.ST_ASSERT_NEG_DISTANCE:
        raise AssertionError("assertion failed: original_destination_to_plane_distance >= 0.0 "
                             "(collision_response.rs:22)")


def probe_fp32_peak(device: int = 0) -> float:
    v = C.c_double()
    check(_lib.lib().xp_probe_fp32_peak(device, C.byref(v)), "xp_probe_fp32_peak")
    return float(v.value)


def debug_check_arith(device: int = 0, n_random: int = 1 << 32, seed: int = 1) -> dict:
    """Device self-check of the exact mode's hand-rolled division / square root and operand-side decisions
    (xp_debug_check_arith): mismatch counters per check, all zero when they are exact."""
    c = (C.c_uint64 * 8)()
    check(_lib.lib().xp_debug_check_arith(device, n_random, seed, c), "xp_debug_check_arith")
    names = ("sqrt", "div_random", "div_mantissa", "edge_param", "face_interval", "vertex_min")
    return {k: int(c[i]) for i, k in enumerate(names)}


def device_count() -> int:
    n = C.c_int()
    rc = _lib.lib().xp_device_count(C.byref(n))
    return int(n.value) if rc == 0 else 0
