#!/usr/bin/env python
"""Stress parity on a GPU box: mid-size scenes (where the sphere ordering, the persistent broadphase with work
stealing and the span-wise narrowphase all engage) against the CPU oracle, bit for bit.  Complements the unit
tests, whose scenes are sized for seconds.   python tools/stress_parity.py [--cases N] [--seed S]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "xp-game-engine_b200")]
import xp_physics_cuda as xp  # noqa: E402
from oracle import xp_oracle as oracle  # noqa: E402
from xp_physics_cuda import scenes  # noqa: E402


def bits(a):
    a = np.ascontiguousarray(a, np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


def fields(col):
    return np.concatenate([col["time_to"][:, None], col["distance_to"][:, None], col["position"], col["intersection"]], axis=1)


def rfields(r):
    return np.concatenate([r["c"], r["r"][:, None], r["movement"]], axis=1)


def make_case(rng, k):
    kind = k % 6
    if kind == 0:                                   # terrain
        nx = int(rng.integers(30, 120))
        tris, h = scenes.sine_terrain(nx, nx + 3, seed=int(rng.integers(1, 1000)))
        resps = scenes.terrain_spheres(h, int(rng.integers(3000, 40000)), seed=int(rng.integers(1, 1000)))
    elif kind == 1:                                 # mesh field (3D, dense)
        tris, lo, hi = scenes.mesh_field(int(rng.integers(2000, 12000)), seed=int(rng.integers(1, 1000)))
        resps = scenes.field_spheres(lo, hi, int(rng.integers(2000, 9000)), seed=int(rng.integers(1, 1000)))
    elif kind == 2:                                 # random soup with degenerate triangles
        nt = int(rng.integers(500, 6000))
        ext = float(rng.choice([5.0, 30.0, 200.0]))
        base = rng.uniform(-ext, ext, (nt, 1, 3))
        tris = (base + rng.normal(0, rng.choice([0.1, 1.0, 5.0]), (nt, 3, 3))).astype(np.float32)
        tris[1] = tris[0]
        tris[2, 2] = tris[2, 1]
        tris[3, 1] = (tris[3, 0] + tris[3, 2]) / 2
        ns = int(rng.integers(1500, 20000))
        c = rng.uniform(-ext - 3, ext + 3, (ns, 3)).astype(np.float32)
        m = rng.normal(0, 1, (ns, 3))
        m = (m / np.linalg.norm(m, axis=1, keepdims=True) * rng.uniform(0.01, 2.5, (ns, 1))).astype(np.float32)
        m[::17, int(rng.integers(0, 3))] = 0.0
        m[5] = 0.0
        resps = xp.make_responses(c, m)
        resps["r"][11] = 0.5                        # the reference panics here: status bit
    elif kind == 3:                                 # clipmap + voxels
        tris = scenes.config4_scene(center_xz=(float(rng.uniform(-50, 50)), float(rng.uniform(-50, 50))))
        resps = scenes.scene_spheres(tris, int(rng.integers(1500, 4000)), seed=int(rng.integers(1, 1000)))
    elif kind == 4:                                 # adversarial shapes: needles, thin strips, huge and tiny triangles
        nt = int(rng.integers(800, 4000))
        base = rng.uniform(-20, 20, (nt, 1, 3))
        a = rng.normal(0, 1, (nt, 3)); b = rng.normal(0, 1, (nt, 3))
        la = rng.choice([1e-3, 0.05, 1.0, 30.0, 300.0], nt); lb = rng.choice([1e-3, 0.05, 1.0, 30.0], nt)
        a = a / np.linalg.norm(a, axis=1, keepdims=True) * la[:, None]
        b = b / np.linalg.norm(b, axis=1, keepdims=True) * lb[:, None]
        thin = rng.random(nt) < 0.3
        b[thin] = a[thin] * rng.uniform(0.2, 3.0, (int(thin.sum()), 1)) + rng.normal(0, 1, (int(thin.sum()), 3)) * rng.choice([1e-7, 1e-5, 1e-3, 1e-2], (int(thin.sum()), 1))
        tris = (base + np.stack([np.zeros_like(a), a, b], axis=1)).astype(np.float32)
        ns = int(rng.integers(2000, 9000))
        c = rng.uniform(-25, 25, (ns, 3)).astype(np.float32)
        m = rng.normal(0, 1, (ns, 3)) * rng.choice([1e-6, 1e-2, 1.0, 100.0], (ns, 1))
        resps = xp.make_responses(c, m.astype(np.float32))
    else:                                           # terrain far from the origin (coordinates ~ 1e4): few mantissa bits left
        nx = int(rng.integers(30, 80))
        tris, h = scenes.sine_terrain(nx, nx, seed=int(rng.integers(1, 1000)))
        resps = scenes.terrain_spheres(h, int(rng.integers(3000, 12000)), seed=int(rng.integers(1, 1000)))
        off = np.array([rng.choice([-1, 1]) * 8192.0, rng.choice([0.0, 4096.0]), rng.choice([-1, 1]) * 16384.0], np.float32)
        tris = tris + off
        resps["c"] += off
    return np.ascontiguousarray(tris, np.float32), resps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=12)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--max-iter", type=int, default=3, help="response iterations (0 = unbounded like the reference)")
    ap.add_argument("--far-exact", type=int, default=0, help="1: run with XP_OPT_FAR_EXACT")
    ap.add_argument("--kinds", default="", help="comma list of case kinds 0..5 to cycle through (default: all)")
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    kinds = [int(x) for x in args.kinds.split(",")] if args.kinds else list(range(6))
    bad = 0
    with xp.Scene(0) as s:
        for k in range(args.cases):
            tris, resps = make_case(rng, kinds[k % len(kinds)])
            t0 = time.time()
            want = oracle.detect_batch(resps, tris)
            wres = oracle.collision_response_batch(resps, tris, max_iter=args.max_iter)
            t_cpu = time.time() - t0
            s.set_options(method="lbvh_pairs", prune=bool(k % 3 == 2), far_exact=bool(args.far_exact))
            s.set_triangles(tris)
            s.build()
            t0 = time.time()
            got = s.detect_batch(resps)
            out, it, st, ln = s.collision_response_batch(resps, max_iter=args.max_iter)
            t_gpu = time.time() - t0
            h = want[1].astype(bool)
            ok = (np.array_equal(got[1], want[1]) and np.array_equal(bits(fields(got[0])[h]), bits(fields(want[0])[h]))
                  and np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
                  and np.array_equal(it, wres[1]) and np.array_equal(st, wres[2])
                  and np.array_equal(bits(rfields(out)), bits(rfields(wres[0]))))
            bad += not ok
            print(f"case {k:2d} kind {kinds[k % len(kinds)]} tris {len(tris):7d} spheres {len(resps):6d} hits {int(h.sum()):6d} "
                  f"prune {int(k % 3 == 2)} cpu {t_cpu:5.1f}s gpu {t_gpu:5.2f}s  {'identical' if ok else 'MISMATCH'}", flush=True)
    print("stress parity:", "ok" if bad == 0 else f"{bad} MISMATCHING CASES")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
