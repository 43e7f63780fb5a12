#!/usr/bin/env python
"""CPU-only search (oracle) for pairs the reference's narrowphase hits but the broadphase predicate P rejects,
on thin triangles that the library does NOT list for exhaustive sweeping ("grazing band").
Spheres are aimed so that the face test's contact point lands a hair outside the triangle's long edge.

    python tools/search_grazing.py [--triangles N] [--seed S]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "xp-game-engine_b200")]
from oracle import xp_oracle as oracle  # noqa: E402


def listed(tri):
    t = tri.astype(np.float32)
    e1, e2 = t[1] - t[0], t[2] - t[0]
    c = np.cross(e1, e2)
    cc, l1, l2 = np.float32(c @ c), np.float32(e1 @ e1), np.float32(e2 @ e2)
    m2 = max(l1, l2)
    return (not cc > np.float32(1e-4) * (l1 * l2)) or (not cc * np.float32(1.7e5) > m2 * m2 * m2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--triangles", type=int, default=400)
    ap.add_argument("--spheres", type=int, default=4000)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    found = skipped = 0
    worst = 0.0
    for k in range(args.triangles):
        L = float(rng.choice([2.0, 6.4, 20.0, 64.0, 200.0]))
        ratio = float(rng.uniform(50, 400))                     # L^2 / w, below the listing threshold
        w = L * L / ratio
        if w > L:
            continue
        # right triangle in its own frame: v0 = 0, v1 = L x, v2 = L x + w y (thin), random rigid motion + offset
        q, _ = np.linalg.qr(rng.normal(0, 1, (3, 3)))
        mode = rng.integers(0, 3)
        if mode == 0:
            q = np.eye(3)
        elif mode == 1:                                         # almost axis aligned: the box is nearly tight along the normal
            a = rng.normal(0, 1, 3) * float(rng.choice([1e-6, 1e-4, 1e-2]))
            K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
            q = np.eye(3) + K + K @ K / 2
        org = rng.uniform(-30, 30, 3) * rng.choice([0.0, 1.0, 100.0])
        local = np.array([[0, 0, 0], [L, 0, 0], [L, w, 0]], np.float64)
        if rng.random() < 0.5:
            local = np.array([[0, 0, 0], [L, w, 0], [L, 0, 0]], np.float64)[[0, 2, 1]] * [1, 1, 1]
        tri = (local @ q.T + org).astype(np.float32)
        if listed(tri):
            skipped += 1
            continue
        t64 = tri.astype(np.float64)
        n = np.cross(t64[1] - t64[0], t64[2] - t64[0])
        n /= np.linalg.norm(n)
        ns = args.spheres
        # contact point: on the long edge v0-v1 (the one the AABB hugs when axis aligned), pushed outwards by delta
        s = rng.uniform(0.02, 0.98, ns)
        edge = t64[1] - t64[0]
        out_dir = np.cross(edge, n)
        out_dir /= np.linalg.norm(out_dir)
        if out_dir @ (t64[2] - t64[0]) > 0:
            out_dir = -out_dir
        delta = rng.uniform(-0.5 * w, 6e-4, ns) * rng.choice([1.0, 0.1], ns)     # inside the face or a hair outside it
        side = rng.choice([-1.0, 1.0], ns)
        p = t64[0] + s[:, None] * edge + delta[:, None] * out_dir + side[:, None] * n      # sphere centre at contact
        m = -side[:, None] * n * rng.uniform(0.2, 3.0, (ns, 1)) + rng.normal(0, 0.3, (ns, 3))
        t0 = rng.uniform(0.05, 2.0, ns)
        c = p - m * t0[:, None]
        resps = oracle.make_responses(c.astype(np.float32), m.astype(np.float32))
        want = oracle.detect_batch(resps, tri[None])
        pi, pj = oracle.broadphase_pairs(resps, tri[None])
        inP = np.zeros(ns, bool)
        inP[pi] = True
        bad = want[1].astype(bool) & ~inP
        if bad.any():
            found += int(bad.sum())
            i = int(np.flatnonzero(bad)[0])
            worst = max(worst, float(delta[bad].max()))
            print(f"tri {k}: L={L} w={w:.4g} L2/w={ratio:.0f}: {int(bad.sum())} hits outside P, e.g. delta={delta[i]:.2e} t={want[0]['time_to'][i]}")
    print(f"searched {args.triangles - skipped} unlisted thin triangles x {args.spheres} grazing spheres: "
          f"{found} pairs hit by the narrowphase but rejected by P (largest overshoot {worst:.2e})")


if __name__ == "__main__":
    main()
