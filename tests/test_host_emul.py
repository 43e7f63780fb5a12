"""CPU check that the RE-ORGANISED device formulas (xp_narrowphase.cuh, compiled for the host by
tests/host_emul) are bit-identical to the oracle's line-by-line restatement of the reference.
This does not replace the GPU parity tests; it finds formula bugs before GPU time is spent."""
import ctypes as C

import numpy as np
import pytest

from tests.conftest import assert_bits_equal, bits
from tests.host_emul import lib as emul_lib
from xp_physics_cuda import scenes


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _emul_detect(E, resp, tri, mode=0):
    out = np.zeros(8, np.float32)
    r = E.emul_detect(_p(resp), _p(tri), _p(out), mode)
    return r, out


def _cases(seed, n):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        kind = rng.integers(0, 11)
        tri = rng.uniform(-3, 3, (3, 3)).astype(np.float32)
        c = rng.uniform(-4, 4, 3).astype(np.float32)
        m = rng.uniform(-1, 1, 3).astype(np.float32)
        if kind == 1:      # sphere above a roughly horizontal triangle, moving down
            tri[:, 1] = rng.uniform(-0.2, 0.2, 3)
            c[1] = rng.uniform(1.0, 3.0)
            m[1] = -abs(m[1]) - 0.05
        elif kind == 2:    # axis-aligned movement (zeros in m)
            m[rng.integers(0, 3)] = 0.0
        elif kind == 3:    # integer lattice: lots of exact ties and zeros
            tri = rng.integers(-3, 4, (3, 3)).astype(np.float32)
            c = rng.integers(-4, 5, 3).astype(np.float32)
            m = rng.integers(-2, 3, 3).astype(np.float32)
        elif kind == 4:    # degenerate triangle (quirk B10)
            tri[2] = tri[1]
        elif kind == 5:    # already touching / embedded (|sd| < 1)
            c = (tri.mean(axis=0) + rng.uniform(-0.8, 0.8, 3)).astype(np.float32)
        elif kind == 6:    # zero movement: a == 0, roots are +-inf / NaN (quirk B11)
            m[:] = 0.0
            c = (tri.mean(axis=0) + rng.uniform(-1.5, 1.5, 3)).astype(np.float32)
        elif kind == 7:    # astronomically large values: b, det overflow to inf
            scale = np.float32(10.0 ** rng.integers(15, 25))
            tri = (tri * scale).astype(np.float32)
            c = (c * scale).astype(np.float32)
            m = (m * np.float32(10.0 ** rng.integers(0, 20))).astype(np.float32)
        elif kind == 8:    # NaN / inf in the inputs
            which = rng.integers(0, 3)
            bad = np.float32([np.nan, np.inf, -np.inf][rng.integers(0, 3)])
            if which == 0:
                tri[rng.integers(0, 3), rng.integers(0, 3)] = bad
            elif which == 1:
                c[rng.integers(0, 3)] = bad
            else:
                m[rng.integers(0, 3)] = bad
        elif kind == 9:    # vanishing movement: |m|^2 underflows (denormal / zero denominators)
            m = (m * np.float32(10.0 ** -rng.integers(18, 30))).astype(np.float32)
            c = (tri.mean(axis=0) + rng.uniform(-1.5, 1.5, 3)).astype(np.float32)
        elif kind == 10:   # movement along an edge: a of the edge quadratic rounds to ~0 of either sign
            k = rng.integers(0, 3)
            m = ((tri[(k + 1) % 3] - tri[k]) * np.float32(rng.uniform(0.05, 0.5))).astype(np.float32)
            c = (tri[k] + rng.uniform(-1.2, 1.2, 3)).astype(np.float32)
        yield c, m, np.ascontiguousarray(tri)


def test_detect_lean_equals_oracle_bitwise(oracle):
    E = emul_lib()
    n_hit = 0
    for c, m, tri in _cases(11, 30000):
        resp = oracle.make_responses([c], [m])
        want = oracle.detect(resp, tri)
        r, got = _emul_detect(E, resp, tri)
        assert (want is not None) == (r == 1), (c, m, tri)
        if want is not None:
            n_hit += 1
            w = np.concatenate([[want["time_to"], want["distance_to"]], want["position"], want["intersection"]])
            assert_bits_equal(got, w.astype(np.float32), "collision fields")
    assert n_hit > 1000


def test_detect_with_stored_edge_lengths_equals_oracle_bitwise(oracle):
    """k_narrow_pairs reads |v1-v0|^2 and |v2-v1|^2 from the triangle record (k_pack computes them with
    xp_edge_len2); the result must not change by a bit."""
    E = emul_lib()
    n_hit = 0
    for c, m, tri in _cases(13, 20000):
        resp = oracle.make_responses([c], [m])
        want = oracle.detect(resp, tri)
        r, got = _emul_detect(E, resp, tri, mode=1)
        r0, got0 = _emul_detect(E, resp, tri, mode=0)
        assert r == r0 and np.array_equal(got.view(np.uint32), got0.view(np.uint32))
        assert (want is not None) == (r == 1), (c, m, tri)
        if want is not None:
            n_hit += 1
            w = np.concatenate([[want["time_to"], want["distance_to"]], want["position"], want["intersection"]])
            assert_bits_equal(got, w.astype(np.float32), "collision fields")
    assert n_hit > 700


def test_detect_lean_equals_oracle_on_terrain(oracle):
    E = emul_lib()
    tris, h = scenes.sine_terrain(24, 24, seed=3)
    resps = scenes.terrain_spheres(h, 64, seed=5)
    out_t = np.zeros(len(tris), np.float32)
    out_hit = np.zeros(len(tris), np.int32)
    flat = np.ascontiguousarray(tris.reshape(-1, 9))
    total_hits = 0
    for i in range(len(resps)):
        E.emul_detect_many(_p(resps[i:i + 1]), _p(flat), len(tris), _p(out_t), _p(out_hit), 0)
        for j in range(len(tris)):
            w = oracle.detect(resps[i:i + 1], tris[j])
            assert (w is not None) == (out_hit[j] == 1)
            if w is not None:
                total_hits += 1
                assert bits(np.float32(w["time_to"])) == bits(out_t[j])
    assert total_hits > 200


def test_detect_x_equals_oracle_bitwise(oracle):
    """xp_detect_x (comparison-based face interval and edge parameter, one vertex division, accumulated range
    guards, literal form when `bad`) against the oracle on the same adversarial cases as the literal form --
    NaN / inf / zero movement / overflow / lattice ties included.  The host build substitutes IEEE division and
    square root for the device sequences, so this checks the RESTRUCTURING and the guards; the device
    sequences themselves are checked against __fdiv_rn / __fsqrt_rn on the GPU (tests/test_gpu_arith.py)."""
    E = emul_lib()
    E.emul_bad_count(1)
    n_hit = n = 0
    for seed, count in ((11, 30000), (23, 30000)):
        for c, m, tri in _cases(seed, count):
            resp = oracle.make_responses([c], [m])
            want = oracle.detect(resp, tri)
            r, got = _emul_detect(E, resp, tri, mode=2)
            n += 1
            assert (want is not None) == (r == 1), (c, m, tri)
            if want is not None:
                n_hit += 1
                w = np.concatenate([[want["time_to"], want["distance_to"]], want["position"], want["intersection"]])
                assert_bits_equal(got, w.astype(np.float32), "collision fields")
    bad = E.emul_bad_count(1)
    print(f"x path: {n} adversarial cases, {n_hit} hits, {bad} took the literal form")
    assert n_hit > 2000
    assert bad < 0.6 * n             # kinds 6-10 (zero movement, 1e20 coordinates, NaN, ...) are SUPPOSED to fall back


def test_detect_x_on_terrain_without_fallback(oracle):
    """On an ordinary scene every pair stays on the guarded path (no fallback) and every result is the oracle's."""
    E = emul_lib()
    tris, h = scenes.sine_terrain(24, 24, seed=3)
    resps = scenes.terrain_spheres(h, 96, seed=5)
    flat = np.ascontiguousarray(tris.reshape(-1, 9))
    out_t, out_hit = np.zeros(len(tris), np.float32), np.zeros(len(tris), np.int32)
    ref_t, ref_hit = np.zeros(len(tris), np.float32), np.zeros(len(tris), np.int32)
    E.emul_bad_count(1)
    hits = 0
    for i in range(len(resps)):
        E.emul_detect_many(_p(resps[i:i + 1]), _p(flat), len(tris), _p(ref_t), _p(ref_hit), 0)   # literal form == oracle (above)
        E.emul_detect_many(_p(resps[i:i + 1]), _p(flat), len(tris), _p(out_t), _p(out_hit), 2)
        assert np.array_equal(out_hit, ref_hit)
        assert np.array_equal(out_t.view(np.uint32), ref_t.view(np.uint32))
        hits += int((ref_hit == 1).sum())
    assert hits > 1000
    assert E.emul_bad_count(1) <= 2, "ordinary terrain pairs must not need the literal form"


def test_detect_f_within_tolerance(oracle):
    """xp_detect_f (contracted arithmetic, half-b quadratics) on terrain: same hit / miss as the literal form except
    on a small fraction of grazing pairs, times of impact within 1e-5 relative on all but a small tail."""
    E = emul_lib()
    tris, h = scenes.sine_terrain(24, 24, seed=3)
    resps = scenes.terrain_spheres(h, 96, seed=5)
    flat = np.ascontiguousarray(tris.reshape(-1, 9))
    out_t, out_hit = np.zeros(len(tris), np.float32), np.zeros(len(tris), np.int32)
    ref_t, ref_hit = np.zeros(len(tris), np.float32), np.zeros(len(tris), np.int32)
    n = differ = beyond = both = 0
    worst = 0.0
    for i in range(len(resps)):
        E.emul_detect_many(_p(resps[i:i + 1]), _p(flat), len(tris), _p(ref_t), _p(ref_hit), 0)
        E.emul_detect_many(_p(resps[i:i + 1]), _p(flat), len(tris), _p(out_t), _p(out_hit), 3)
        n += len(tris)
        differ += int((out_hit != ref_hit).sum())
        b = (out_hit == 1) & (ref_hit == 1)
        rel = np.abs(out_t[b] - ref_t[b]) / np.maximum(1.0, np.abs(ref_t[b]))
        both += int(b.sum())
        beyond += int((rel > 1e-5).sum())
        worst = max(worst, float(rel.max()) if len(rel) else 0.0)
    print(f"f path (host emulation): {n} pairs, hit/miss differs on {differ}, {beyond} of {both} common hits beyond 1e-5 (worst {worst:.2e})")
    assert differ < 2e-4 * n
    assert beyond < 0.01 * both and worst < 5e-3


def test_respond_equals_oracle_bitwise(oracle):
    E = emul_lib()
    rng = np.random.default_rng(7)
    n_ok = n_assert = 0
    for c, m, tri in _cases(13, 20000):
        m = (m * rng.uniform(0.2, 1.5)).astype(np.float32)
        resp = oracle.make_responses([c], [m])
        col = oracle.detect(resp, tri)
        if col is None:
            continue
        want, wn = oracle.respond(resp, np.array([col]))
        out = np.zeros(7, np.float32)
        nrm = np.zeros(3, np.float32)
        r = E.emul_respond(_p(resp), _p(np.array([col])), _p(out), _p(nrm))
        assert_bits_equal(nrm, wn, "slide normal")
        if want is None:
            assert r == 1
            n_assert += 1
        else:
            assert r == 0
            n_ok += 1
            assert_bits_equal(out, np.concatenate([want["c"], [want["r"]], want["movement"]]).astype(np.float32), "response")
    assert n_ok > 500 and n_assert > 10


def test_slab_forms_agree_with_predicate_P(oracle):
    """header form (explicit m==0 branch) == branch-free kernel form == oracle's P, incl. boundary cases."""
    E = emul_lib()
    rng = np.random.default_rng(17)
    n_acc = 0
    for it in range(40000):
        tri = rng.integers(-3, 4, (3, 3)).astype(np.float32) if it % 2 else rng.uniform(-3, 3, (3, 3)).astype(np.float32)
        c = rng.integers(-5, 6, 3).astype(np.float32) if it % 3 == 0 else rng.uniform(-5, 5, 3).astype(np.float32)
        m = rng.uniform(-1, 1, 3).astype(np.float32)
        for k in range(3):
            if rng.random() < 0.25:
                m[k] = 0.0 if rng.random() < 0.5 else -0.0
        if it % 5 == 0:      # centre exactly on an inflated face of the box
            lo, hi = oracle.triangle_aabbs(tri[None])
            k = rng.integers(0, 3)
            c[k] = lo[0, k] if rng.random() < 0.5 else hi[0, k]
        resp = oracle.make_responses([c], [m])
        for horizon in (np.inf, 1.0):
            want = oracle.lib().xpo_broadphase_accept(_p(resp), _p(np.ascontiguousarray(tri)), horizon)
            got = E.emul_broadphase(_p(resp), _p(np.ascontiguousarray(tri)), horizon)
            assert got == (3 if want else 0), (c, m, tri, horizon, want, got)
            n_acc += want
    assert n_acc > 5000


def test_P_is_conservative(oracle):
    """every (sphere, triangle) with detect != None satisfies P(horizon = inf)."""
    tris, h = scenes.sine_terrain(32, 32, seed=9)
    resps = scenes.terrain_spheres(h, 200, seed=10)
    si, ti = oracle.broadphase_pairs(resps, tris)
    P = set(zip(si.tolist(), ti.tolist()))
    hits = 0
    for i in range(len(resps)):
        for j in range(len(tris)):
            if oracle.detect(resps[i:i + 1], tris[j]) is not None:
                hits += 1
                assert (i, j) in P
    assert hits > 1000
    assert len(P) < 0.2 * len(resps) * len(tris)


def _grid_eps(x0, z0, sp, nx, nz):
    """xp_scene_set_heightmap's choice (xp_api.cu)"""
    big = max(abs(x0) + sp * nx, abs(z0) + sp * nz) + 2.0
    err = 32.0 * 1.1920929e-7 * big
    assert err < 0.25 * sp
    return float(np.float32(err))


@pytest.mark.parametrize("case", ["config2-like", "offset-fine", "coarse", "far-origin"])
def test_grid_walk_enumerates_exactly_predicate_P(oracle, case):
    """The height-grid walk (csrc/xp_gridwalk.cuh, the functions k_broadphase_grid runs) against the oracle's
    predicate P over the emitted triangle soup: the same candidate set for every ray -- rays inside, outside,
    along grid lines, with zero components, vertical, pointing away, unit and infinite horizon."""
    E = emul_lib()
    nx, nz, x0, z0, sp, seed = {"config2-like": (40, 37, 0.0, 0.0, 1.0, 3), "offset-fine": (33, 29, -5.5, 2.25, 0.37, 4),
                                "coarse": (9, 12, 100.0, -50.0, 7.5, 5), "far-origin": (24, 24, 4000.0, -3000.0, 1.0, 6)}[case]
    h = scenes.sine_terrain_heights(nx, nz, seed=seed)
    tris = scenes.heightmap_from_heights(h, x0=x0, z0=z0, spacing=sp)
    eps = _grid_eps(x0, z0, sp, nx, nz)
    rng = np.random.default_rng(seed)
    n = 400
    resps = scenes.terrain_spheres(h, n, seed=seed + 10)
    resps["c"][:, 0] = resps["c"][:, 0] * sp + x0
    resps["c"][:, 2] = resps["c"][:, 2] * sp + z0
    k = np.arange(n)
    resps["movement"][k % 9 == 1, 0] = 0.0                            # movement inside a z-column
    resps["movement"][k % 9 == 2, 2] = 0.0
    resps["movement"][k % 9 == 3, 0::2] = 0.0                         # vertical
    resps["movement"][k % 9 == 4] = 0.0                               # zero movement
    resps["movement"][k % 9 == 5, 1] *= -1.0                          # upwards
    resps["c"][k % 9 == 6] += rng.uniform(-3, 3, ((k % 9 == 6).sum(), 3)).astype(np.float32) * np.float32([nx * sp, 1, nz * sp])   # outside
    on_line = k % 9 == 7                                              # centres exactly on grid lines / inflated faces
    resps["c"][on_line, 0] = (x0 + sp * rng.integers(0, nx + 1, on_line.sum())).astype(np.float32)
    resps["c"][k % 9 == 8, 2] = np.float32(z0 - 1.0001)
    resps["movement"][k % 9 == 8, 2] = 0.0
    hf = np.ascontiguousarray(h, np.float32)
    cap = 1 << 16
    out = np.zeros(cap, np.uint32)
    cells = C.c_long(0)
    total = tested = 0
    for horizon in (np.inf, 1.0):
        wi, wj = oracle.broadphase_pairs(resps, tris, horizon=horizon)
        for i in range(n):
            cnt = E.emul_grid_walk(_p(hf), nx, nz, x0, z0, sp, eps, _p(resps[i:i + 1]), horizon, _p(out), cap, C.byref(cells))
            assert cnt <= cap
            got = np.sort(out[:cnt])
            want = np.sort(wj[wi == i])
            assert np.array_equal(got, want), (case, horizon, i, resps[i], len(got), len(want))
            total += cnt
            tested += cells.value
    print(f"{case}: {total} candidates, {2 * tested} triangle boxes tested ({2 * tested / max(total, 1):.2f} per candidate)")
    assert total > 800
