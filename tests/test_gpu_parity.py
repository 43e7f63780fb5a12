"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs and against the committed golden vectors.

Bars (BASELINE.json north_star / SURVEY.md Appendix F):
  * broadphase candidate set: bit-exact (sorted pair lists identical to predicate P);
  * XP_ARITH_EXACT: every output bit-identical to the oracle (t, position, intersection, resolved
    Response, slide normal), hit / tri_index / iterations / status equal -- zero mismatches;
  * XP_ARITH_FAST: hit/miss equal except grazing rows (oracle float64 shadow margin < 1e-5),
    time of impact within 1e-5 * max(1, |t|), positions within 1e-5 * max(1, |c|_inf).
"""
import os

import numpy as np
import pytest

import xp_physics_cuda as xp
from tests.conftest import assert_bits_equal
from xp_physics_cuda import scenes

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
METHODS = [("lbvh_pairs", True), ("q4_pairs", False), ("wide_fused", False), ("wide_fused", True), ("wide_pairs", False), ("lbvh_fused", False),
           ("lbvh_fused", True), ("lbvh_pairs", False), ("brute", False)]


def col_fields(col):
    return np.concatenate([col["time_to"][:, None], col["distance_to"][:, None], col["position"],
                           col["intersection"]], axis=1)


def resp_fields(r):
    return np.concatenate([r["c"], r["r"][:, None], r["movement"]], axis=1)


def check_detect_exact(got, want, what, tri_index=True):
    gcol, ghit, gtri, gst = got
    wcol, whit, wtri, wst = want[:4]
    assert np.array_equal(ghit, whit), f"{what}: hit/miss differs on {np.flatnonzero(ghit != whit)[:8]}"
    h = whit.astype(bool)
    assert_bits_equal(col_fields(gcol)[h], col_fields(wcol)[h], what)
    if tri_index:
        assert np.array_equal(gtri, wtri), f"{what}: tri_index differs"
    assert np.array_equal(gst, wst), f"{what}: status differs"


@pytest.fixture(scope="module")
def terrain(oracle):
    tris, h = scenes.sine_terrain(40, 40, seed=31)
    resps = scenes.terrain_spheres(h, 3000, seed=32)
    want = oracle.detect_batch(resps, tris)
    return tris, h, resps, want


# ---- known answers of the reference's own tests, on the device -------------------------------
def test_reference_kats_on_device():
    T, R, S = xp.Triangle, xp.Response, xp.Sphere
    c = xp.sphere_triangle_detect_collision(R(S([0, 4, 0], 1.0), [0, -2, 0]), T([-2, 2, -2], [-2, 2, 2], [2, 2, 0]))
    assert c.time_to == 0.5                                                  # collision_detect.rs:222-232
    assert c.distance_to == 1.0 and c.position == [0, 3, 0] and c.intersection == [0, 2, 0]
    c = xp.sphere_triangle_detect_collision(R(S([0, 4, 0], 1.0), [0, -8, 0]), T([0, 0, 0], [-2, -1, 0], [2, -1, 0]))
    assert c.time_to == 0.375                                                # collision_detect.rs:235-246
    c = xp.sphere_triangle_detect_collision(R(S([0, 4, 0], 1.0), [0, -8, 0]), T([0, -2, 0], [-2, -1, 0], [2, -1, 0]))
    assert c.time_to == 0.5                                                  # collision_detect.rs:249-260
    c = xp.sphere_triangle_detect_collision(R(S([0, 40, 0], 1.0), [0, -8, 0]), T([0, 0, 0], [-2, -1, 0], [2, -1, 0]))
    assert c.time_to == 4.875                                                # quirk B2
    assert xp.sphere_triangle_detect_collision(R(S([0, 0, 0], 1.0), [0, 2, 0]), T([-2, 2, -2], [-2, 2, 2], [2, 2, 0])) is None
    with pytest.raises(AssertionError):                                      # collision_detect.rs:161
        xp.sphere_triangle_detect_collision(R(S([0, 4, 0], 0.5), [0, -2, 0]), T([-2, 2, -2], [-2, 2, 2], [2, 2, 0]))


def test_point_in_triangle_kat_through_the_face_path(oracle):
    # triangle.rs:62-76: the six points, reached through detect's face path (sphere dropped along -n)
    tri = np.array([[-1, 0, 0], [0, 1, 1], [1, 0, 0]], np.float32)
    n = np.cross(tri[1] - tri[0], tri[2] - tri[0])
    n = n / np.linalg.norm(n)
    pts = [[0, .5, .5], [0, .6, .6], [0, .99, .99], [0, 1.001, 1.001], [0, 0, 0], [0, 1, 1]]
    resps = xp.make_responses([np.asarray(p) + 3 * n for p in pts], [-2.5 * n] * len(pts))
    with xp.Scene(0, tri[None]) as s:
        got = s.detect_batch(resps)
    check_detect_exact(got, oracle.detect_batch(resps, tri[None]), "pit kat")


def test_single_response_call_matches_oracle(oracle):
    tri = xp.Triangle([-2, 2, -2], [-2, 2, 2], [2, 2, 0])
    r = xp.Response(xp.Sphere([0.1, 3.2, 0.05]), [0.3, -0.5, 0.1])
    col = xp.sphere_triangle_detect_collision(r, tri)
    out = xp.sphere_triangle_calculate_response(r, col)
    want, _ = oracle.respond(r.as_record(), col.as_record())
    assert_bits_equal(np.array(out.sphere.c + [out.sphere.r] + out.movement, np.float32), resp_fields(np.array([want]))[0], "respond")
    bad = xp.Response(xp.Sphere([0, 4, 0]), [0, -4, 0])
    with pytest.raises(AssertionError):                                      # collision_response.rs:22
        xp.sphere_triangle_calculate_response(bad, xp.sphere_triangle_detect_collision(bad, tri))


def test_reference_free_functions(oracle):
    tri = np.array([[[-2, 2, -2], [-2, 2, 2], [2, 2, 0]]], np.float32)
    r = xp.Response(xp.Sphere([0.1, 3.2, 0.05]), [0.3, -0.5, 0.1])
    want, it, st, _ = oracle.collision_response(r.as_record(), tri)
    out = xp.collision_response(r, tri)
    assert_bits_equal(np.array(out.sphere.c + [out.sphere.r] + out.movement, np.float32), resp_fields(np.array([want]))[0], "collision_response")
    out2 = xp.collision_response_non_trianulated(r, tri.reshape(-1, 3))
    assert out2 == out
    with pytest.raises(IndexError):                                          # lib.rs:19-24
        xp.collision_response_non_trianulated(r, np.zeros((4, 3), np.float32))
    with pytest.raises(AssertionError):
        xp.collision_response(xp.Response(xp.Sphere([0, 4, 0]), [0, -4, 0]), tri)


# ---- detect: every method, exact arithmetic, bit-identical -----------------------------------
@pytest.mark.parametrize("method,prune", METHODS)
@pytest.mark.parametrize("sort", [True, False])
def test_detect_exact_bitwise(terrain, method, prune, sort):
    tris, h, resps, want = terrain
    with xp.Scene(0, tris, method=method, prune=prune, sort_spheres=sort) as s:
        got = s.detect_batch(resps)
        st = s.stats()
    check_detect_exact(got, want, f"{method} prune={prune} sort={sort}")
    if method == "brute":
        assert st["narrowphase_tests"] == len(resps) * len(tris)
    assert st["kernel_launches"] > 0


def test_broadphase_pairs_bit_exact(terrain, oracle):
    tris, h, resps, _ = terrain
    for horizon in ("inf", "unit"):
        wi, wj = oracle.broadphase_pairs(resps, tris, np.inf if horizon == "inf" else 1.0)
        w = np.sort(wi.astype(np.uint64) << 32 | wj)
        for method in ("wide_fused", "lbvh_fused", "q4_pairs"):      # every tree must yield exactly the set P
            with xp.Scene(0, tris, method=method) as s:
                gi, gj = s.broadphase_pairs(resps, horizon=horizon)
                s.detect_batch(resps, horizon=horizon)
                st_ = s.stats()
            g = np.sort(gi.astype(np.uint64) << 32 | gj)
            assert np.array_equal(w, g), f"candidate set differs ({horizon}, {method}): {len(w)} vs {len(g)}"
            assert st_["narrowphase_tests"] == len(w)     # detect is evaluated on exactly the pairs passing P
            if method == "q4_pairs":                       # stage 1 emitted a (small) superset, stage 2 filtered it
                assert len(w) <= st_["broadphase_pairs"] < 1.3 * len(w)


def test_golden_vectors():
    g = np.load(os.path.join(GOLDEN, "golden_oracle.npz"))
    tris, resps = g["triangles"], g["responses"]
    with xp.Scene(0, tris) as s:
        got = s.detect_batch(resps)
        check_detect_exact(got, (g["det_col"], g["det_hit"], g["det_tri"], np.zeros(len(resps), np.uint32)), "golden detect")
        out, it, st, ln = s.collision_response_batch(resps)
        gi, gj = s.broadphase_pairs(resps)
    assert_bits_equal(resp_fields(out), resp_fields(g["resp_out"]), "golden response")
    assert np.array_equal(it, g["resp_iter"]) and np.array_equal(st, g["resp_status"])
    assert_bits_equal(ln, g["resp_normal"], "golden slide normal")
    assert np.array_equal(np.sort(gi.astype(np.uint64) << 32 | gj),
                          np.sort(g["pair_sphere"].astype(np.uint64) << 32 | g["pair_tri"]))


def test_pair_buffer_overflow_is_recovered(terrain):
    """a tiny candidate buffer forces the optimistic two-stage sweep to overflow; the library must
    notice, redo the sweep in its synchronous chunk-halving mode and return the same answer."""
    tris, h, resps, want = terrain
    for method in ("lbvh_pairs", "wide_pairs", "q4_pairs"):
        with xp.Scene(0, tris, method=method) as s:
            s.set_options(max_pairs=4096)
            got = s.detect_batch(resps)
            check_detect_exact(got, want, f"overflow recovery {method}")
            st_ = s.stats()
            assert st_["narrowphase_tests"] <= st_["broadphase_pairs"]
            if method == "wide_pairs":                     # (q4 and the leaf runs of lbvh_pairs emit a superset and filter it)
                assert st_["narrowphase_tests"] == st_["broadphase_pairs"]
            out, it, st, ln = s.collision_response_batch(resps[:300], max_iter=2)
            assert it.max() == 2


@pytest.mark.parametrize("scene_kind", ["soup", "heightmap"])
def test_pruning_changes_no_contact(oracle, scene_kind):
    """XP_OPT_PRUNE = 1 on the two-stage sweep (distance windows; a heightmap scene then walks its LBVH): fewer tests,
    the same closest collision, bit for bit."""
    h = scenes.sine_terrain_heights(90, 80, seed=21)
    tris = scenes.heightmap_from_heights(h)
    resps = scenes.terrain_spheres(h, 20000, seed=22)
    want = oracle.detect_batch(resps[:3000], tris)
    with xp.Scene(0, prune=False) as s:
        if scene_kind == "heightmap":
            s.set_heightmap(h)
        else:
            s.set_triangles(tris)
        s.build()
        a = s.detect_batch(resps)
        tests_all = s.stats()["narrowphase_tests"]
        s.set_options(prune=True)
        b = s.detect_batch(resps)
        tests_pruned = s.stats()["narrowphase_tests"]
    check_detect_exact(tuple(x[:3000] for x in a), want, "exhaustive")
    check_detect_exact(b, a, f"pruned ({scene_kind})", tri_index=False)
    assert tests_pruned < tests_all


def test_compact_results_rebuild_the_full_collision(terrain, oracle):
    """xp_detect_batch_compact (8 B per sphere: time_to, tri_index) + collision_from_hit == xp_detect_batch, bit for bit;
    the asynchronous compact form shares tickets with the full one."""
    import torch
    tris, h, resps, want = terrain
    with xp.Scene(0, tris) as s:
        full = s.detect_batch(resps)
        hits, st = s.detect_batch_compact(resps)
        n = len(resps)
        h_in = torch.from_numpy(resps.view(np.float32).reshape(n, 7).copy()).pin_memory()
        h_out = [torch.zeros((n, 2), dtype=torch.int32).pin_memory() for _ in range(2)]
        t0 = s.detect_batch_compact_submit(h_in.data_ptr(), n, h_out[0].data_ptr())
        t1 = s.detect_batch_compact_submit(h_in.data_ptr(), n, h_out[1].data_ptr())
        s.detect_batch_wait(t0); s.detect_batch_wait(t1)
    col, hit = xp.collision_from_hit(resps, hits)
    check_detect_exact((col, hit, hits["tri_index"], st), want, "compact + collision_from_hit")
    check_detect_exact((col, hit, hits["tri_index"], st), full, "compact vs full records")
    assert (hits["time_to"][hit == 0] == 0).all()
    for o in h_out:
        assert np.array_equal(o.numpy().view(np.uint32), hits.view(np.uint32).reshape(n, 2))


def test_host_pipeline_chunks(terrain, monkeypatch):
    """the host-pointer sweep cut into H2D / compute / D2H pipeline chunks gives the same answer as one piece"""
    tris, h, resps, want = terrain
    with xp.Scene(0, tris) as s:
        for k in ("1", "3", "7"):
            monkeypatch.setenv("XP_IO_CHUNKS", k)
            got = s.detect_batch(resps)
            check_detect_exact(got, want, f"{k} pipeline chunks")
            assert 0 < s.stats()["narrowphase_tests"] <= s.stats()["broadphase_pairs"]
    monkeypatch.delenv("XP_IO_CHUNKS")


# ---- response loop ---------------------------------------------------------------------------
@pytest.mark.parametrize("method,prune", METHODS)
def test_collision_response_exact_bitwise(terrain, oracle, method, prune):
    tris, h, resps, _ = terrain
    resps = resps[:1500]
    for max_iter in (0, 1, 4):
        want_out, want_it, want_st, want_n, _ = oracle.collision_response_batch(resps, tris, max_iter=max_iter)
        with xp.Scene(0, tris, method=method, prune=prune) as s:
            out, it, st, ln = s.collision_response_batch(resps, max_iter=max_iter)
        what = f"{method} prune={prune} max_iter={max_iter}"
        assert np.array_equal(it, want_it), what
        assert np.array_equal(st, want_st), what
        assert_bits_equal(resp_fields(out), resp_fields(want_out), what)
        assert_bits_equal(ln, want_n, what + " slide normal")
    assert want_it.max() >= 3


def test_response_loop_runs_on_the_device_and_survives_overflow(terrain, oracle, monkeypatch):
    """The device-resident loop (no host read-back between iterations; live counts, next rays and lists stay on
    the GPU) gives the synchronous loop's answer bit for bit -- bounded, unbounded (one look per 8 iterations), and
    when a sweep overflows the optimistic candidate buffer mid-loop (the call is redone from a copy of the input)."""
    tris, h, resps, _ = terrain
    for max_iter in (0, 2, 5, 11):
        w = oracle.collision_response_batch(resps, tris, max_iter=max_iter)
        with xp.Scene(0, tris, timing=True) as s:
            out, it, st, ln = s.collision_response_batch(resps, max_iter=max_iter)
            stats = s.stats()
        assert np.array_equal(it, w[1]) and np.array_equal(st, w[2])
        assert_bits_equal(resp_fields(out), resp_fields(w[0]), f"device loop max_iter={max_iter}")
        assert_bits_equal(ln, w[3], "slide normal")
        assert w[1].max() <= stats["iterations_run"] <= w[1].max() + 1
    w = oracle.collision_response_batch(resps, tris, max_iter=4)
    with xp.Scene(0, tris) as s:
        s.set_options(max_pairs=4096)                      # every sweep overflows: the redo path
        out, it, st, ln = s.collision_response_batch(resps, max_iter=4)
    assert np.array_equal(it, w[1]) and np.array_equal(st, w[2])
    assert_bits_equal(resp_fields(out), resp_fields(w[0]), "device loop, overflow redo")
    monkeypatch.setenv("XP_RESPONSE_SYNC", "1")             # read once per process: only effective if set before the first call
    with xp.Scene(0, tris) as s:
        out2 = s.collision_response_batch(resps, max_iter=4)
    monkeypatch.delenv("XP_RESPONSE_SYNC")
    assert_bits_equal(resp_fields(out2[0]), resp_fields(w[0]), "loop (either form)")


def test_response_status_bits(oracle):
    tri = np.array([[[-2, 2, -2], [-2, 2, 2], [2, 2, 0]]], np.float32)
    resps = xp.make_responses([[0, 4, 0], [0, 4, 0], [0.1, 3.2, 0.05], [0, 9, 0]],
                              [[0, -4, 0], [0, -2, 0], [0.3, -0.5, 0.1], [0, 1, 0]], [1.0, 0.5, 1.0, 1.0])
    want_out, want_it, want_st, want_n, _ = oracle.collision_response_batch(resps, tri)
    with xp.Scene(0, tri) as s:
        out, it, st, ln = s.collision_response_batch(resps)
    assert list(st) == [xp._lib.ST_ASSERT_NEG_DISTANCE, xp._lib.ST_RADIUS_NE_1, 0, 0] == list(want_st)
    assert np.array_equal(it, want_it)
    assert_bits_equal(resp_fields(out), resp_fields(want_out), "status cases")
    assert_bits_equal(ln, want_n, "status normals")
    with xp.Scene(0, tri) as s:
        out1, it1, st1, _ = s.collision_response_batch(resps[2:3], max_iter=1)
    w = oracle.collision_response_batch(resps[2:3], tri, max_iter=1)
    assert list(st1) == list(w[2]) and list(it1) == list(w[1])


# ---- edge cases ------------------------------------------------------------------------------
def test_empty_and_tiny_inputs(oracle):
    tris, h = scenes.sine_terrain(4, 4, seed=2)
    resps = scenes.terrain_spheres(h, 50, seed=3)
    with xp.Scene(0, tris) as s:
        col, hit, tri, st = s.detect_batch(resps[:0])
        assert len(col) == 0
        out, it, st, ln = s.collision_response_batch(resps[:0])
        assert len(out) == 0
    with xp.Scene(0, tris[:0]) as s:                       # no triangles at all
        col, hit, tri, st = s.detect_batch(resps)
        assert not hit.any() and (tri == 0xFFFFFFFF).all()
        out, it, st, ln = s.collision_response_batch(resps)
        assert_bits_equal(resp_fields(out), resp_fields(resps), "no triangles: unchanged")
        assert not it.any()
        # r != 1 against an EMPTY slice: the assert of collision_detect.rs:161 sits inside the per-triangle call,
        # which the loop of lib.rs:33-35 never reaches -- the reference returns the Response unchanged
        odd = resps.copy()
        odd["r"][::2] = 0.5
        col, hit, tri, st = s.detect_batch(odd)
        w = oracle.detect_batch(odd, tris[:0])
        assert not hit.any() and np.array_equal(st, w[3]) and not st.any()
        out, it, st, ln = s.collision_response_batch(odd)
        w = oracle.collision_response_batch(odd, tris[:0])
        assert np.array_equal(st, w[2]) and not st.any() and not it.any()
        assert_bits_equal(resp_fields(out), resp_fields(odd), "no triangles, r != 1: unchanged")
    r = xp.Response(xp.Sphere([0, 3, 0], 0.5), [0, -1, 0])
    assert xp.collision_response(r, tris[:0]) == r            # no panic, like the reference
    for n in (1, 2, 3, 5):
        with xp.Scene(0, tris[:n]) as s:
            got = s.detect_batch(resps)
            gi, gj = s.broadphase_pairs(resps)
        check_detect_exact(got, oracle.detect_batch(resps, tris[:n]), f"{n} triangles")
        wi, wj = oracle.broadphase_pairs(resps, tris[:n])
        assert np.array_equal(np.sort(gi.astype(np.uint64) << 32 | gj), np.sort(wi.astype(np.uint64) << 32 | wj))


def test_quirk_inputs(oracle):
    """zero / axis-aligned movement, degenerate and duplicated triangles, r != 1, embedded spheres."""
    rng = np.random.default_rng(5)
    tris, h = scenes.sine_terrain(12, 12, seed=6)
    tris = np.concatenate([tris, tris[:40], np.repeat(tris[50:51, :1], 3, axis=1)])      # duplicates + zero-area
    tris[7, 2] = tris[7, 1]                                                                # degenerate
    resps = scenes.terrain_spheres(h, 600, seed=7)
    resps["movement"][::7] = 0.0
    resps["movement"][1::7, 0] = 0.0
    resps["movement"][2::7, [0, 2]] = 0.0
    resps["movement"][3::7, 1] = -0.0
    resps["c"][4::7, 1] -= 1.5                                                             # start embedded
    resps["r"][5::50] = 0.5
    resps["c"][6::7] = np.round(resps["c"][6::7])                                          # on lattice planes
    want = oracle.detect_batch(resps, tris)
    for method, prune in METHODS:
        with xp.Scene(0, tris, method=method, prune=prune) as s:
            got = s.detect_batch(resps)
            gi, gj = s.broadphase_pairs(resps)
        check_detect_exact(got, want, f"quirks {method} {prune}", tri_index=not prune)
    wi, wj = oracle.broadphase_pairs(resps, tris)
    valid = resps["r"][gi] == 1.0
    assert np.array_equal(np.sort(gi[valid].astype(np.uint64) << 32 | gj[valid]),
                          np.sort(wi[resps["r"][wi] == 1.0].astype(np.uint64) << 32 | wj[resps["r"][wi] == 1.0]))
    w = oracle.collision_response_batch(resps, tris, max_iter=6)
    with xp.Scene(0, tris) as s:
        out, it, st, ln = s.collision_response_batch(resps, max_iter=6)
    assert np.array_equal(it, w[1]) and np.array_equal(st, w[2])
    assert_bits_equal(resp_fields(out), resp_fields(w[0]), "quirk response")


def test_non_finite_and_extreme_spheres(oracle):
    """NaN / astronomically large / vanishing sphere inputs: same answers as the oracle's brute force.
    (Contract, include/xp_physics_cuda.h: triangles are finite; a sphere with an INFINITE component is outside
    the contract, and so is a non-zero movement so small that |m|^2 underflows to 0 (|m| < ~1e-19) -- the
    reference turns both into "hits" at t = +inf with infinite positions, which no bounded query volume contains.)"""
    rng = np.random.default_rng(11)
    tris, h = scenes.sine_terrain(14, 14, seed=12)
    resps = scenes.terrain_spheres(h, 900, seed=13)
    c, m = resps["c"], resps["movement"]
    c[0::30, rng.integers(0, 3)] = np.nan
    m[3::30, rng.integers(0, 3)] = np.nan
    m[7::30] *= np.float32(1e-15)                         # tiny but |m|^2 still representable (see contract)
    m[8::30] *= np.float32(1e25)                          # b, det overflow
    c[9::30] *= np.float32(1e20)
    m[10::30] = 0.0
    want = oracle.detect_batch(resps, tris)
    w = oracle.collision_response_batch(resps, tris, max_iter=3)
    for method, prune in METHODS:
        with xp.Scene(0, tris, method=method, prune=prune) as s:
            got = s.detect_batch(resps)
            out, it, st, ln = s.collision_response_batch(resps, max_iter=3)
        check_detect_exact(got, want, f"non-finite {method} {prune}", tri_index=not prune)
        assert np.array_equal(it, w[1]) and np.array_equal(st, w[2]), method
        assert_bits_equal(resp_fields(out), resp_fields(w[0]), f"non-finite response {method}")


def test_random_soups_fuzz(oracle):
    """many small random triangle soups (arbitrary orientation, slivers, duplicates, zero-area triangles, 1..400
    triangles) x random spheres: detect, broadphase pairs and a 3-iteration response against the oracle."""
    rng = np.random.default_rng(2024)
    total_hits = 0
    with xp.Scene(0) as s:
        for case in range(48):
            nt = int(rng.choice([1, 2, 3, 7, 33, 64, 65, 200, 400]))
            ext = float(rng.choice([2.0, 8.0, 40.0]))
            base = rng.uniform(-ext, ext, (nt, 1, 3))
            tris = (base + rng.normal(0, rng.choice([0.05, 1.0, 4.0]), (nt, 3, 3))).astype(np.float32)
            if nt > 3:
                tris[1] = tris[0]                              # duplicate
                tris[2, 2] = tris[2, 1]                        # zero area
                tris[3, 1] = (tris[3, 0] + tris[3, 2]) / 2     # collinear sliver
            ns = 150
            c = rng.uniform(-ext - 3, ext + 3, (ns, 3)).astype(np.float32)
            m = rng.normal(0, 1, (ns, 3))
            m = (m / np.linalg.norm(m, axis=1, keepdims=True) * rng.uniform(0.01, 0.95, (ns, 1))).astype(np.float32)
            m[::17, rng.integers(0, 3)] = 0.0
            resps = xp.make_responses(c, m)
            s.set_options(method=str(rng.choice(["lbvh_pairs", "q4_pairs", "wide_pairs", "lbvh_fused", "wide_fused"])),
                          prune=bool(rng.integers(0, 2)))
            s.set_triangles(tris)
            s.build()
            got = s.detect_batch(resps)
            want = oracle.detect_batch(resps, tris)
            check_detect_exact(got, want, f"fuzz case {case}", tri_index=False)
            total_hits += int(want[1].sum())
            gi, gj = s.broadphase_pairs(resps)
            wi, wj = oracle.broadphase_pairs(resps, tris)
            assert np.array_equal(np.sort(gi.astype(np.uint64) << 32 | gj), np.sort(wi.astype(np.uint64) << 32 | wj)), case
            out, it, st, ln = s.collision_response_batch(resps, max_iter=3)
            w = oracle.collision_response_batch(resps, tris, max_iter=3)
            assert np.array_equal(it, w[1]) and np.array_equal(st, w[2]), case
            assert_bits_equal(resp_fields(out), resp_fields(w[0]), f"fuzz response {case}")
    assert total_hits > 500


def test_large_coordinates(oracle):
    tris, h = scenes.sine_terrain(16, 16, seed=8)
    tris = tris + np.float32([4000, 0, 3000])
    resps = scenes.terrain_spheres(h, 400, seed=9)
    resps["c"] += np.float32([4000, 0, 3000])
    with xp.Scene(0, tris) as s:
        got = s.detect_batch(resps)
    check_detect_exact(got, oracle.detect_batch(resps, tris), "large coordinates")


# ---- fast (FMA-contracted) arithmetic: tolerance + grazing protocol ----------------------------
def test_fast_mode_within_tolerance(terrain, oracle):
    tris, h, resps, want = terrain
    wcol, whit, wtri, wst = want[:4]
    with xp.Scene(0, tris, arith="fast") as s:
        col, hit, tri, st = s.detect_batch(resps)
        out, it, rst, ln = s.collision_response_batch(resps[:1000], max_iter=1)
    margins = oracle.grazing_margins(resps, tris)
    grazing = margins < 1e-5                                   # the stated epsilon
    differ = hit != whit
    assert not (differ & ~grazing).any(), "hit/miss differs on non-grazing rows"
    assert grazing.mean() < 0.05          # ~37 candidates x ~10 branch operands per sphere
    both = (hit & whit).astype(bool) & ~grazing
    t, wt = col["time_to"][both], wcol["time_to"][both]
    rel = np.abs(t - wt) / np.maximum(1.0, np.abs(wt))
    # The 1e-5 bar holds wherever the quadratic is well conditioned.  (-b - sqrt(det)) / 2a cancels
    # catastrophically for a few contacts, where fp32 itself (the reference included) only carries
    # ~1e-4; contraction changes those low bits.  Report and bound that tail; XP_ARITH_EXACT, the
    # default, has no such tail (bit-identical, see test_detect_exact_bitwise).
    frac_out = float((rel > 1e-5).mean())
    print(f"fast mode: max rel dt {rel.max():.3e}, fraction beyond 1e-5: {frac_out:.4%}, grazing rows {grazing.mean():.3%}")
    assert frac_out < 0.01 and rel.max() < 2e-3
    tol = 1e-5 * np.maximum(1.0, np.abs(wcol["position"][both]).max(axis=1, keepdims=True))
    pos_out = float((np.abs(col["position"][both] - wcol["position"][both]) > tol).any(axis=1).mean())
    assert pos_out < 0.01
    w = oracle.collision_response_batch(resps[:1000], tris, max_iter=1)
    ok = (it == w[1]) & ~grazing[:1000]
    tol = 1e-5 * np.maximum(1.0, np.abs(w[0]["c"]).max(axis=1, keepdims=True))
    assert float((np.abs(out["c"][ok] - w[0]["c"][ok]) > tol[ok]).any(axis=1).mean()) < 0.01
    assert float((np.abs(ln[ok] - w[3][ok]) > 1e-5).any(axis=1).mean()) < 0.01
    assert (it != w[1])[~grazing[:1000]].sum() == 0


# ---- the benchmark configs as parity cases -----------------------------------------------------
def test_config1_player_vs_heightmap(oracle):
    """config 1: one swept player sphere vs the res/map heightmap crop (80 000 triangles), one frame."""
    red = np.load(os.path.join(GOLDEN, "heightmap_crop_red.npy"))
    tris = scenes.heightmap_from_pixels(red)
    h00 = red[100, 100] / 10.0
    resps = xp.make_responses([[0, h00 + 1.5, 0], [0, h00 + 1.02, 0], [10.5, red[110, 95] / 10.0 + 1.2, -5.25]],
                              [[0.05, -0.05, 0], [0.05, -0.05, 0], [0.03, -0.06, 0.02]])
    want = oracle.detect_batch(resps, tris)
    w = oracle.collision_response_batch(resps, tris)
    for method, prune in METHODS:
        with xp.Scene(0, tris, method=method, prune=prune) as s:
            got = s.detect_batch(resps)
            out, it, st, ln = s.collision_response_batch(resps)
        check_detect_exact(got, want, f"config1 {method}", tri_index=not prune)
        assert np.array_equal(it, w[1]) and np.array_equal(st, w[2])
        assert_bits_equal(resp_fields(out), resp_fields(w[0]), f"config1 response {method}")


def test_small_batch_path_and_its_fall_through(oracle):
    """collision_response for a handful of spheres runs as ONE launch (k_small_response: a warp per sphere).  The same
    batch through that path, through the batch kernels (XP_NO_SMALL_PATH) and through the small path giving up (a work
    list of one entry, XP_SMALL_QUEUE: the call falls through) must agree bit for bit with each other and with the
    oracle -- host call and device-pointer call in place, stats included.  The switches are read once per process."""
    import subprocess
    import sys
    import tempfile
    code = (
        "import numpy as np, torch, xp_physics_cuda as xp\n"
        "from xp_physics_cuda import scenes\n"
        "tris, h = scenes.sine_terrain(48, 48, seed=21)\n"
        "resps = scenes.terrain_spheres(h, 300, seed=22)\n"
        "resps['r'][7] = 2.0\n"
        "with xp.Scene(0, tris) as s:\n"
        "    out, it, st, ln = s.collision_response_batch(resps, max_iter=6)\n"
        "    stats = s.stats()\n"
        "    d = torch.from_numpy(np.ascontiguousarray(resps).view(np.float32).reshape(-1, 7)).cuda()\n"
        "    di = torch.zeros(len(resps), dtype=torch.int32, device='cuda'); ds = torch.zeros_like(di)\n"
        "    s.collision_response_batch_dev(d.data_ptr(), len(resps), d.data_ptr(), 6, di.data_ptr(), ds.data_ptr())\n"
        "    torch.cuda.synchronize()\n"
        "np.savez(r'%s', out=out.view(np.float32), it=it, st=st, ln=ln.view(np.float32), dev=d.cpu().numpy(), di=di.cpu().numpy(),\n"
        "         ds=ds.cpu().numpy(), launches=stats['kernel_launches'], tests=stats['narrowphase_tests'])\n")
    res = {}
    with tempfile.TemporaryDirectory() as dd:
        for tag, env_add in (("small", {}), ("batch", {"XP_NO_SMALL_PATH": "1"}), ("giveup", {"XP_SMALL_QUEUE": "1"})):
            env = dict(os.environ)
            env["PYTHONPATH"] = os.pathsep.join([os.path.dirname(os.path.dirname(xp.__file__)), env.get("PYTHONPATH", "")])
            for k in ("XP_NO_SMALL_PATH", "XP_SMALL_QUEUE"):
                env.pop(k, None)
            env.update(env_add)
            path = os.path.join(dd, tag + ".npz")
            subprocess.run([sys.executable, "-c", code % path], check=True, env=env, timeout=300)
            z = np.load(path)
            res[tag] = {k: z[k] for k in z.files}
    assert res["small"]["launches"] == 1 and res["batch"]["launches"] > 5 and res["giveup"]["launches"] > 5
    for tag in ("batch", "giveup"):
        for k in ("out", "it", "st", "ln", "dev", "di", "ds"):
            assert np.array_equal(res["small"][k].view(np.uint32), res[tag][k].view(np.uint32)), (tag, k)
    assert res["small"]["tests"] == res["batch"]["tests"]                     # the same candidate sets were tested
    assert np.array_equal(res["small"]["out"].view(np.uint32), res["small"]["dev"].view(np.uint32).reshape(res["small"]["out"].shape))
    tris, h = scenes.sine_terrain(48, 48, seed=21)
    resps = scenes.terrain_spheres(h, 300, seed=22)
    resps["r"][7] = 2.0
    w = oracle.collision_response_batch(resps, tris, max_iter=6)
    assert np.array_equal(res["small"]["it"], w[1]) and np.array_equal(res["small"]["st"], w[2])
    assert_bits_equal(res["small"]["out"].reshape(-1, 7), resp_fields(w[0]).reshape(-1, 7), "small path vs oracle")
    assert (res["small"]["st"] & 4).any() or (res["small"]["it"] >= 1).any()


def test_config3_mesh_field_with_rebuild(oracle):
    """config 3 (scaled down): cubes + icospheres, LBVH rebuilt for every frame's translated meshes."""
    with xp.Scene(0) as s:
        for frame in range(3):
            tris, lo, hi = scenes.mesh_field(4000, seed=2, frame=frame)
            resps = scenes.field_spheres(lo, hi, 1500, seed=3 + frame)
            s.set_triangles(tris)
            s.build()
            got = s.detect_batch(resps)
            check_detect_exact(got, oracle.detect_batch(resps, tris), f"config3 frame {frame}")
            assert got[1].sum() > 20


def test_config4_four_iterations(oracle):
    """config 4 (scaled down): 4-iteration sweep-and-slide; ITER_CAP where the reference would go on."""
    tris, h = scenes.sine_terrain(48, 48, seed=41)
    resps = scenes.terrain_spheres(h, 4000, seed=42)
    w = oracle.collision_response_batch(resps, tris, max_iter=4)
    with xp.Scene(0, tris, prune=True) as s:
        out, it, st, ln = s.collision_response_batch(resps, max_iter=4)
        assert s.stats()["iterations_run"] == 4
    assert np.array_equal(it, w[1]) and np.array_equal(st, w[2])
    assert (st & xp._lib.ST_ITER_CAP).any()
    assert_bits_equal(resp_fields(out), resp_fields(w[0]), "config4")


def test_config4_clipmap_and_voxels(oracle):
    """config 4 geometry (scaled-down batch): the game's clipmap terrain soup + greedy-meshed voxel chunks,
    4 sweep-and-slide iterations."""
    tris = scenes.config4_scene()
    resps = scenes.scene_spheres(tris, 160, seed=3)
    want = oracle.detect_batch(resps, tris)
    w = oracle.collision_response_batch(resps, tris, max_iter=4)
    with xp.Scene(0, tris) as s:
        got = s.detect_batch(resps)
        out, it, st, ln = s.collision_response_batch(resps, max_iter=4)
        gi, gj = s.broadphase_pairs(resps)
    check_detect_exact(got, want, "config4 scene")
    assert want[1].sum() > 100
    assert np.array_equal(it, w[1]) and np.array_equal(st, w[2])
    assert_bits_equal(resp_fields(out), resp_fields(w[0]), "config4 scene response")
    wi, wj = oracle.broadphase_pairs(resps, tris)
    assert np.array_equal(np.sort(gi.astype(np.uint64) << 32 | gj), np.sort(wi.astype(np.uint64) << 32 | wj))


# ---- full size (BASELINE config 2: 1M spheres x 1M triangles): size-independent properties ------
@pytest.fixture(scope="module")
def full_scene():
    tris, h = scenes.sine_terrain(708, 707, seed=1)
    resps = scenes.terrain_spheres(h, 1 << 20, seed=1)
    return tris, h, resps


def test_full_size_properties(full_scene, oracle):
    tris, h, resps = full_scene
    with xp.Scene(0, tris) as s:
        a = s.detect_batch(resps)
        tests_all = s.stats()["narrowphase_tests"]
        b = s.detect_batch(resps)                                   # idempotence
        s.set_options(method="lbvh_fused", prune=True)
        c = s.detect_batch(resps)                                   # pruning must not change any contact
        tests_pruned = s.stats()["narrowphase_tests"]
        s.set_options(method="lbvh_pairs", prune=True)             # distance-window passes of the two-stage sweep
        c2 = s.detect_batch(resps)
        tests_windowed = s.stats()["narrowphase_tests"]
        s.set_options(method="lbvh_pairs", prune=False, sort_spheres=False)
        d = s.detect_batch(resps)                                   # processing order must not matter
        u = s.detect_batch(resps, horizon="unit")
        s.set_options(sort_spheres=True)
        perm = np.random.default_rng(0).permutation(len(resps))
        e = s.detect_batch(resps[perm])                             # permutation equivariance
    for other, name in ((b, "rerun"), (d, "unsorted")):
        check_detect_exact(other, a, name)
    check_detect_exact(c, a, "pruned", tri_index=False)
    check_detect_exact(c2, a, "windowed", tri_index=False)
    assert tests_windowed < tests_all
    check_detect_exact(e, tuple(x[perm] for x in a), "permuted")
    col, hit, tri, st = a
    assert hit.mean() > 0.5
    assert tests_pruned < tests_all
    assert 8 * len(resps) < tests_all < 200 * len(resps)
    hh = hit.astype(bool)
    assert (col["time_to"][hh] >= 0).all()
    # a unit-horizon hit is also the infinite-horizon answer whenever t <= 1 ... and never earlier
    uh = u[1].astype(bool)
    assert (uh <= hh).all()
    assert (u[0]["time_to"][uh] >= col["time_to"][uh]).all()
    # the reported triangle really produces the reported time (single-pair call on the device data)
    idx = np.flatnonzero(hh)[:: max(1, hh.sum() // 64)][:64]
    sub = oracle.detect_batch(resps[idx], tris, threads=0)          # brute force: 64 x 1M tests on the CPU
    check_detect_exact(tuple(x[idx] for x in a), sub, "64-sphere sample vs brute-force oracle")


@pytest.fixture(scope="module")
def full_brute(full_scene):
    """The reference's literal algorithm (lib.rs:33-35: every triangle for every sphere) on the DEVICE for the
    whole benchmark workload: 2^20 x 1 001 112 = 1.05e12 evaluations of a11 (k_brute, itself bit-compared with
    the oracle at small size by test_detect_exact_bitwise[brute])."""
    tris, h, resps = full_scene
    with xp.Scene(0, tris, method="brute") as s:
        return s.detect_batch(resps)


@pytest.mark.parametrize("far_exact", [False, True])
@pytest.mark.parametrize("arith", ["exact"])
def test_full_size_every_sphere_against_brute_force(full_scene, full_brute, far_exact, arith):
    """Predicate P + the LBVH are conservative on the benchmark workload itself: the two-stage sweep agrees with
    the all-pairs sweep for ALL 2^20 spheres -- hit, triangle, every bit of the collision -- with and without the
    far-exact pass."""
    tris, h, resps = full_scene
    with xp.Scene(0, tris, far_exact=far_exact, arith=arith) as s:
        got = s.detect_batch(resps)
        pairs = s.stats()["narrowphase_tests"]
    check_detect_exact(got, full_brute, f"lbvh_pairs(far_exact={far_exact}) vs brute, all spheres")
    assert pairs < len(resps) * 200                                    # ... with ~1e-5 of the all-pairs work


def test_full_size_host_batch_above_pipeline_threshold(full_scene):
    """3 x 2^20 spheres through the host ABI: above 2^21 the call is pipelined in chunks by default"""
    tris, h, resps = full_scene
    big = np.concatenate([resps, resps[::-1], resps])
    with xp.Scene(0, tris) as s:
        a = s.detect_batch(resps)
        b = s.detect_batch(big)
    n = len(resps)
    check_detect_exact(tuple(x[:n] for x in b), a, "first third")
    check_detect_exact(tuple(x[n:2 * n][::-1] for x in b), a, "second third (reversed)")
    check_detect_exact(tuple(x[2 * n:] for x in b), a, "last third")


def test_full_size_response_sample(full_scene, oracle):
    tris, h, resps = full_scene
    n = 1 << 18
    with xp.Scene(0, tris, prune=True) as s:
        out, it, st, ln = s.collision_response_batch(resps[:n], max_iter=4)
    idx = np.arange(0, n, n // 24)[:24]
    w = oracle.collision_response_batch(resps[idx], tris, max_iter=4)
    assert np.array_equal(it[idx], w[1]) and np.array_equal(st[idx], w[2])
    assert_bits_equal(resp_fields(out[idx]), resp_fields(w[0]), "response sample")
    assert it.max() == 4


# ---- row f3: soups ingested from the reference's OBJ / glTF / .vox assets ----------------------
def ingested_scene():
    """The committed golden soups (tests/golden/make_ingest_golden.py) assembled into one scene in
    unit-sphere space: two greedy-meshed treehouse chunks, the ground plane, and the small props."""
    g = np.load(os.path.join(GOLDEN, "ingest_scenes.npz"))
    radius = np.float32(0.5)                                         # player sphere, main.rs:122
    parts = [g["vox_chunk_1_1_1"], g["vox_chunk_2_0_1"],
             g["obj_ground_plane_20x20"] + np.array([5, 0, 5], np.float32),
             g["obj_axis"] + np.array([4, 4, 4], np.float32), g["obj_arrow"] + np.array([6, 5, 3], np.float32),
             g["gltf_Cube_0"] + np.array([8, 1, 2], np.float32), g["gltf_Cube_1"] + np.array([2, 7, 6], np.float32)]
    return np.ascontiguousarray(np.concatenate(parts) / radius, dtype=np.float32)


@pytest.mark.parametrize("method", ["lbvh_pairs", "wide_fused", "q4_pairs"])
def test_ingested_assets_scene(oracle, method):
    tris = ingested_scene()
    assert len(tris) == 3016 + 1788 + 128 + 52 + 12 + 12 + 12
    rng = np.random.default_rng(77)
    lo, hi = tris.reshape(-1, 3).min(axis=0), tris.reshape(-1, 3).max(axis=0)
    n = 2500
    c = rng.uniform(lo - 1.5, hi + 1.5, (n, 3)).astype(np.float32)
    m = rng.normal(0, 1.2, (n, 3)).astype(np.float32)
    resps = xp.make_responses(c, m)
    want = oracle.detect_batch(resps, tris)
    w = oracle.collision_response_batch(resps, tris, max_iter=3)
    with xp.Scene(0, tris, method=method) as s:
        got = s.detect_batch(resps)
        out, it, st, ln = s.collision_response_batch(resps, max_iter=3)
        gi, gj = s.broadphase_pairs(resps)
    check_detect_exact(got, want, f"ingested scene {method}")
    assert want[1].sum() > 500
    assert np.array_equal(it, w[1]) and np.array_equal(st, w[2])
    assert_bits_equal(resp_fields(out), resp_fields(w[0]), f"ingested scene response {method}")
    wi, wj = oracle.broadphase_pairs(resps, tris)
    assert np.array_equal(np.sort(gi.astype(np.uint64) << 32 | gj), np.sort(wi.astype(np.uint64) << 32 | wj))


# ---- asynchronous stream of batches: xp_detect_batch_submit / _wait ----------------------------
def test_async_stream_of_batches(terrain, oracle):
    import torch
    tris, h, resps, want = terrain
    n = len(resps)
    batches = [resps, scenes.terrain_spheres(h, n, seed=91), scenes.terrain_spheres(h, n, seed=92)]
    wants = [want, oracle.detect_batch(batches[1], tris), oracle.detect_batch(batches[2], tris)]

    def pinned(dtype):
        return torch.from_numpy(np.zeros(n, dtype).view(np.uint8)).pin_memory()
    bufs = [{"in": pinned(xp.RESPONSE_DTYPE), "col": pinned(xp.COLLISION_DTYPE), "hit": pinned(np.uint8),
             "tri": pinned(np.uint32), "st": pinned(np.uint32)} for _ in range(2)]
    with xp.Scene(0, tris) as s:
        tickets = {}
        results = {}

        def collect(k):
            s.detect_batch_wait(tickets.pop(k))
            b = bufs[k % 2]
            assert s.stats()["narrowphase_tests"] > 0
            results[k] = (b["col"].numpy().view(xp.COLLISION_DTYPE).copy(), b["hit"].numpy().copy(),
                          b["tri"].numpy().view(np.uint32).copy(), b["st"].numpy().view(np.uint32).copy())
        for rep in range(2):
            for k, r in enumerate(batches):
                k += 3 * rep
                if k >= 2:
                    collect(k - 2)                                   # frees the slot and the buffers of batch k - 2
                b = bufs[k % 2]
                b["in"].numpy()[:] = np.ascontiguousarray(r).view(np.uint8)
                tickets[k] = s.detect_batch_submit(b["in"].data_ptr(), n, b["col"].data_ptr(), b["hit"].data_ptr(),
                                                   b["tri"].data_ptr(), b["st"].data_ptr())
                if k == 1:                                           # both slots in flight: a third submit is refused
                    with pytest.raises(xp.XpError):
                        s.detect_batch_submit(b["in"].data_ptr(), n)
        collect(4)
        collect(5)
        with pytest.raises(xp.XpError):
            s.detect_batch_wait(0)                                   # not in flight
        # a synchronous call still works after the stream
        check_detect_exact(s.detect_batch(resps), want, "sync after async")
    for k in range(6):
        check_detect_exact(results[k], wants[k % 3], f"async batch {k}")


def test_async_batch_overflow_is_redone(terrain):
    import torch
    tris, h, resps, want = terrain
    n = len(resps)
    h_in = torch.from_numpy(np.ascontiguousarray(resps).view(np.uint8).copy()).pin_memory()
    h_col = torch.zeros(n * 32, dtype=torch.uint8).pin_memory()
    h_hit = torch.zeros(n, dtype=torch.uint8).pin_memory()
    h_tri = torch.zeros(n * 4, dtype=torch.uint8).pin_memory()
    h_st = torch.zeros(n * 4, dtype=torch.uint8).pin_memory()
    with xp.Scene(0, tris) as s:
        s.set_options(max_pairs=4096)                                # far too small: candidates are dropped
        t = s.detect_batch_submit(h_in.data_ptr(), n, h_col.data_ptr(), h_hit.data_ptr(), h_tri.data_ptr(), h_st.data_ptr())
        s.detect_batch_wait(t)
    got = (h_col.numpy().view(xp.COLLISION_DTYPE), h_hit.numpy(), h_tri.numpy().view(np.uint32), h_st.numpy().view(np.uint32))
    check_detect_exact(got, want, "async overflow redo")


# ---- a loop of steps through the asynchronous fused exchange (one process: the only peer is the local buffer)
def test_async_exchange_loop_single_process(terrain, oracle):
    import torch
    from xp_physics_cuda.distributed import ContactExchange, contacts_as_records
    tris, h, resps, want = terrain
    n = len(resps)
    batches = [resps, scenes.terrain_spheres(h, n, seed=93), scenes.terrain_spheres(h, n - 7, seed=94)]
    wants = [want, oracle.detect_batch(batches[1], tris), oracle.detect_batch(batches[2], tris)]
    dev = torch.device("cuda", 0)
    with xp.Scene(0, tris) as s:
        s.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        ex = ContactExchange(s, n, dev)
        assert ex.ok and ex.solo and ex.mode == "p2p"
        d_in = [torch.from_numpy(np.ascontiguousarray(b).view(np.float32).reshape(-1, 7)).to(dev) for b in batches]
        pending, got = [], {}

        def finish():
            k, t = pending.pop(0)
            buf = ex.wait(t)
            assert s.stats()["narrowphase_tests"] > 0
            got[k] = contacts_as_records(buf[0, :len(batches[k % 3])])
        for k in range(7):
            if len(pending) == 2:
                finish()
            pending.append((k, ex.sweep_async(d_in[k % 3].data_ptr(), len(batches[k % 3]))))
        while pending:
            finish()
        ex.close()
    for k in range(7):
        rec, w = got[k], wants[k % 3]
        hit = (rec["flags"] & 1).astype(np.uint8)
        assert np.array_equal(hit, w[1]), f"step {k}"
        hm = hit.astype(bool)
        assert np.array_equal(rec["tri_index"][hm], w[2][hm])
        assert_bits_equal(rec["time_to"][hm], w[0]["time_to"][hm], f"async exchange step {k}")
        assert_bits_equal(rec["position"][hm], w[0]["position"][hm], f"async exchange step {k}")
        assert_bits_equal(rec["intersection"][hm], w[0]["intersection"][hm], f"async exchange step {k}")


def test_compact_contact_records(terrain, oracle):
    """XP_OPT_CONTACT_RECORD = 1: xp_contact16 = (centre at the contact, triangle), 16 B instead of 40 -- through the
    contacts-only entry point, the synchronous exchange and the loop of steps; misses are (0, 0, 0, 0xFFFFFFFF)."""
    import torch
    from xp_physics_cuda import CONTACT16_DTYPE
    from xp_physics_cuda.distributed import ContactExchange, contacts_as_records
    tris, h, resps, want = terrain
    n = len(resps)
    col, hit, tri = want[0], want[1].astype(bool), want[2]
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(np.ascontiguousarray(resps).view(np.float32).reshape(-1, 7)).to(dev)

    def check(rec, what):
        assert rec.dtype == CONTACT16_DTYPE and len(rec) == n, what
        assert np.array_equal(rec["tri_index"], np.where(hit, tri, 0xFFFFFFFF).astype(np.uint32)), what
        assert_bits_equal(rec["position"][hit], col["position"][hit], what)
        assert not rec["position"][~hit].any(), what
    with xp.Scene(0, tris) as s:
        s.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        s.set_options(contact_record="compact")
        out = torch.full((n + 1, 4), 7.0, dtype=torch.float32, device=dev)
        s.detect_contacts_dev(d_in.data_ptr(), n, out.data_ptr())
        torch.cuda.synchronize()
        assert (out[n] == 7.0).all()                                       # nothing written past n records of 16 B
        check(contacts_as_records(out[:n]), "xp_detect_contacts_dev")
        s.detect_contacts_dev(d_in.data_ptr(), n, out.data_ptr() + 4)      # 4-byte aligned only: the per-sphere writer
        torch.cuda.synchronize()
        check(contacts_as_records(out.reshape(-1)[1:1 + 4 * n].reshape(n, 4).clone()), "xp_detect_contacts_dev, unaligned")
        ex = ContactExchange(s, n, dev, record="compact")
        assert ex.ok and ex.view.shape[-1] == 4
        ex.sweep(d_in.data_ptr(), n)
        torch.cuda.synchronize()
        check(contacts_as_records(ex.view[0, :n]), "synchronous exchange")
        t = ex.sweep_async(d_in.data_ptr(), n)
        check(contacts_as_records(ex.wait(t)[0, :n]), "loop of steps")
        ex.close()
        s.set_options(contact_record="full")                               # and back: the 40 B records again
        full = torch.zeros((n, 10), dtype=torch.float32, device=dev)
        s.detect_contacts_dev(d_in.data_ptr(), n, full.data_ptr())
        torch.cuda.synchronize()
        rec = contacts_as_records(full)
        assert np.array_equal((rec["flags"] & 1).astype(bool), hit)
        assert_bits_equal(rec["position"][hit], col["position"][hit], "full records after compact")


# ---- ill-conditioned triangles: the reference's face test reports hits the geometric broadphase never offers
SLIVER = [[-0.7752354741096497, -0.6470972299575806, 5.695098400115967],
          [-0.21221216022968292, -0.4316372871398926, 5.311792373657227],
          [0.3508111536502838, -0.21617737412452698, 4.928485870361328]]      # v1 = fl((v0 + v2) / 2): nearly collinear
SLIVER_SPHERES = [([7.1984725, 3.6347277, -6.416019], [1.4218696, -0.8297933, -0.7164174]),
                  ([7.556139, 4.7300577, -2.830135], [0.78011984, -2.0790448, 0.52724224]),
                  ([5.215561, 3.7503188, 2.4745898], [0.48196483, -1.0705745, 0.48744422])]


@pytest.mark.parametrize("method,prune", METHODS)
def test_sliver_triangle_far_face_hits(oracle, method, prune):
    """Found by tools/stress_parity.py: with a rounding-dominated normal the reference 'hits' this sliver from far
    outside its inflated box.  The library lists such triangles at build time and sweeps them exhaustively."""
    rng = np.random.default_rng(3)
    others = (rng.uniform(-6, 6, (200, 1, 3)) + rng.normal(0, 1.0, (200, 3, 3))).astype(np.float32)
    tris = np.concatenate([others[:3], np.asarray([SLIVER], np.float32), others[3:]])
    extra = xp.make_responses(rng.uniform(-8, 8, (1500, 3)).astype(np.float32), rng.normal(0, 1, (1500, 3)).astype(np.float32))
    resps = np.concatenate([xp.make_responses([c for c, m in SLIVER_SPHERES], [m for c, m in SLIVER_SPHERES]), extra])
    want = oracle.detect_batch(resps, tris)
    assert want[1][:3].all() and (want[2][:2] == 3).all()             # the reference hits the sliver (index 3)
    assert not oracle.broadphase_pairs(resps[:3], tris[3:4])[0].size  # ... although P rejects those pairs
    w = oracle.collision_response_batch(resps, tris, max_iter=3)
    with xp.Scene(0, tris, method=method, prune=prune) as s:
        got = s.detect_batch(resps)
        out, it, st, ln = s.collision_response_batch(resps, max_iter=3)
    check_detect_exact(got, want, f"sliver {method}")
    assert np.array_equal(it, w[1]) and np.array_equal(st, w[2])
    assert_bits_equal(resp_fields(out), resp_fields(w[0]), f"sliver response {method}")


# ---- far-range grazing: the reference's quadratics lose ~3e-7 D^2 of miss distance ----------
def far_grazing_case(seed=4):
    rng = np.random.default_rng(seed)
    tris = (rng.uniform(-4, 4, (8, 1, 3)) + rng.normal(0, 0.05, (8, 3, 3))).astype(np.float32)
    cs, ms = [], []
    for D in (30.0, 100.0, 300.0, 1000.0):
        ns = 1500
        k = rng.integers(0, 8, ns)
        axis, sgn = rng.integers(0, 3, ns), rng.choice([-1.0, 1.0], ns)
        d = 1.0 + rng.uniform(-0.2, 1.5, ns) * 6e-8 * D * D / 2 + rng.uniform(0, 2e-4, ns)
        off = np.zeros((ns, 3))
        off[np.arange(ns), axis] = sgn * d                       # closest approach: next to vertex 0 of triangle k, along an axis
        dirs = rng.normal(0, 1, (ns, 3))
        dirs[np.arange(ns), axis] = 0
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        cs.append(tris[k, 0].astype(np.float64) + off - dirs * D)
        ms.append(dirs * rng.uniform(0.05, 2.0, (ns, 1)))
    return tris, xp.make_responses(np.concatenate(cs).astype(np.float32), np.concatenate(ms).astype(np.float32))


def test_far_exact_reproduces_the_references_far_grazing_noise(oracle):
    tris, resps = far_grazing_case()
    want = oracle.detect_batch(resps, tris)
    w = oracle.collision_response_batch(resps, tris, max_iter=2)
    with xp.Scene(0, tris) as s:
        plain = s.detect_batch(resps)
        s.set_options(far_exact=True)
        got = s.detect_batch(resps)
        tests_far = s.stats()["narrowphase_tests"]
        out, it, st, ln = s.collision_response_batch(resps, max_iter=2)
    # default mode: the documented gap -- the reference "hits" triangles whose inflated box the path never enters
    n_gap = int((plain[1] != want[1]).sum())
    print("far-grazing rows where the default mode differs from the reference:", n_gap, "of", len(resps))
    assert n_gap > 0
    assert not (plain[1].astype(bool) & ~want[1].astype(bool)).any()        # never the other way round
    # XP_OPT_FAR_EXACT closes it, bit for bit
    check_detect_exact(got, want, "far exact")
    assert tests_far > 0
    assert np.array_equal(it, w[1]) and np.array_equal(st, w[2])
    assert_bits_equal(resp_fields(out), resp_fields(w[0]), "far exact response")
    # and changes nothing where there is no far noise
    tris2, h = scenes.sine_terrain(40, 40, seed=31)
    resps2 = scenes.terrain_spheres(h, 3000, seed=32)
    with xp.Scene(0, tris2) as s:
        a = s.detect_batch(resps2)
        s.set_options(far_exact=True)
        b = s.detect_batch(resps2)
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))


def test_leaf_runs_change_nothing_but_the_visits(terrain, oracle):
    """XP_OPT_LEAF_RUNS: two-leaf children of the LBVH are emitted unvisited and stage 2 applies P per triangle -- same
    tested set (|P|), same contacts bit for bit, fewer node visits; with and without the option, both horizons, through
    the response loop, and in a pruned sweep (a run belongs to the windowed pass that holds its union box's entry time, so a
    triangle may be tested a pass earlier than without runs: the tested set may differ slightly, the contacts may not)."""
    tris, h, resps, want = terrain
    n_p = {hz: len(oracle.broadphase_pairs(resps, tris, np.inf if hz == "inf" else 1.0)[0]) for hz in ("inf", "unit")}
    for prune in (False, True):
        res = {}
        for runs in (False, True):
            with xp.Scene(0, tris, prune=prune) as s:
                s.set_options(leaf_runs=runs)
                got = s.detect_batch(resps)
                st_ = s.stats()
                unit = s.detect_batch(resps, horizon="unit")
                st_u = s.stats()
                out = s.collision_response_batch(resps[:600], max_iter=3)
            res[runs] = (got, st_, unit, st_u, out)
            check_detect_exact(got, want, f"leaf runs {runs} prune {prune}", tri_index=not prune)
            if not prune:
                assert st_["narrowphase_tests"] == n_p["inf"] and st_u["narrowphase_tests"] == n_p["unit"]
                assert (st_["broadphase_pairs"] == n_p["inf"]) == (not runs) and st_["broadphase_pairs"] >= n_p["inf"]
        (a, sa, ua, sua, oa), (b, sb, ub, sub_, ob) = res[False], res[True]
        check_detect_exact(b, a, f"runs on vs off (prune {prune})", tri_index=not prune)
        check_detect_exact(ub, ua, f"runs on vs off, unit horizon (prune {prune})", tri_index=not prune)
        if prune:
            assert abs(sb["narrowphase_tests"] - sa["narrowphase_tests"]) < 0.1 * sa["narrowphase_tests"]
        else:
            assert sa["narrowphase_tests"] == sb["narrowphase_tests"] and sua["narrowphase_tests"] == sub_["narrowphase_tests"]
        assert sb["node_visits"] < 0.9 * sa["node_visits"]
        assert sb["broadphase_pairs"] < 1.25 * sa["broadphase_pairs"]
        print(f"leaf runs (prune {prune}): {sa['node_visits']} -> {sb['node_visits']} node visits, {sb['broadphase_pairs']} pairs emitted, "
              f"{sa['narrowphase_tests']} / {sb['narrowphase_tests']} tests (|P| = {n_p['inf']})")
        for x, y in zip(oa, ob):
            assert np.array_equal(np.ascontiguousarray(x).view(np.uint8), np.ascontiguousarray(y).view(np.uint8))
