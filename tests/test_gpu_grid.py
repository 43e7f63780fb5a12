"""Heightmap scenes swept by walking the height grid (k_broadphase_grid, csrc/xp_gridwalk.cuh) instead of the LBVH:
the candidate set is predicate P bit for bit, every contact is the oracle's bit for bit, and rays the walk cannot
take (non-finite) fall through to the LBVH walk inside the same sweep."""
import numpy as np
import pytest

import xp_physics_cuda as xp
from tests.conftest import assert_bits_equal
from tests.test_gpu_parity import check_detect_exact, resp_fields
from xp_physics_cuda import scenes

pytestmark = pytest.mark.gpu

CASES = {"config2-like": (60, 57, 0.0, 0.0, 1.0, 3), "offset-fine": (33, 29, -5.5, 2.25, 0.37, 4),
         "coarse": (9, 12, 100.0, -50.0, 7.5, 5), "far-origin": (24, 24, 4000.0, -3000.0, 1.0, 6)}


def make_case(case, n=3000):
    nx, nz, x0, z0, sp, seed = CASES[case]
    h = scenes.sine_terrain_heights(nx, nz, seed=seed)
    tris = scenes.heightmap_from_heights(h, x0=x0, z0=z0, spacing=sp)
    rng = np.random.default_rng(seed)
    resps = scenes.terrain_spheres(h, n, seed=seed + 10)
    resps["c"][:, 0] = resps["c"][:, 0] * sp + x0
    resps["c"][:, 2] = resps["c"][:, 2] * sp + z0
    k = np.arange(n)
    resps["movement"][k % 11 == 1, 0] = 0.0
    resps["movement"][k % 11 == 2, 2] = 0.0
    resps["movement"][k % 11 == 3, 0::2] = 0.0                        # vertical
    resps["movement"][k % 11 == 4] = 0.0                              # zero movement
    resps["movement"][k % 11 == 5, 1] *= -1.0                         # upwards
    out = k % 11 == 6
    resps["c"][out] += rng.uniform(-3, 3, (out.sum(), 3)).astype(np.float32) * np.float32([nx * sp, 1, nz * sp])
    on_line = k % 11 == 7
    resps["c"][on_line, 0] = (x0 + sp * rng.integers(0, nx + 1, on_line.sum())).astype(np.float32)
    return h, tris, resps, (x0, z0, sp)


@pytest.mark.parametrize("case", list(CASES))
def test_grid_walk_candidates_and_contacts(oracle, case):
    h, tris, resps, (x0, z0, sp) = make_case(case)
    with xp.Scene(0) as s:
        s.set_heightmap(h, x0=x0, z0=z0, spacing=sp)
        s.build()
        for horizon in ("inf", "unit"):
            gi, gj = s.broadphase_pairs(resps, horizon=horizon)
            wi, wj = oracle.broadphase_pairs(resps, tris, horizon=np.inf if horizon == "inf" else 1.0)
            assert np.array_equal(np.sort(gi.astype(np.uint64) << 32 | gj), np.sort(wi.astype(np.uint64) << 32 | wj)), (case, horizon)
        cells = s.stats()["node_visits"]
        got = s.detect_batch(resps)
        tests = s.stats()["narrowphase_tests"]
        out = s.collision_response_batch(resps, max_iter=3)
        s.set_options(grid_walk=False)                                 # the LBVH walk over the same scene
        lb = s.detect_batch(resps)
        tests_lb = s.stats()["narrowphase_tests"]
    want = oracle.detect_batch(resps, tris)
    check_detect_exact(got, want, f"grid walk {case}")
    check_detect_exact(lb, want, f"lbvh walk {case}")
    assert tests == tests_lb == len(oracle.broadphase_pairs(resps, tris)[0])
    w = oracle.collision_response_batch(resps, tris, max_iter=3)
    assert np.array_equal(out[1], w[1]) and np.array_equal(out[2], w[2])
    assert_bits_equal(resp_fields(out[0]), resp_fields(w[0]), "response loop on the grid walk")
    print(f"{case}: {tests} candidates, {cells} cells walked")


def test_grid_walk_leaves_non_finite_rays_to_the_lbvh(oracle):
    h, tris, resps, (x0, z0, sp) = make_case("config2-like", n=2000)
    bad = np.arange(0, len(resps), 7)
    vals = np.float32([np.nan, np.inf, -np.inf])
    for j, i in enumerate(bad):
        f = "c" if j % 2 else "movement"
        resps[f][i, j % 3] = vals[(j // 3) % 3]
    with xp.Scene(0) as s:
        s.set_heightmap(h, x0=x0, z0=z0, spacing=sp)
        s.build()
        got = s.detect_batch(resps)
        gi, gj = s.broadphase_pairs(resps)
    check_detect_exact(got, oracle.detect_batch(resps, tris), "non-finite rays among regular ones")
    wi, wj = oracle.broadphase_pairs(resps, tris)
    assert np.array_equal(np.sort(gi.astype(np.uint64) << 32 | gj), np.sort(wi.astype(np.uint64) << 32 | wj))


def test_grid_walk_full_size_against_the_lbvh(oracle):
    """config 2 at full size through xp_scene_set_heightmap: the grid walk and the LBVH walk produce the same
    number of candidates and identical contacts for all 2^20 spheres (the LBVH path is itself compared with the
    all-pairs sweep for every sphere in test_gpu_parity.py)."""
    h = scenes.sine_terrain_heights(708, 707, seed=1)
    resps = scenes.terrain_spheres(h, 1 << 20, seed=1)
    with xp.Scene(0) as s:
        s.set_heightmap(h)
        s.build()
        a = s.detect_batch(resps)
        ta, cells = s.stats()["narrowphase_tests"], s.stats()["node_visits"]
        fa = s.detect_batch(resps[: 1 << 18], horizon="unit")
        s.set_options(grid_walk=False)
        b = s.detect_batch(resps)
        tb, nodes = s.stats()["narrowphase_tests"], s.stats()["node_visits"]
        fb = s.detect_batch(resps[: 1 << 18], horizon="unit")
    check_detect_exact(a, b, "grid walk vs LBVH walk, all spheres")
    check_detect_exact(fa, fb, "unit horizon")
    assert ta == tb
    print(f"config 2: {ta} candidates; grid walk {cells} cells ({2 * cells / ta:.2f} boxes per candidate), LBVH {nodes} node visits")
