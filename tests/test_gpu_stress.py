"""A slice of tools/stress_parity.py in the GPU tier: mid-size adversarial scenes (random soups with degenerate
triangles; needles, strips, huge and tiny triangles; terrain far from the origin) against the oracle, where the
sphere ordering, the persistent broadphase with work stealing and the exhaustive list all engage."""
import os
import sys

import numpy as np
import pytest

import xp_physics_cuda as xp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import stress_parity as sp  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,seed,max_iter", [(2, 7, 3), (2, 70, 0), (4, 8, 3), (4, 80, 0), (5, 9, 3), (1, 10, 2)])
def test_stress_slice(oracle, kind, seed, max_iter):
    rng = np.random.default_rng(seed)
    tris, resps = sp.make_case(rng, kind)
    want = oracle.detect_batch(resps, tris)
    wres = oracle.collision_response_batch(resps, tris, max_iter=max_iter)
    for prune in (False, True):
        with xp.Scene(0, tris, prune=prune) as s:
            got = s.detect_batch(resps)
            out, it, st, ln = s.collision_response_batch(resps, max_iter=max_iter)
        h = want[1].astype(bool)
        assert np.array_equal(got[1], want[1])
        assert np.array_equal(sp.bits(sp.fields(got[0])[h]), sp.bits(sp.fields(want[0])[h]))
        assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
        assert np.array_equal(it, wres[1]) and np.array_equal(st, wres[2])
        assert np.array_equal(sp.bits(sp.rfields(out)), sp.bits(sp.rfields(wres[0])))
    assert h.sum() > 100
