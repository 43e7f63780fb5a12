"""The exact mode's building blocks on the device (csrc/xp_debug.cu): xp_detect_x is bit-identical to the
reference only if its division / square-root sequences are correctly rounded over the whole guarded operand
range and its operand-side decisions equal the reference's quotient-side ones.  Checked against the compiler's
IEEE intrinsics: every square-root input in [2^-100, 2^100] (1.7e9 bit patterns), 2^33 hashed quotient pairs,
every denominator mantissa x 48 numerators x 3 exponents (1.2e9), and 2^33 hashed + structured operand sets for
each decision."""
import pytest

import xp_physics_cuda as xp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 20261020])
def test_handrolled_division_and_square_root_are_ieee(seed):
    got = xp.debug_check_arith(0, n_random=1 << 33, seed=seed)
    print(got)
    assert all(v == 0 for v in got.values()), got


@pytest.mark.parametrize("arith", ["exact", "exact_literal"])
def test_both_exact_kernels_agree_with_the_oracle(oracle, arith):
    """the round-1 kernel (IEEE intrinsics) stays selectable as the A/B baseline; both are the oracle bit for bit"""
    import numpy as np
    from tests.test_gpu_parity import check_detect_exact
    from xp_physics_cuda import scenes
    tris, h = scenes.sine_terrain(48, 48, seed=41)
    resps = scenes.terrain_spheres(h, 4000, seed=42)
    resps["movement"][::7] *= np.float32(1e-3)            # slow movers: small discriminants
    resps["movement"][5::11, 0] = 0.0                     # axis-aligned components
    with xp.Scene(0, tris, arith=arith) as s:
        got = s.detect_batch(resps)
    check_detect_exact(got, oracle.detect_batch(resps, tris), arith)
