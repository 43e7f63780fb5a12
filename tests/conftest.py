import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "xp-game-engine_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box via gpurun)")


def bits(a):
    """float32 array -> uint32 bit patterns with every NaN canonicalised (x86 and GPU NaN payloads differ)."""
    a = np.ascontiguousarray(a, np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


def assert_bits_equal(a, b, what=""):
    ba, bb = bits(a), bits(b)
    if not np.array_equal(ba, bb):
        bad = np.argwhere(ba != bb)
        raise AssertionError(f"{what}: {len(bad)} of {ba.size} float32 values differ bitwise; first at {bad[0]}: "
                             f"{np.asarray(a).reshape(ba.shape)[tuple(bad[0])]!r} vs "
                             f"{np.asarray(b).reshape(bb.shape)[tuple(bad[0])]!r}")


@pytest.fixture(scope="session")
def oracle():
    from oracle import xp_oracle
    xp_oracle.lib()
    return xp_oracle
