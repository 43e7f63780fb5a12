"""Structure of the device-built LBVH (new code: the reference has no broadphase)."""
import numpy as np
import pytest

import xp_physics_cuda as xp
from xp_physics_cuda import scenes

pytestmark = pytest.mark.gpu


def _expand(v):
    v = (v * 0x00010001) & 0xFF0000FF
    v = (v * 0x00000101) & 0x0F00F00F
    v = (v * 0x00000011) & 0xC30C30C3
    v = (v * 0x00000005) & 0x49249249
    return v


def _check_tree(tris, oracle):
    n = len(tris)
    with xp.Scene(0, tris) as s:
        sorted_tri, morton, child, nlo, nhi = s.debug_tree()
        run = s.debug_leaf_runs()
    assert np.array_equal(np.sort(sorted_tri), np.arange(n))                 # a permutation
    assert (np.diff(morton.astype(np.int64)) >= 0).all()                     # sorted
    assert (morton < (1 << 30)).all()
    # stable: equal codes keep upload order
    same = morton[1:] == morton[:-1]
    assert (sorted_tri[1:][same] > sorted_tri[:-1][same]).all()
    # codes are the 30-bit Morton codes of the inflated-box centres
    lo, hi = oracle.triangle_aabbs(tris)
    c = 0.5 * (lo + hi)
    mn, mx = c.min(axis=0), c.max(axis=0)
    ext = np.float32((mx - mn).max())                                        # isotropic: cubic cells
    scale = np.float32(1024.0) / ext if ext > 0 else np.float32(0)
    q = np.clip((c - mn) * scale, 0, 1023).astype(np.uint32)
    code = (_expand(q[:, 0].astype(np.uint64)) << 2) | (_expand(q[:, 1].astype(np.uint64)) << 1) | _expand(q[:, 2].astype(np.uint64))
    assert (code[sorted_tri] == morton).mean() > 0.99       # float rounding at cell borders may differ
    if n < 2:
        return
    # every leaf and every internal node except the root has exactly one parent; boxes nest
    seen_leaf = np.zeros(n, np.int32)
    seen_node = np.zeros(n - 1, np.int32)
    llo, lhi = lo[sorted_tri], hi[sorted_tri]
    for i in range(n - 1):
        for c_ in child[i]:
            if c_ < 0:
                seen_leaf[~c_] += 1
                blo, bhi = llo[~c_], lhi[~c_]
            else:
                seen_node[c_] += 1
                blo, bhi = nlo[c_], nhi[c_]
            assert (nlo[i] <= blo).all() and (nhi[i] >= bhi).all()
    assert (seen_leaf == 1).all() and seen_node[0] == 0 and (seen_node[1:] == 1).all()
    # leaf runs (XP_OPT_LEAF_RUNS): marked on exactly the children that are internal nodes over two leaves, with the first leaf
    kid = child.astype(np.int64)
    internal = kid >= 0
    grand = child[np.where(internal, kid, 0)]                               # (n-1, 2, 2): the children's own links
    two_leaves = internal & (grand < 0).all(axis=2)
    assert np.array_equal(run > 0, two_leaves)
    assert np.array_equal(run[two_leaves] - 1, ~grand[two_leaves][:, 0])
    assert np.array_equal(~grand[two_leaves][:, 1], run[two_leaves])
    assert two_leaves.sum() > 0 or n < 3
    # leaf ranges: an in-order walk visits leaves 0..n-1 in order
    order = []
    stack = [0]
    while stack:
        x = stack.pop()
        if x < 0:
            order.append(~x)
        else:
            stack.append(child[x][1]); stack.append(child[x][0])
    assert order == list(range(n))
    assert np.array_equal(nlo[0], llo.min(axis=0)) and np.array_equal(nhi[0], lhi.max(axis=0))


def test_tree_invariants_terrain(oracle):
    tris, _ = scenes.sine_terrain(30, 31, seed=4)
    _check_tree(tris, oracle)


def test_tree_invariants_duplicates_and_small(oracle):
    tris, _ = scenes.sine_terrain(6, 6, seed=4)
    _check_tree(np.concatenate([tris, tris, tris[:9]]), oracle)              # many identical Morton codes
    for n in (1, 2, 3, 4, 7):
        _check_tree(tris[:n], oracle)
    _check_tree(np.repeat(tris[:1], 37, axis=0), oracle)                      # all codes equal
    # around the 256-leaf segments of the bottom-up tree pass, with and without runs of equal codes
    field, _, _ = scenes.mesh_field(1100, seed=11)
    for n in (255, 256, 257, 511, 512, 513, 1025):
        _check_tree(field[:n], oracle)
    _check_tree(np.concatenate([np.repeat(field[:3], 200, axis=0), field[:300]]), oracle)


def test_tree_large_sort(oracle):
    """radix sort across many tiles (T > 4096 * several) incl. a ragged last tile."""
    tris, lo, hi = scenes.mesh_field(70001, seed=5)
    with xp.Scene(0, tris) as s:
        sorted_tri, morton, child, nlo, nhi = s.debug_tree()
    assert np.array_equal(np.sort(sorted_tri), np.arange(len(tris)))
    assert (np.diff(morton.astype(np.int64)) >= 0).all()
    same = morton[1:] == morton[:-1]
    assert (sorted_tri[1:][same] > sorted_tri[:-1][same]).all()


def test_tree_work_list_overflow_path(oracle):
    """The bottom-up tree pass hands subtrees a 256-leaf segment cannot attach alone to a work list; when the list is
    full the lane climbs through the global protocol in place.  Forced here with a list of 4 entries (XP_TREE_WORK_CAP,
    read once per process, hence the subprocess): same tree, bit for bit."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, xp_physics_cuda as xp\n"
        "from xp_physics_cuda import scenes\n"
        "tris, lo, hi = scenes.mesh_field(30011, seed=7)\n"
        "with xp.Scene(0, tris) as s:\n"
        "    t = s.debug_tree()\n"
        "np.savez(r'%s', *t)\n")
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        outs = []
        for cap in (None, "3"):
            env = dict(os.environ)
            env["PYTHONPATH"] = os.pathsep.join([os.path.dirname(os.path.dirname(xp.__file__)), env.get("PYTHONPATH", "")])
            env.pop("XP_TREE_WORK_CAP", None)
            if cap is not None:
                env["XP_TREE_WORK_CAP"] = cap
            path = os.path.join(d, f"tree_{cap}.npz")
            subprocess.run([sys.executable, "-c", code % path], check=True, env=env, timeout=300)
            z = np.load(path)
            outs.append([z[k] for k in z.files])
    for a, b in zip(*outs):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    tris, _, _ = scenes.mesh_field(30011, seed=7)
    _check_tree(tris, oracle)


def test_device_heightmap_emission_equals_host_soup(oracle):
    """xp_scene_set_heightmap (row f2) builds exactly the scene xp_scene_set_triangles builds from the
    host-generated soup: same Morton order, same tree, same contacts."""
    h = scenes.sine_terrain_heights(37, 23, seed=9)
    tris = scenes.heightmap_from_heights(h, x0=-5.5, z0=2.25, spacing=0.75)
    resps = scenes.terrain_spheres(h, 500, seed=4)
    resps["c"][:, 0] = resps["c"][:, 0] * 0.75 - 5.5
    resps["c"][:, 2] = resps["c"][:, 2] * 0.75 + 2.25
    with xp.Scene(0, tris) as a, xp.Scene(0) as b:
        b.set_heightmap(h, x0=-5.5, z0=2.25, spacing=0.75)
        b.build()
        assert b.triangle_count == len(tris)
        ta, tb = a.debug_tree(), b.debug_tree()
        for x, y in zip(ta, tb):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
        ra, rb = a.detect_batch(resps), b.detect_batch(resps)
        for x, y in zip(ra, rb):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
        assert ra[1].sum() > 100


def test_device_clipmap_emission_equals_host_soup(oracle):
    """xp_scene_set_clipmap (row f2, clipmap half): the device emits, from the toroidal height store, exactly the
    soup the host mirror emits -- also after an incremental update around a moved centre."""
    from xp_physics_cuda import clipmap as cm
    c = cm.Clipmap(5)
    for centre in ((37.0, -21.0), (44.5, -3.0), (-300.0, 910.0)):
        c.update_heightmap(centre, scenes.sine_height)
        tris = c.triangles()
        assert np.array_equal(tris.view(np.uint32),
                              scenes.clipmap_triangles(centre, scenes.sine_height).view(np.uint32))
        resps = scenes.scene_spheres(tris, 600, seed=3, extent=40.0)
        with xp.Scene(0, tris) as a, xp.Scene(0) as b:
            b.set_clipmap(c)
            b.build()
            assert b.triangle_count == len(tris) == 127016
            for x, y in zip(a.debug_tree(), b.debug_tree()):
                assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
            ra, rb = a.detect_batch(resps), b.detect_batch(resps)
            for x, y in zip(ra, rb):
                assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
            assert ra[1].sum() > 100
    want = oracle.detect_batch(resps[:150], tris)
    assert np.array_equal(rb[1][:150], want[1])
