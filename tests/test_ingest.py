"""Mesh ingest: OBJ / glTF / .vox loaders and the greedy voxel mesher.

The loaders are checked three ways: (1) hand-written files whose triangles are known by construction,
(2) the reference's own pinned facts (world.rs `offset_test`, loader.rs `load_gltf_test`, face counts
of its assets), (3) tests/golden/ingest_scenes.npz -- soups this loader produced from the reference's
assets (make_ingest_golden.py); where /root/reference is mounted the loaders must reproduce them.
"""
import base64
import json
import os
import struct

import numpy as np
import pytest

from xp_physics_cuda import ingest, scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RES = "/root/reference/res/"
have_reference = pytest.mark.skipif(not os.path.isdir(RES), reason="/root/reference is not mounted here")


def golden():
    return np.load(os.path.join(GOLDEN, "ingest_scenes.npz"))


# ------------------------------------------------------------------------------------------ OBJ
OBJ_TEXT = """
# a unit quad, a pentagon written with relative indices, and a second object
mtllib demo.mtl
o quad
v 0 0 0
v 1 0 0
v 1 0 1
v 0 0 1
usemtl red
f 1/1/1 2/2/1 3/3/1 4/4/1
o fan
v 0 1 0
v 2 1 0
v 3 1 1
v 2 1 2
v 0 1 2
usemtl nothing
f -5 -4 -3 -2 -1
g tri
f 1//1 3//1 5//1
"""
MTL_TEXT = "newmtl red\nKd 0.8 0.1 0.05\n"


def test_obj_fan_triangulation_models_and_materials():
    sc = ingest.parse_obj(OBJ_TEXT, lambda name: MTL_TEXT)
    assert [m.name for m in sc.models] == ["quad", "fan", "tri"]
    assert [len(m.positions) for m in sc.models] == [2, 3, 1]
    t = sc.triangles()
    assert t.dtype == np.float32 and t.shape == (6, 3, 3)
    q = [[0, 0, 0], [1, 0, 0], [1, 0, 1], [0, 0, 1]]
    np.testing.assert_array_equal(t[0], [q[0], q[1], q[2]])          # quad a b c d -> a b c, a c d
    np.testing.assert_array_equal(t[1], [q[0], q[2], q[3]])
    p = [[0, 1, 0], [2, 1, 0], [3, 1, 1], [2, 1, 2], [0, 1, 2]]
    np.testing.assert_array_equal(t[2:5], [[p[0], p[1], p[2]], [p[0], p[2], p[3]], [p[0], p[3], p[4]]])
    np.testing.assert_array_equal(t[5], [q[0], q[2], p[0]])          # absolute indices 1, 3, 5
    cols = sc.diffuse_colors()
    np.testing.assert_allclose(cols[0], [0.8, 0.1, 0.05])
    assert cols[2] is None and cols[5] is None                       # unknown material = None (obj.rs:38-48)


def test_obj_errors():
    with pytest.raises(IndexError):
        ingest.parse_obj("v 0 0 0\nv 1 0 0\nf 1 2 3\n")
    with pytest.raises(ValueError):
        ingest.parse_obj("v 0 0 0\nv 1 0 0\nf 1 2\n")
    assert ingest.parse_obj("# nothing\n").triangles().shape == (0, 3, 3)


def test_obj_golden_counts():
    g = golden()
    # face statistics of the reference's files: arrow 8 faces (4 quads + 4 triangles), axis 52 triangles,
    # ground plane 64 quads
    assert g["obj_arrow"].shape == (12, 3, 3)
    assert g["obj_axis"].shape == (52, 3, 3)
    assert g["obj_ground_plane_20x20"].shape == (128, 3, 3)
    gp = g["obj_ground_plane_20x20"]
    assert abs(gp[..., 0].min() + 10) < 1e-4 and abs(gp[..., 2].max() - 10) < 1e-4


@have_reference
def test_obj_loader_reproduces_golden():
    g = golden()
    for name in ("arrow", "axis", "ground-plane-20x20"):
        t = ingest.load_obj(RES + f"obj/{name}.obj").triangles()
        np.testing.assert_array_equal(t, g["obj_" + name.replace("-", "_")])
    # every `f` line is a polygon of k corners -> k - 2 triangles
    for name in ("arrow", "axis", "ground-plane-20x20"):
        with open(RES + f"obj/{name}.obj") as f:
            want = sum(len(line.split()) - 3 for line in f if line.startswith("f "))
        assert len(ingest.load_obj(RES + f"obj/{name}.obj").triangles()) == want


# ------------------------------------------------------------------------------------------ glTF
def make_gltf(index_type=5123, stride=None, glb=False, mode=None, uri_prefix="data:application/octet-stream;base64,"):
    pos = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    idx = np.array([0, 1, 2, 0, 2, 3], {5121: np.uint8, 5123: np.uint16, 5125: np.uint32}[index_type])
    if stride:
        rows = np.zeros((4, stride), np.uint8)
        rows[:, :12] = pos.view(np.uint8).reshape(4, 12)
        pos_bytes = rows.tobytes()
    else:
        pos_bytes = pos.tobytes()
    idx_bytes = idx.tobytes()
    pad = b"\0" * (-len(idx_bytes) % 4)
    blob = idx_bytes + pad + pos_bytes
    view_pos = {"buffer": 0, "byteOffset": len(idx_bytes) + len(pad), "byteLength": len(pos_bytes)}
    if stride:
        view_pos["byteStride"] = stride
    prim = {"attributes": {"POSITION": 1}, "indices": 0}
    if mode is not None:
        prim["mode"] = mode
    doc = {
        "asset": {"version": "2.0"},
        "nodes": [{"mesh": 0, "name": "Quad", "translation": [5, 0, 0]}, {"mesh": 0}, {"mesh": 0, "name": "Again"}],
        "meshes": [{"primitives": [prim, {"attributes": {"POSITION": 1}}]}],      # the 2nd has no indices: skipped
        "buffers": [{"byteLength": len(blob)}],
        "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": len(idx_bytes)}, view_pos],
        "accessors": [{"bufferView": 0, "componentType": index_type, "count": 6, "type": "SCALAR"},
                      {"bufferView": 1, "componentType": 5126, "count": 4, "type": "VEC3"}],
    }
    if glb:
        js = json.dumps(doc).encode()
        js += b" " * (-len(js) % 4)
        bb = blob + b"\0" * (-len(blob) % 4)
        body = struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(bb), 0x004E4942) + bb
        return b"glTF" + struct.pack("<II", 2, 12 + len(body)) + body
    doc["buffers"][0]["uri"] = uri_prefix + base64.b64encode(blob).decode()
    return json.dumps(doc).encode()


@pytest.mark.parametrize("kw", [{}, {"index_type": 5121}, {"index_type": 5125}, {"stride": 16}, {"glb": True}])
def test_gltf_named_nodes_indices_strides_glb(kw):
    out = ingest.load_gltf(make_gltf(**kw))
    assert list(out) == ["Quad", "Again"]                            # unnamed nodes are not reported (loader.rs:29)
    want = np.array([[[0, 0, 0], [1, 0, 0], [1, 1, 0]], [[0, 0, 0], [1, 1, 0], [0, 1, 0]]], np.float32)
    np.testing.assert_array_equal(out["Quad"], want)                 # node translation is ignored, like the reference
    np.testing.assert_array_equal(out["Again"], want)


def test_gltf_errors_match_the_reference_variants():
    with pytest.raises(ingest.MeshLoadError) as e:
        ingest.load_gltf(make_gltf(mode=1))
    assert e.value.kind == "UnsupportedPrimitiveMode"
    with pytest.raises(ingest.MeshLoadError) as e:
        ingest.load_gltf(make_gltf(uri_prefix="file:"))
    assert e.value.kind == "UnsupportedBufferFormat"
    with pytest.raises(ingest.MeshLoadError) as e:
        ingest.load_gltf(make_gltf(uri_prefix="data:application/octet-stream;base64,@@"))
    assert e.value.kind == "Decode"
    doc = json.loads(make_gltf())
    del doc["buffers"][0]["uri"]
    with pytest.raises(ingest.MeshLoadError) as e:
        ingest.load_gltf(json.dumps(doc).encode())
    assert e.value.kind == "MissingBlob"
    with pytest.raises(ingest.MeshLoadError) as e:
        ingest.load_gltf(b"{ not json")
    assert e.value.kind == "Gltf"


def test_gltf_golden_is_two_cubes():
    g = golden()
    for k in ("gltf_Cube_0", "gltf_Cube_1"):
        t = g[k]
        assert t.shape == (12, 3, 3) and t.min() == -1 and t.max() == 1
        n = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
        c = t.mean(axis=1)
        assert ((n * c).sum(axis=1) > 0).all()                       # outward winding


@have_reference
def test_gltf_loader_reproduces_golden():
    # loader.rs:112-124 `load_gltf_test`: the file loads; plus our golden soups
    out = ingest.load_gltf(RES + "gltf/test.gltf")
    g = golden()
    assert list(out) == ["Cube.0", "Cube.1"]
    np.testing.assert_array_equal(out["Cube.0"], g["gltf_Cube_0"])
    np.testing.assert_array_equal(out["Cube.1"], g["gltf_Cube_1"])


# ------------------------------------------------------------------------------------------ .vox
def test_vox_round_trip_and_axis_swap():
    rng = np.random.default_rng(5)
    ids = (rng.random((5, 7, 3)) < 0.4) * rng.integers(1, 200, (5, 7, 3))        # [x, y_up, z]
    ids = ids.astype(np.uint8)
    pal = [(0xFF000000 | (i * 0x010203)) & 0xFFFFFFFF for i in range(256)]
    data = ingest.write_vox(ids, pal)
    dv = ingest.parse_vox(data)
    assert dv.version == 150 and dv.models[0].size == (5, 3, 7)                  # stored z-up: (x, engine z, engine y)
    vox = ingest.load_vox(data)
    assert (vox.x_size, vox.y_size, vox.z_size) == (5, 7, 3)                     # vox_loader.rs:41
    for x, y, z in np.ndindex(5, 7, 3):
        want = None if ids[x, y, z] == 0 else int(ids[x, y, z]) - 1              # dot_vox stores index - 1
        assert vox.get(x, y, z) == want
    cid = int(ids[ids > 0][0]) - 1
    r, g, b = vox.get_color(cid)
    v = pal[cid]
    assert (r, g, b) == (np.float32(v & 255) / np.float32(255), np.float32((v >> 8) & 255) / np.float32(255),
                         np.float32((v >> 16) & 255) / np.float32(255))          # vox_loader.rs:52-55
    np.testing.assert_array_equal(vox.solid, ids > 0)
    with pytest.raises(ValueError):
        ingest.parse_vox(b"NOPE" + data[4:])


def test_chunk_number_and_offset_kat():
    f = ingest.World.chunk_number_and_offset
    # world.rs:371-376 `offset_test`
    assert f(-5, 32) == (-1, 27)
    assert f(2, 4) == (0, 2)
    assert f(5, 4) == (1, 1)
    # the fourth line of that test expects (-2, 3); the code under test computes 4 - (5 % (4 + 1)) = 4
    assert f(-5, 4) == (-2, 4)
    assert f(-1, 32) == (-1, 31) and f(-32, 32) == (-1, 0) and f(-33, 32) == (-2, 32)   # the `% (n + 1)` quirk


def test_position_to_chunk_kat():
    f = ingest.World.position_to_chunk                     # chunks.rs:137-162 `position_to_chunk_test` (32, 0.1)
    assert f([0.0001] * 3) == [0, 0, 0]
    assert f([-0.0001] * 3) == [-1, -1, -1]
    assert f([3.1999] * 3) == [0, 0, 0]
    assert f([3.2001] * 3) == [1, 1, 1]
    assert f([-3.1999] * 3) == [-1, -1, -1]
    assert f([-3.2001] * 3) == [-2, -2, -2]


def visible_faces(grid):
    """Unit faces world.rs:262-279 exposes: (axis, plane, a, b, sign) for every voxel side whose neighbour is
    empty, out of the grid, or of a DIFFERENT colour id."""
    out = {}
    sx, sy, sz = grid.shape
    for x, y, z in np.ndindex(sx, sy, sz):
        c = grid[x, y, z]
        if c < 0:
            continue
        p = (x, y, z)
        for u in range(3):
            for sign in (-1, 1):
                q = list(p)
                q[u] += sign
                inside = 0 <= q[u] < grid.shape[u]
                nb = grid[tuple(q)] if inside else -1
                if nb == c:
                    continue
                plane = p[u] + (1 if sign == 1 else 0)
                rest = tuple(p[k] for k in range(3) if k != u)
                out[(u, plane, rest, sign)] = int(c)
    return out


def rasterise(tris, colors, voxel=0.1):
    """Quads (pairs of triangles) -> the unit faces they cover, with the outward direction of the winding."""
    got = {}
    t = np.asarray(tris, np.float64) / voxel
    for k in range(0, len(t), 2):
        quad = np.concatenate([t[k], t[k + 1]])
        lo, hi = np.rint(quad.min(axis=0)).astype(int), np.rint(quad.max(axis=0)).astype(int)
        u = int(np.argmin(hi - lo))
        assert hi[u] == lo[u]
        n0 = np.cross(t[k, 1] - t[k, 0], t[k, 2] - t[k, 0])
        n1 = np.cross(t[k + 1, 1] - t[k + 1, 0], t[k + 1, 2] - t[k + 1, 0])
        assert n0[u] * n1[u] > 0                                     # both triangles wind the same way
        sign = 1 if n0[u] > 0 else -1
        rest_axes = [a for a in range(3) if a != u]
        for a in range(lo[rest_axes[0]], hi[rest_axes[0]]):
            for b in range(lo[rest_axes[1]], hi[rest_axes[1]]):
                key = (u, int(lo[u]), (a, b), sign)
                assert key not in got, "two quads cover the same face"
                got[key] = colors[k]
    return got


def test_greedy_mesh_covers_exactly_the_visible_faces():
    rng = np.random.default_rng(11)
    ids = ((rng.random((6, 5, 7)) < 0.55) * rng.integers(1, 4, (6, 5, 7))).astype(np.uint8)   # 3 colours, dense
    vox = ingest.load_vox(ingest.write_vox(ids))
    tris, cols = ingest.greedy_mesh(vox, with_colors=True)
    assert len(tris) % 2 == 0 and len(tris) == len(cols)
    grid = vox.data.transpose(2, 1, 0).astype(int)
    want = visible_faces(grid)
    color_key = {vox.get_color(c): c for c in vox.palette}
    got = rasterise(tris, [color_key[tuple(c)] for c in cols.tolist()] if len(color_key) == len(vox.palette) else None)
    assert set(got) == set(want)
    # fewer quads than faces: the merge works
    assert len(tris) // 2 < len(want)


def test_greedy_mesh_emission_order_and_vertices():
    # one voxel at (1, 2, 3): six quads, descriptor order -x +x -y +y -z +z (world.rs:238-245), vertices
    # base, base+dv+dw, base+dv, base+dw with indices 0 1 2 / 0 3 1 (step 1) or 0 2 1 / 0 1 3 (world.rs:317-351)
    ids = np.zeros((4, 4, 4), np.uint8)
    ids[1, 2, 3] = 7
    tris = ingest.greedy_mesh(ingest.load_vox(ingest.write_vox(ids)))
    assert tris.shape == (12, 3, 3)
    f = np.float32
    b = np.array([f(1) / f(10), f(2) / f(10), f(3) / f(10)], np.float32)
    e = np.float32(1) / np.float32(10)
    dx, dy, dz = np.array([e, 0, 0], np.float32), np.array([0, e, 0], np.float32), np.array([0, 0, e], np.float32)
    # -x face: u=0, v=1 (dv=dy), w=2 (dw=dz), step 1
    np.testing.assert_array_equal(tris[0], [b, b + dy + dz, b + dy])
    np.testing.assert_array_equal(tris[1], [b, b + dz, b + dy + dz])
    # +x face: base[u] = 1/10 + 1/10, step -1
    bx = b.copy()
    bx[0] = f(1) / f(10) + f(1) / f(10)
    np.testing.assert_array_equal(tris[2], [bx, bx + dy, bx + dy + dz])
    np.testing.assert_array_equal(tris[3], [bx, bx + dy + dz, bx + dz])
    n = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0])
    want_axis = [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2]
    want_sign = [-1, -1, 1, 1, -1, -1, 1, 1, -1, -1, 1, 1]
    for k in range(12):
        assert np.sign(n[k, want_axis[k]]) == want_sign[k] and np.count_nonzero(n[k]) == 1
    assert ingest.greedy_mesh(ingest.Vox(3, 3, 3)) is None           # untouched grid: no mesh (world.rs:355-363)


def test_greedy_mesh_agrees_with_the_boolean_mesher_on_one_colour():
    solid = scenes.procedural_chunk((0, -1, 0))
    ids = solid.astype(np.uint8) * 200
    tris = ingest.greedy_mesh(ingest.load_vox(ingest.write_vox(ids)))
    other = scenes.greedy_voxel_mesh(solid)

    def area(t):
        return 0.5 * np.linalg.norm(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]), axis=1).sum()
    assert abs(area(tris) - area(other)) < 1e-3 * area(other)
    lo, hi = tris.reshape(-1, 3).min(axis=0), tris.reshape(-1, 3).max(axis=0)
    np.testing.assert_allclose(lo, other.reshape(-1, 3).min(axis=0), atol=1e-6)
    np.testing.assert_allclose(hi, other.reshape(-1, 3).max(axis=0), atol=1e-6)


def test_world_places_models_on_the_chunk_grid():
    ids = np.zeros((40, 10, 34), np.uint8)
    ids[:, :, :] = 3
    vox = ingest.load_vox(ingest.write_vox(ids))
    w = ingest.World()
    w.add(vox, [30, 32, 0])                                         # x: chunk 0 gets 2 columns, chunk 1 gets 32, chunk 2 gets 6
    assert w.occupied_chunks() == [(0, 1, 0), (0, 1, 1), (1, 1, 0), (1, 1, 1), (2, 1, 0), (2, 1, 1)]
    ent, src, dst, size = w.chunk_entity_map[(1, 1, 1)]
    assert (src, dst, size) == ((2, 0, 32), (0, 0, 0), (32, 10, 2))
    cv = w.chunk_vox((0, 1, 0))
    assert cv.get(30, 0, 0) == 2 and cv.get(31, 9, 31) == 2 and cv.get(29, 0, 0) is None     # id 3 is stored as 2
    t = w.chunk_triangles((2, 1, 0))                                # y in [3.2, 6.4): above the procedural fill
    assert len(t) and abs(t[..., 0].min() - 6.4) < 1e-5 and abs(t[..., 0].max() - 7.0) < 1e-5
    assert abs(t[..., 1].min() - 3.2) < 1e-5 and abs(t[..., 1].max() - 4.2) < 1e-5
    # chunk (0,-1,0) holds only the procedural fill: same surface as scenes.procedural_chunk
    t = w.chunk_triangles((0, -1, 0))
    assert len(t) and t[..., 1].max() <= 1e-6


def test_vox_golden_chunks():
    g = golden()
    n_tris, n_vox, sx, sy, sz = g["vox_whole_count"]
    assert (sx, sy, sz) == (126, 126, 126) and n_vox == 117429       # SIZE / XYZI chunks of #treehouse.vox
    assert n_tris == 97924
    for k in ("vox_chunk_1_1_1", "vox_chunk_2_0_1"):
        t = g[k]
        assert len(t) % 2 == 0 and t.dtype == np.float32
        c = [int(v) for v in k.split("_")[2:]]
        lo, hi = t.reshape(-1, 3).min(axis=0), t.reshape(-1, 3).max(axis=0)
        assert (lo >= np.array(c) * 3.2 - 1e-4).all() and (hi <= (np.array(c) + 1) * 3.2 + 1e-4).all()


@have_reference
def test_vox_loader_reproduces_golden():
    g = golden()
    vox = ingest.load_vox(RES + "vox-models/#treehouse/#treehouse.vox")
    assert int((vox.data >= 0).sum()) == 117429
    w = ingest.World()
    w.add(vox, [0, 0, 0])
    assert len(w.occupied_chunks()) == 64
    np.testing.assert_array_equal(w.chunk_triangles((1, 1, 1)), g["vox_chunk_1_1_1"])
    np.testing.assert_array_equal(w.chunk_triangles((2, 0, 1)), g["vox_chunk_2_0_1"])
