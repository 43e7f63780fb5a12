"""Mesh ingest: OBJ / glTF / MagicaVoxel .vox -> triangle soups.

Host-side mirrors of the reference's asset loaders, so that the scenes the reference feeds its
collision backends can be handed to `Scene.set_triangles` unchanged:

* `load_obj`      -- xp_mesh/src/obj.rs:13-75 (`Obj::load` + the triangle iterator; tobj with
                     `triangulate = true`)
* `load_gltf`     -- low_poly_nice_graphics/src/gltf/loader.rs:25-110
* `load_vox`      -- low_poly_nice_graphics/src/vox/vox_loader.rs:39-50 on top of the `dot_vox` 4.1
                     container parser (third-party, not in the reference tree: the published
                     MagicaVoxel .vox chunk format is restated here)
* `greedy_mesh`   -- low_poly_nice_graphics/src/world/world.rs:234-365, quad for quad and in the same
                     emission order (colour-aware merge)
* `World`         -- world.rs:60-190: voxel models placed on a 32^3 chunk grid, a chunk = procedural
                     sin*sin fill overlaid with the model's voxels, meshed and translated

Everything returns float32 arrays shaped (n, 3, 3) -- n triangles x 3 vertices x xyz -- in the order
the reference produces them.  Decimal text is parsed through Python's float (binary64) and then
rounded to binary32; Rust parses straight to binary32, which can differ only for literals with more
than ~16 significant digits that sit on a rounding boundary.
"""
from __future__ import annotations

import base64
import json
import os
import struct
from dataclasses import dataclass, field

import numpy as np


# ---------------------------------------------------------------------------------------------
# OBJ
# ---------------------------------------------------------------------------------------------
@dataclass
class ObjModel:
    """One tobj model: a run of faces under one `o`/`g` name and one material."""
    name: str
    positions: np.ndarray          # (n_tris, 3, 3) float32
    material: str | None


@dataclass
class ObjScene:
    models: list[ObjModel]
    materials: dict[str, np.ndarray] = field(default_factory=dict)     # name -> Kd (3,) float32

    def triangles(self) -> np.ndarray:
        """All triangles, model by model, in face order (the `Obj` iterator, obj.rs:25-75)."""
        parts = [m.positions for m in self.models if len(m.positions)]
        if not parts:
            return np.zeros((0, 3, 3), np.float32)
        return np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)

    def diffuse_colors(self) -> list[np.ndarray | None]:
        """Per triangle: the material's Kd, or None when the model has no (known) material (obj.rs:38-48)."""
        out: list[np.ndarray | None] = []
        for m in self.models:
            kd = self.materials.get(m.material) if m.material is not None else None
            out.extend([kd] * len(m.positions))
        return out


def _parse_mtl(text: str) -> dict[str, np.ndarray]:
    mats: dict[str, np.ndarray] = {}
    cur = None
    for line in text.splitlines():
        tok = line.split()
        if not tok or tok[0].startswith("#"):
            continue
        if tok[0] == "newmtl":
            cur = " ".join(tok[1:])
            mats[cur] = np.zeros(3, np.float32)
        elif tok[0] == "Kd" and cur is not None:
            mats[cur] = np.array([float(t) for t in tok[1:4]], np.float32)
    return mats


def _obj_index(tok: str, n_positions: int) -> int:
    """Position index of a face corner `v`, `v/vt`, `v//vn` or `v/vt/vn`; negative = relative to the end."""
    v = int(tok.split("/")[0])
    if v > 0:
        return v - 1
    if v < 0:
        return n_positions + v
    raise ValueError("OBJ face index 0 is invalid")


def parse_obj(text: str, mtl_loader=None) -> ObjScene:
    """Parse OBJ text the way `tobj::load_obj(file, true)` presents it to obj.rs: polygons are fan-
    triangulated around their first corner (a quad a b c d -> a b c, a c d), a new model starts at
    every `o` / `g` / `usemtl` that follows faces, and triangles keep file order."""
    positions: list[tuple[float, float, float]] = []
    models: list[ObjModel] = []
    materials: dict[str, np.ndarray] = {}
    name, material = "unnamed_object", None
    faces: list[tuple[int, int, int]] = []

    def flush():
        nonlocal faces
        if faces:
            p = np.asarray(positions, np.float64)
            idx = np.asarray(faces, np.int64)
            if idx.min() < 0 or idx.max() >= len(p):
                raise IndexError("OBJ face references a vertex that does not exist")
            models.append(ObjModel(name, p[idx].astype(np.float32), material))
        faces = []

    for raw in text.splitlines():
        line = raw.strip()
        if not line or line.startswith("#"):
            continue
        tok = line.split()
        key = tok[0]
        if key == "v":
            positions.append((float(tok[1]), float(tok[2]), float(tok[3])))
        elif key == "f":
            corners = [_obj_index(t, len(positions)) for t in tok[1:]]
            if len(corners) < 3:
                raise ValueError("OBJ face with fewer than 3 corners")
            a, b = corners[0], corners[1]
            for c in corners[2:]:
                faces.append((a, b, c))
                b = c
        elif key in ("o", "g"):
            flush()
            name = " ".join(tok[1:]) if len(tok) > 1 else "unnamed_object"
        elif key == "usemtl":
            flush()
            material = " ".join(tok[1:]) if len(tok) > 1 else None
        elif key == "mtllib" and mtl_loader is not None:
            for lib in tok[1:]:
                try:
                    materials.update(_parse_mtl(mtl_loader(lib)))
                except OSError:
                    pass                                  # tobj reports a missing .mtl and carries on without it
    flush()
    for m in models:                                      # `usemtl (null)` & co: unknown name = no material
        if m.material is not None and m.material not in materials:
            m.material = None
    return ObjScene(models, materials)


def load_obj(path: str) -> ObjScene:
    """`Obj::load(file_name)` (obj.rs:13-23).  `.triangles()` is the soup for `Scene.set_triangles`."""
    base = os.path.dirname(os.path.abspath(path))
    with open(path, "r", encoding="utf-8", errors="replace") as f:
        text = f.read()

    def mtl(name):
        with open(os.path.join(base, name), "r", encoding="utf-8", errors="replace") as g:
            return g.read()
    return parse_obj(text, mtl)


# ---------------------------------------------------------------------------------------------
# glTF
# ---------------------------------------------------------------------------------------------
class MeshLoadError(Exception):
    """loader.rs:4-11; `.kind` is the variant name."""
    def __init__(self, kind: str, detail: str = ""):
        self.kind = kind
        super().__init__(f"{kind}: {detail}" if detail else kind)


_OCTET = "data:application/octet-stream;base64,"
_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_WIDTH = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT2": 4, "MAT3": 9, "MAT4": 16}


def _split_glb(data: bytes):
    if data[:4] != b"glTF":
        return json.loads(data.decode("utf-8")), None
    version, length = struct.unpack_from("<II", data, 4)
    if version != 2:
        raise MeshLoadError("Gltf", f"unsupported GLB version {version}")
    off, doc, blob = 12, None, None
    while off + 8 <= min(length, len(data)):
        n, kind = struct.unpack_from("<II", data, off)
        body = data[off + 8: off + 8 + n]
        if kind == 0x4E4F534A:
            doc = json.loads(body.decode("utf-8"))
        elif kind == 0x004E4942 and blob is None:
            blob = bytes(body)
        off += 8 + n + (-n % 4)
    if doc is None:
        raise MeshLoadError("Gltf", "GLB without a JSON chunk")
    return doc, blob


def _gltf_buffers(doc, blob):
    """loader.rs:87-110: only embedded base64 octet streams and the GLB blob are accepted."""
    out = []
    for buf in doc.get("buffers", []):
        uri = buf.get("uri")
        if uri is not None:
            if not uri.startswith(_OCTET):
                raise MeshLoadError("UnsupportedBufferFormat", uri[:40])
            try:
                out.append(base64.b64decode(uri[len(_OCTET):], validate=True))
            except Exception as e:                       # noqa: BLE001 -- base64::DecodeError
                raise MeshLoadError("Decode", str(e)) from e
        else:
            if blob is None:
                raise MeshLoadError("MissingBlob")
            out.append(blob)
    return out


def _gltf_accessor(doc, buffers, index: int) -> np.ndarray:
    acc = doc["accessors"][index]
    if "sparse" in acc:
        raise MeshLoadError("Gltf", "sparse accessors are not supported")
    dtype = np.dtype(_COMPONENT[acc["componentType"]]).newbyteorder("<")
    width = _WIDTH[acc["type"]]
    count = acc["count"]
    if "bufferView" not in acc:
        return np.zeros((count, width), dtype)
    view = doc["bufferViews"][acc["bufferView"]]
    raw = buffers[view["buffer"]]
    start = view.get("byteOffset", 0) + acc.get("byteOffset", 0)
    elem = dtype.itemsize * width
    stride = view.get("byteStride") or elem
    if count and start + stride * (count - 1) + elem > len(raw):
        raise MeshLoadError("Gltf", "accessor runs past the end of its buffer")
    if stride == elem:
        return np.frombuffer(raw, dtype, count * width, start).reshape(count, width)
    rows = np.lib.stride_tricks.as_strided(np.frombuffer(raw, np.uint8, len(raw) - start, start),
                                           (count, elem), (stride, 1))
    return np.ascontiguousarray(rows).view(dtype).reshape(count, width)


def load_gltf(data: bytes | str) -> dict[str, np.ndarray]:
    """`load_gltf(bytes, named_mesh)` (loader.rs:25-85): one entry per NAMED node, in document order,
    holding the node's mesh as a triangle soup in mesh-local space (the reference ignores node
    transforms, `loader.rs:28-30`).  Later nodes with an earlier node's name replace it, like repeated
    `named_mesh` callbacks into a map would.  `data` is the file's bytes (JSON or GLB) or a path."""
    if isinstance(data, str):
        with open(data, "rb") as f:
            data = f.read()
    try:
        doc, blob = _split_glb(bytes(data))
    except (ValueError, UnicodeDecodeError) as e:
        raise MeshLoadError("Gltf", str(e)) from e
    buffers = _gltf_buffers(doc, blob)
    out: dict[str, np.ndarray] = {}
    for node in doc.get("nodes", []):
        name = node.get("name")
        if name is None:
            continue
        if "mesh" not in node:
            raise ValueError(f"named glTF node {name!r} has no mesh (the reference unwraps it, loader.rs:30)")
        parts = []
        for prim in doc["meshes"][node["mesh"]].get("primitives", []):
            if prim.get("mode", 4) != 4:
                raise MeshLoadError("UnsupportedPrimitiveMode", str(prim.get("mode")))
            pos_acc = prim.get("attributes", {}).get("POSITION")
            if pos_acc is None or "indices" not in prim:
                continue                                  # loader.rs:39-44: silently skipped
            pos = _gltf_accessor(doc, buffers, pos_acc).astype(np.float32)
            idx = _gltf_accessor(doc, buffers, prim["indices"]).astype(np.int64).reshape(-1)
            if len(idx) % 3:
                raise AssertionError("glTF index count is not a multiple of 3 (loader.rs:45)")
            if len(idx) and (idx.max() >= len(pos)):
                raise IndexError("glTF index out of range")
            parts.append(pos[idx].reshape(-1, 3, 3))
        out[name] = (np.ascontiguousarray(np.concatenate(parts), dtype=np.float32) if parts
                     else np.zeros((0, 3, 3), np.float32))
    return out


# ---------------------------------------------------------------------------------------------
# MagicaVoxel .vox
# ---------------------------------------------------------------------------------------------
class Vox:
    """vox_loader.rs:3-37: a dense grid of optional colour ids + palette; `data[z][y][x]`, y up."""
    def __init__(self, x_size: int, y_size: int, z_size: int):
        self.x_size, self.y_size, self.z_size = int(x_size), int(y_size), int(z_size)
        self.data = np.full((self.z_size, self.y_size, self.x_size), -1, np.int16)     # -1 = None
        self.palette: dict[int, tuple[float, float, float]] = {}
        self.touched = False

    def set(self, x: int, y: int, z: int, color_id: int, color=(1.0, 0.0, 0.0)):
        self.touched = True
        self.data[z, y, x] = color_id
        self.palette[int(color_id)] = tuple(color)

    def get(self, x: int, y: int, z: int):
        v = int(self.data[z, y, x])
        return None if v < 0 else v

    def get_color(self, color_id: int):
        return self.palette[int(color_id)]

    @property
    def solid(self) -> np.ndarray:
        """bool[x, y, z] occupancy (the layout `scenes.greedy_voxel_mesh` takes)."""
        return np.ascontiguousarray((self.data >= 0).transpose(2, 1, 0))


def _default_palette() -> list[int]:
    # MagicaVoxel's built-in palette follows a fixed pattern; only files without an RGBA chunk use it
    # and only for colours, which no geometry depends on: the 6x6x6 cube + ramps, as published.
    pal = [0x00000000]
    levels = [0xFF, 0xCC, 0x99, 0x66, 0x33, 0x00]
    for b in levels:
        for g in levels:
            for r in levels:
                pal.append(0xFF000000 | (b << 16) | (g << 8) | r)
    pal = pal[:216]
    ramp = [0xEE, 0xDD, 0xBB, 0xAA, 0x88, 0x77, 0x55, 0x44, 0x22, 0x11]
    for v in ramp: pal.append(0xFF000000 | v)                        # reds
    for v in ramp: pal.append(0xFF000000 | (v << 8))                 # greens
    for v in ramp: pal.append(0xFF000000 | (v << 16))                # blues
    for v in ramp: pal.append(0xFF000000 | (v << 16) | (v << 8) | v) # greys
    return (pal + [0xFF000000] * 256)[:256]


@dataclass
class DotVoxModel:
    size: tuple[int, int, int]               # (x, y, z) as stored: z is up in MagicaVoxel
    voxels: np.ndarray                       # (n, 4) uint8: x, y, z, i  with i = stored index - 1


@dataclass
class DotVoxData:
    version: int
    models: list[DotVoxModel]
    palette: list[int]                       # 256 x 0xAABBGGRR


def parse_vox(data: bytes) -> DotVoxData:
    """The container: 'VOX ' + version, a MAIN chunk whose children are SIZE / XYZI pairs (one per
    model) and an optional RGBA palette.  Like dot_vox, a voxel's colour index is stored minus one so
    that it indexes the palette directly."""
    if len(data) < 20 or data[:4] != b"VOX ":
        raise ValueError("not a MagicaVoxel .vox file")
    version = struct.unpack_from("<i", data, 4)[0]
    cid, n, m = struct.unpack_from("<4sii", data, 8)
    if cid != b"MAIN":
        raise ValueError(".vox file without a MAIN chunk")
    off, end = 20 + n, min(len(data), 20 + n + m)
    models: list[DotVoxModel] = []
    palette = _default_palette()
    size = None
    while off + 12 <= end:
        cid, n, m = struct.unpack_from("<4sii", data, off)
        body = off + 12
        if cid == b"SIZE":
            size = struct.unpack_from("<iii", data, body)
        elif cid == b"XYZI":
            cnt = struct.unpack_from("<i", data, body)[0]
            vox = np.frombuffer(data, np.uint8, cnt * 4, body + 4).reshape(cnt, 4).copy()
            vox[:, 3] = np.where(vox[:, 3] > 0, vox[:, 3] - 1, 0)
            if size is None:
                raise ValueError("XYZI chunk without a SIZE chunk")
            models.append(DotVoxModel(tuple(int(v) for v in size), vox))
            size = None
        elif cid == b"RGBA":
            palette = [int(v) for v in np.frombuffer(data, "<u4", 256, body)]
        off = body + n + m
    return DotVoxData(version, models, palette)


def _palette_to_color(v: int):
    """vox_loader.rs:52-55."""
    r, g, b = v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF
    return (np.float32(r) / np.float32(255.0), np.float32(g) / np.float32(255.0), np.float32(b) / np.float32(255.0))


def load_vox(data: bytes | str) -> Vox:
    """vox_loader.rs:39-50: the FIRST model of the file, axes swapped so that y is up
    (`Vox::new(size.x, size.z, size.y)`, `set(v.x, v.z, v.y, ..)`).  `data`: bytes or a path."""
    if isinstance(data, str):
        with open(data, "rb") as f:
            data = f.read()
    dv = parse_vox(bytes(data))
    if not dv.models:
        raise IndexError(".vox file holds no model (the reference indexes models[0])")
    model = dv.models[0]
    vox = Vox(model.size[0], model.size[2], model.size[1])
    v = model.voxels
    if len(v):
        vox.touched = True
        vox.data[v[:, 1].astype(np.intp), v[:, 2].astype(np.intp), v[:, 0].astype(np.intp)] = v[:, 3].astype(np.int16)
        for cid in np.unique(v[:, 3]):
            vox.palette[int(cid)] = _palette_to_color(dv.palette[int(cid)])
    return vox


def write_vox(solid_ids: np.ndarray, palette: list[int] | None = None) -> bytes:
    """A minimal .vox file (for tests and synthetic scenes): `solid_ids[x, y_up, z]` holds a colour
    index 1..255 or 0 for empty; written z-up like MagicaVoxel does."""
    ids = np.asarray(solid_ids, np.uint8)
    sx, sy, sz = ids.shape                                   # engine axes: x, y (up), z
    x, y, z = np.nonzero(ids)
    xyzi = np.stack([x, z, y, ids[x, y, z]], axis=1).astype(np.uint8)      # stored: x, y := engine z, z := engine y
    size = struct.pack("<4sii", b"SIZE", 12, 0) + struct.pack("<iii", sx, sz, sy)
    body = struct.pack("<i", len(xyzi)) + xyzi.tobytes()
    chunk = struct.pack("<4sii", b"XYZI", len(body), 0) + body
    children = size + chunk
    if palette is not None:
        children += struct.pack("<4sii", b"RGBA", 1024, 0) + np.asarray(palette, "<u4").tobytes()
    return b"VOX " + struct.pack("<i", 150) + struct.pack("<4sii", b"MAIN", 0, len(children)) + children


# ---------------------------------------------------------------------------------------------
# greedy mesher + chunked world
# ---------------------------------------------------------------------------------------------
# (u, v, w, step, normal): world.rs:238-245.  The mask of a slice is indexed (v, w) = (x, y) below.
_DESCRIPTORS = (
    (0, 1, 2, 1, (1, 0, 0)), (0, 1, 2, -1, (-1, 0, 0)),
    (1, 2, 0, 1, (0, 1, 0)), (1, 2, 0, -1, (0, -1, 0)),
    (2, 0, 1, 1, (0, 0, 1)), (2, 0, 1, -1, (0, 0, -1)),
)
_F10 = np.float32(10.0)


def greedy_mesh(vox: Vox, with_colors: bool = False):
    """world.rs:234-365.  For each of the six face directions and each slice, a mask holds the colour
    id of every voxel face that is visible (a face between two voxels of the SAME colour id is hidden,
    one between different ids is emitted from both sides, world.rs:272-276); the mask is swept row by
    row, each run is widened then heightened while the id matches, and a quad of two triangles is
    emitted (vertex order world.rs:317-345).  Returns the triangles in chunk-local coordinates
    (voxel = 0.1), in emission order, or None for an untouched grid; with_colors adds (n, 3) float32."""
    if not vox.touched:
        return (None, None) if with_colors else None
    grid = vox.data.transpose(2, 1, 0)                       # [x, y, z] int16, -1 = None
    size = (vox.x_size, vox.y_size, vox.z_size)
    tris: list[np.ndarray] = []
    cols: list[tuple] = []
    for (u, v, w, step, normal) in _DESCRIPTORS:
        qu = 1 if step != 1 else 0                           # q = [1,0,0] for the -1 descriptors: base[u] += 0.1
        for s_i in range(size[u]):
            sl = s_i if step == 1 else size[u] - (s_i + 1)
            cur = np.take(grid, sl, axis=u)                  # remaining axes in increasing axis order
            back_i = sl - normal[u]
            if 0 <= back_i < size[u]:
                back = np.take(grid, back_i, axis=u)
                mask = np.where((back >= 0) & (cur >= 0) & (back == cur), -1, cur)
            else:
                mask = cur.copy()
            # bring the slice to [v, w] order: the remaining axes are sorted ascending, (v, w) is
            # (1,2), (2,0) or (0,1) -> a transpose is needed exactly when v > w
            if v > w:
                mask = mask.T
            mask = np.array(mask, np.int16)                  # own copy, [v][w]
            if not (mask >= 0).any():
                continue
            nv, nw = size[v], size[w]
            for y in range(nw):
                row = mask[:, y]
                if not (row >= 0).any():
                    continue
                x = 0
                while x < nv:
                    cid = int(mask[x, y])
                    if cid < 0:
                        x += 1
                        continue
                    width = 1
                    while x + width < nv and mask[x + width, y] == cid:
                        width += 1
                    height = 1
                    while y + height < nw and (mask[x:x + width, y + height] == cid).all():
                        height += 1
                    base = [np.float32(0)] * 3
                    base[u] = np.float32(sl) / _F10 + np.float32(qu) / _F10
                    base[v] = np.float32(x) / _F10 + np.float32(0) / _F10
                    base[w] = np.float32(y) / _F10 + np.float32(0) / _F10
                    dv = [np.float32(0)] * 3
                    dw = [np.float32(0)] * 3
                    dv[v] = np.float32(width) / _F10
                    dw[w] = np.float32(height) / _F10
                    p0 = base
                    p1 = [np.float32(np.float32(base[k] + dv[k]) + dw[k]) for k in range(3)]
                    p2 = [np.float32(base[k] + dv[k]) for k in range(3)]
                    p3 = [np.float32(base[k] + dw[k]) for k in range(3)]
                    if step == 1:                            # indices 0 1 2, 0 3 1
                        quad = ((p0, p1, p2), (p0, p3, p1))
                    else:                                    # indices 0 2 1, 0 1 3
                        quad = ((p0, p2, p1), (p0, p1, p3))
                    tris.append(np.array(quad, np.float32))
                    if with_colors:
                        cols.extend([vox.get_color(cid)] * 2)
                    mask[x:x + width, y:y + height] = -1
                    x += width
    out = np.concatenate(tris).reshape(-1, 3, 3) if tris else np.zeros((0, 3, 3), np.float32)
    out = np.ascontiguousarray(out, dtype=np.float32)
    if with_colors:
        return out, np.asarray(cols, np.float32).reshape(-1, 3)
    return out


class World:
    """world.rs:60-190: voxel models registered at integer voxel positions, cut into 32^3 chunks
    (voxel 0.1 -> 3.2 units per chunk); a generated chunk is the procedural `sin(x) * sin(z) > y` fill
    (world.rs:146-157) overlaid with the part of a model that falls into it (world.rs:158-179)."""
    CHUNK_SIZE = 32
    VOXEL = np.float32(0.1)

    def __init__(self):
        self.entities: list[tuple[Vox, tuple[int, int, int]]] = []
        self.chunk_entity_map: dict[tuple[int, int, int], tuple] = {}

    @staticmethod
    def chunk_number_and_offset(start: int, chunk_size: int) -> tuple[int, int]:
        """world.rs:78-88, literally: Rust's integer `/` and `%` truncate towards zero and the negative
        branch takes the remainder by `chunk_size + 1`.  The reference's own `offset_test`
        (world.rs:371-376) pins (-5, 32) -> (-1, 27), (2, 4) -> (0, 2), (5, 4) -> (1, 1); its fourth
        expectation, (-5, 4) -> (-2, 3), contradicts the code it tests, which yields (-2, 4) -- this
        mirror follows the code."""
        if start >= 0:
            return start // chunk_size, start % chunk_size
        num = start + 1
        chunk_number = -((-num) // chunk_size) - 1           # trunc((start + 1) / chunk_size) - 1
        offset = chunk_size - ((-start) % (chunk_size + 1))
        return chunk_number, offset

    @classmethod
    def position_to_chunk(cls, position, chunk_size: int | None = None, voxel_size=None) -> list[int]:
        """chunks.rs:60-65 (`Chunks::position_to_chunk`): the chunk a world position falls into."""
        cs = np.float32(cls.CHUNK_SIZE if chunk_size is None else chunk_size)
        vs = cls.VOXEL if voxel_size is None else np.float32(voxel_size)
        return [int(np.floor(np.float32(p) / (cs * vs))) for p in position]

    def add(self, vox: Vox, position) -> None:
        """world.rs:86-136: remember which part of the model lands in which chunk (a later model
        replaces an earlier one chunk by chunk, as the HashMap insert does)."""
        self.entities.append((vox, tuple(int(p) for p in position)))
        ent = len(self.entities) - 1
        cs = self.CHUNK_SIZE

        def spans(start, size):
            number, target = self.chunk_number_and_offset(start, cs)
            source = 0
            while size != 0:
                cur = min(size, cs - target)
                if cur <= 0:
                    raise OverflowError("chunk offset outside the chunk (usize underflow in the reference)")
                yield number, source, target, cur
                number += 1
                source += cur
                target = 0
                size -= cur

        for zn, zs, zt, zc in spans(int(position[2]), vox.z_size):
            for yn, ys, yt, yc in spans(int(position[1]), vox.y_size):
                for xn, xs, xt, xc in spans(int(position[0]), vox.x_size):
                    self.chunk_entity_map[(xn, yn, zn)] = (ent, (xs, ys, zs), (xt, yt, zt), (xc, yc, zc))

    def chunk_vox(self, chunk) -> Vox:
        """The voxel grid `generate_chunk` meshes (world.rs:138-179)."""
        cs = self.CHUNK_SIZE
        out = Vox(cs, cs, cs)
        i = np.arange(cs, dtype=np.float32)
        f = np.float32
        xw = f(chunk[0]) * f(cs) * self.VOXEL + i * self.VOXEL
        yw = f(chunk[1]) * f(cs) * self.VOXEL + i * self.VOXEL
        zw = f(chunk[2]) * f(cs) * self.VOXEL + i * self.VOXEL
        Z, Y, X = np.meshgrid(zw, yw, xw, indexing="ij")
        fill = (Y > f(-5.0)) & ((np.sin(X) * np.sin(Z)).astype(np.float32) > Y)
        if fill.any():
            out.touched = True
            out.data[fill] = 255
            out.palette[255] = (1.0, 0.0, 0.0)
        hit = self.chunk_entity_map.get(tuple(int(c) for c in chunk))
        if hit is not None:
            ent, src, dst, size = hit
            vox = self.entities[ent][0]
            part = vox.data[src[2]:src[2] + size[2], src[1]:src[1] + size[1], src[0]:src[0] + size[0]]
            view = out.data[dst[2]:dst[2] + size[2], dst[1]:dst[1] + size[1], dst[0]:dst[0] + size[0]]
            sel = part >= 0
            if sel.any():
                out.touched = True
                view[sel] = part[sel]
                for cid in np.unique(part[sel]):
                    out.palette[int(cid)] = vox.get_color(int(cid))
        return out

    def chunk_triangles(self, chunk) -> np.ndarray:
        """Mesh of one chunk in world space: `greedy_mesh` + the chunk entity's translation
        (world.rs:180-193); empty for an untouched chunk."""
        tris = greedy_mesh(self.chunk_vox(chunk))
        if tris is None or not len(tris):
            return np.zeros((0, 3, 3), np.float32)
        f = np.float32
        t = np.array([f(chunk[k]) * f(self.CHUNK_SIZE) * self.VOXEL for k in range(3)], np.float32)
        return np.ascontiguousarray(tris + t, dtype=np.float32)

    def triangles(self, chunks) -> np.ndarray:
        parts = [self.chunk_triangles(c) for c in chunks]
        parts = [p for p in parts if len(p)]
        return np.concatenate(parts) if parts else np.zeros((0, 3, 3), np.float32)

    def occupied_chunks(self) -> list[tuple[int, int, int]]:
        return sorted(self.chunk_entity_map)
