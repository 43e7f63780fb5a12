"""Generates tests/golden/ingest_scenes.npz: triangle soups produced by xp_physics_cuda.ingest from the
reference's own assets (res/obj/*.obj, res/gltf/test.gltf, res/vox-models/#treehouse/#treehouse.vox).
Run here, where /root/reference is mounted; the GPU box only sees the committed arrays.

    python tests/golden/make_ingest_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "xp-game-engine_b200"))
from xp_physics_cuda import ingest  # noqa: E402

RES = "/root/reference/res/"


def main():
    out = {}
    for name in ("arrow", "axis", "ground-plane-20x20"):
        out["obj_" + name.replace("-", "_")] = ingest.load_obj(RES + f"obj/{name}.obj").triangles()
    for k, v in ingest.load_gltf(RES + "gltf/test.gltf").items():
        out["gltf_" + k.replace(".", "_")] = v
    vox = ingest.load_vox(RES + "vox-models/#treehouse/#treehouse.vox")
    world = ingest.World()
    world.add(vox, [0, 0, 0])                      # main.rs:83
    out["vox_chunk_1_1_1"] = world.chunk_triangles((1, 1, 1))
    out["vox_chunk_2_0_1"] = world.chunk_triangles((2, 0, 1))
    whole = ingest.greedy_mesh(vox)
    out["vox_whole_count"] = np.array([len(whole), int((vox.data >= 0).sum()), vox.x_size, vox.y_size, vox.z_size], np.int64)
    out["vox_whole_bounds"] = np.stack([whole.reshape(-1, 3).min(axis=0), whole.reshape(-1, 3).max(axis=0)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ingest_scenes.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
