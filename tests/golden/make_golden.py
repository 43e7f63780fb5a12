"""Generates the committed fixtures under tests/golden/.  Run in the build container (it reads
/root/reference for the heightmap image; the GPU box has no /root/reference, hence the fixtures).

  heightmap_crop_red.npy  uint8 [201, 201]: red channel of res/map/heightmap.jpg at
                          pixel (x + 2800, z + 2000), x,z in 0..200 -- exactly the pixels
                          low_poly_nice_graphics/src/world/heightmap_loader.rs:12-51 reads
                          (decoded with PIL here; the Rust `image` crate's JPEG decoder may
                          differ from libjpeg by +-1 in isolated pixels).
  golden_oracle.npz       seeded inputs + the oracle's outputs (detect, response, broadphase
                          pairs) for a small terrain scene, so parity tests also have a frozen
                          copy of the oracle's answers that does not depend on the local gcc.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "xp-game-engine_b200")]


def heightmap():
    from PIL import Image
    img = np.asarray(Image.open("/root/reference/res/map/heightmap.jpg").convert("RGB"))
    # image::get_pixel(x, y): x = column, y = row.  red[x, z] = pixel(x + 2800, z + 2000)
    crop = img[2000:2000 + 201, 2800:2800 + 201, 0].T.copy()
    assert crop.shape == (201, 201)
    np.save(os.path.join(HERE, "heightmap_crop_red.npy"), crop.astype(np.uint8))
    print("heightmap crop", crop.shape, crop.min(), crop.max())


def oracle_vectors():
    from oracle import xp_oracle as o
    from xp_physics_cuda import scenes
    tris, h = scenes.sine_terrain(20, 20, seed=21)
    resps = scenes.terrain_spheres(h, 256, seed=22)
    col, hit, ti, st, _ = o.detect_batch(resps, tris)
    out, it, rst, ln, _ = o.collision_response_batch(resps, tris, max_iter=0)
    si, tj = o.broadphase_pairs(resps, tris)
    np.savez_compressed(os.path.join(HERE, "golden_oracle.npz"), triangles=tris, responses=resps,
                        det_col=col, det_hit=hit, det_tri=ti, resp_out=out, resp_iter=it, resp_status=rst,
                        resp_normal=ln, pair_sphere=si, pair_tri=tj)
    print("golden oracle vectors:", len(tris), "triangles", len(resps), "spheres", int(hit.sum()), "hits",
          len(si), "pairs", "iterations hist", np.bincount(it))


if __name__ == "__main__":
    heightmap()
    oracle_vectors()
