"""World-size-2 gloo test of the multi-GPU host logic (sharding + contact all-gather).  No GPU
here, so each rank's contact records come from the CPU oracle (test infrastructure) standing in
for xp_detect_contacts_dev; what is under test is shard_range / gather_contacts and the record
layout the GPU path uses."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "xp-game-engine_b200")]
    import torch
    import torch.distributed as dist
    from oracle import xp_oracle as o
    from xp_physics_cuda import CONTACT_DTYPE, scenes
    from xp_physics_cuda.distributed import gather_contacts, shard_capacity, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    tris, h = scenes.sine_terrain(10, 10, seed=5)
    resps = scenes.terrain_spheres(h, n, seed=6)
    lo, hi = shard_range(n, rank, world)
    col, hit, tri, st, _ = o.detect_batch(resps[lo:hi], tris, threads=1)
    rec = np.zeros(shard_capacity(n, world), CONTACT_DTYPE)
    for f in ("time_to", "distance_to", "position", "intersection"):
        rec[f][:hi - lo] = col[f] * hit[:, None] if col[f].ndim == 2 else col[f] * hit
    rec["tri_index"][:hi - lo] = tri
    rec["flags"][:hi - lo] = hit.astype(np.uint32) | (st << 8)
    local = torch.from_numpy(rec.view(np.float32).reshape(-1, 10).copy())
    full = gather_contacts(local, n)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [257, 64])
def test_sharded_contact_allgather_gloo(tmp_path, n, oracle):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    sys.path[:0] = [os.path.join(ROOT, "xp-game-engine_b200")]
    from xp_physics_cuda import CONTACT_DTYPE, scenes
    tris, h = scenes.sine_terrain(10, 10, seed=5)
    resps = scenes.terrain_spheres(h, n, seed=6)
    col, hit, tri, st, _ = oracle.detect_batch(resps, tris, threads=1)
    got = [np.load(tmp_path / f"rank{r}.npy").view(CONTACT_DTYPE).reshape(-1) for r in range(world)]
    assert np.array_equal(got[0].view(np.uint32), got[1].view(np.uint32))       # every rank holds the same list
    g = got[0]
    assert len(g) == n
    assert np.array_equal(g["flags"] & 1, hit)
    assert np.array_equal(g["tri_index"], tri)
    hh = hit.astype(bool)
    assert np.array_equal(g["time_to"][hh].view(np.uint32), col["time_to"][hh].view(np.uint32))
    assert np.array_equal(g["position"][hh].view(np.uint32), col["position"][hh].view(np.uint32))


def test_shard_ranges_partition():
    from xp_physics_cuda.distributed import shard_capacity, shard_range
    for n in (0, 1, 7, 8, 9, 1 << 20, (1 << 27) + 3):
        for world in (1, 2, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) <= shard_capacity(n, world)
