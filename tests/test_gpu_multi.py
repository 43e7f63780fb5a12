"""Multi-GPU path on real devices: ShardedSweep (sphere shards, replicated scene, NCCL all-gather of
contact records) against the oracle.  Uses 2 ranks when the box has >= 2 GPUs, else 1 rank (the
collective then degenerates to a copy but the same code runs)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir, exchange="auto"):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "xp-game-engine_b200")]
    import torch
    import torch.distributed as dist
    from xp_physics_cuda import scenes
    from xp_physics_cuda.distributed import ShardedSweep
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    tris, h = scenes.sine_terrain(40, 40, seed=15)
    resps = scenes.terrain_spheres(h, n, seed=16)
    sw = ShardedSweep(tris, rank, exchange=exchange)
    full = sw.sweep(resps)
    full2 = sw.sweep(resps)                      # buffers are reused across sweeps
    torch.cuda.synchronize()
    assert torch.equal(full.view(torch.int32), full2.view(torch.int32))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full.cpu().numpy())
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write(sw.last_exchange)
    sw.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["nccl", "auto"])
def test_sharded_sweep(tmp_path, oracle, exchange):
    import torch
    import torch.multiprocessing as mp
    from xp_physics_cuda import CONTACT_DTYPE, scenes
    world = min(torch.cuda.device_count(), 2)
    n = 5001
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path), exchange), nprocs=world, join=True)
    print("exchange used:", open(tmp_path / "rank0.txt").read())
    tris, h = scenes.sine_terrain(40, 40, seed=15)
    resps = scenes.terrain_spheres(h, n, seed=16)
    col, hit, tri, st, _ = oracle.detect_batch(resps, tris)
    for r in range(world):
        g = np.load(tmp_path / f"rank{r}.npy").view(CONTACT_DTYPE).reshape(-1)
        assert len(g) == n
        assert np.array_equal(g["flags"] & 1, hit) and np.array_equal(g["tri_index"], tri)
        hh = hit.astype(bool)
        for f in ("time_to", "distance_to", "position", "intersection"):
            assert np.array_equal(g[f][hh].view(np.uint32), col[f][hh].view(np.uint32)), f
