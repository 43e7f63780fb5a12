"""Multi-GPU path on real devices, one process per GPU over EVERY visible GPU: ShardedSweep (sphere shards,
replicated scene, contact records gathered through NCCL or through mapped peer memory) against the oracle, and
the asynchronous loop of steps (four gather buffers round robin, copy-engine exchange under the next sweep)
checked record for record on every rank with deliberately skewed rank timing.

Skipped on a box with fewer than two GPUs: with one rank nothing crosses a link and the test would claim
coverage it does not have (the single-process form of the loop is tests/test_gpu_parity.py::
test_async_exchange_loop_single_process)."""
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


def _world():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs, this box shows {n}")
    return n


def _init(rank, world, port):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "xp-game-engine_b200")]
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    return torch, dist


def _worker(rank, world, port, n, out_dir, exchange="auto", record="full"):
    torch, dist = _init(rank, world, port)
    from xp_physics_cuda import scenes
    from xp_physics_cuda.distributed import ShardedSweep
    tris, h = scenes.sine_terrain(40, 40, seed=15)
    resps = scenes.terrain_spheres(h, n, seed=16)
    sw = ShardedSweep(tris, rank, exchange=exchange, record=record)
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
    import torch.multiprocessing as mp
    from xp_physics_cuda import CONTACT_DTYPE, scenes
    world = _world()
    n = 5001
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path), exchange), nprocs=world, join=True)
    used = open(tmp_path / "rank0.txt").read()
    print(f"world {world}, exchange used: {used}")
    if exchange == "auto":
        assert used in ("p2p", "multicast"), "peer-memory exchange unavailable on a multi-GPU box"
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


@pytest.mark.parametrize("exchange", ["nccl", "auto"])
def test_sharded_sweep_compact_records(tmp_path, oracle, exchange):
    """XP_OPT_CONTACT_RECORD = 1: the 16 B record (centre at the contact, triangle) through both transports --
    the fields a peer can use without the owner's Response, bit-identical to the oracle's on every rank."""
    import torch.multiprocessing as mp
    from xp_physics_cuda import CONTACT16_DTYPE, scenes
    world = _world()
    n = 5001
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path), exchange, "compact"), nprocs=world, join=True)
    tris, h = scenes.sine_terrain(40, 40, seed=15)
    resps = scenes.terrain_spheres(h, n, seed=16)
    col, hit, tri, st, _ = oracle.detect_batch(resps, tris)
    hh = hit.astype(bool)
    assert hh.sum() > 1000 and (~hh).sum() > 0
    for r in range(world):
        g = np.load(tmp_path / f"rank{r}.npy")
        assert g.shape == (n, 4)
        g = np.ascontiguousarray(g).view(CONTACT16_DTYPE).reshape(-1)
        assert np.array_equal(g["tri_index"], tri)
        assert np.array_equal(g["position"][hh].view(np.uint32), col["position"][hh].view(np.uint32))
        assert not g["position"][~hh].any() and (g["tri_index"][~hh] == 0xFFFFFFFF).all()


def _loop_worker(rank, world, port, out_dir, steps):
    """The asynchronous loop of steps: `steps` steps over three different batches, two in flight, every rank's
    gather buffer compared with the NCCL all-gather of the same batch after every step.  Ranks are skewed on
    purpose -- host sleeps and device-side spins that differ per rank and per step -- so a fast rank runs ahead
    of a slow one as far as the protocol allows (the race df560e8 fixed needed exactly that)."""
    import time
    torch, dist = _init(rank, world, port)
    from xp_physics_cuda import scenes
    from xp_physics_cuda.distributed import ContactExchange, ShardedSweep, shard_capacity, shard_range
    tris, h = scenes.sine_terrain(120, 120, seed=5)
    sizes = (60001, 60001, 45000)
    batches = [scenes.terrain_spheres(h, n, seed=6 + i) for i, n in enumerate(sizes)]
    sw = ShardedSweep(tris, rank, exchange="nccl")
    want = [sw.sweep(b).clone() for b in batches]                      # NCCL all-gather: the reference transport
    dev = sw.device
    ex = ContactExchange(sw.scene, shard_capacity(max(sizes), world), dev)
    ok = ex.ok and ex.mode == "p2p"
    bad_steps = []
    if ok:
        d_in = []
        for b in batches:
            lo, hi = shard_range(len(b), rank, world)
            d_in.append((torch.from_numpy(np.ascontiguousarray(b[lo:hi]).view(np.float32).reshape(-1, 7)).to(dev), hi - lo, len(b)))
        pending = []

        def finish():
            k, t = pending.pop(0)
            buf = ex.wait(t)
            n = d_in[k % 3][2]
            got = torch.cat([buf[r, :shard_range(n, r, world)[1] - shard_range(n, r, world)[0]] for r in range(world)])
            if not torch.equal(got.view(torch.int32), want[k % 3].view(torch.int32)):
                bad_steps.append(k)

        stream = torch.cuda.current_stream(dev)
        for k in range(steps):
            if len(pending) == 2:
                finish()
            if (k * 5 + rank * 3) % 7 == 0:
                time.sleep(0.003 * ((k + rank) % 4))                   # host-side skew
            if (k + 2 * rank) % 5 == 0:
                torch.cuda._sleep(int(2e6) * (1 + (rank + k) % 3))     # device-side skew (~1-3 ms) before the sweep
            t, n_loc, _ = d_in[k % 3]
            pending.append((k, ex.sweep_async(t.data_ptr(), n_loc)))
        while pending:
            finish()
        stream.synchronize()
    with open(os.path.join(out_dir, f"loop{rank}.txt"), "w") as f:
        f.write(f"{int(ok)} {ex.mode} {len(bad_steps)} {bad_steps[:8]}")
    ex.close()
    sw.close()
    dist.destroy_process_group()


def test_exchange_loop_every_rank_every_step(tmp_path):
    import torch.multiprocessing as mp
    world = _world()
    steps = 40
    mp.spawn(_loop_worker, args=(world, _free_port(), str(tmp_path), steps), nprocs=world, join=True)
    for r in range(world):
        ok, mode, n_bad, which = open(tmp_path / f"loop{r}.txt").read().split(" ", 3)
        assert ok == "1", f"rank {r}: peer-memory exchange unavailable ({mode})"
        assert n_bad == "0", f"rank {r}: gather buffer differs from the NCCL all-gather on steps {which}"
    print(f"world {world}: {steps} asynchronous steps, every rank's gather buffer identical to the NCCL all-gather")
