"""torchrun --nproc-per-node N tools/check_exchange.py : the fused exchange (multicast if the box offers it,
else P2P) against the NCCL all-gather of the same sweep, record for record."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "xp-game-engine_b200"))
from xp_physics_cuda import scenes  # noqa: E402
from xp_physics_cuda.distributed import ShardedSweep  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
tris, h = scenes.sine_terrain(200, 200, seed=5)
resps = scenes.terrain_spheres(h, 200001, seed=6)
out = {}
for mode in ("nccl", "auto"):
    sw = ShardedSweep(tris, int(os.environ["LOCAL_RANK"]), exchange=mode)
    for rep in range(3):
        got = sw.sweep(resps)
    out[mode] = got.clone()
    used = sw.last_exchange
    note = getattr(sw._ex, "multicast_note", "") if sw._ex is not None else ""
    sw.close()
    if rank == 0:
        print(f"exchange={mode}: used {used} {note}", flush=True)
# the asynchronous loop: three different batches, two steps in flight, each against the NCCL result
from xp_physics_cuda.distributed import ContactExchange, shard_capacity, shard_range  # noqa: E402
batches = [resps, scenes.terrain_spheres(h, 200001, seed=7), scenes.terrain_spheres(h, 150000, seed=8)]
sw = ShardedSweep(tris, int(os.environ["LOCAL_RANK"]), exchange="nccl")
want = [sw.sweep(b).clone() for b in batches]
dev = sw.device
ex = ContactExchange(sw.scene, shard_capacity(200001, world), dev)
d_in = []
for b in batches:
    lo, hi = shard_range(len(b), rank, world)
    d_in.append((torch.from_numpy(np.ascontiguousarray(b[lo:hi]).view(np.float32).reshape(-1, 7)).to(dev), hi - lo, len(b)))
async_ok = True
pending = []
def finish():
    global async_ok
    k, t = pending.pop(0)
    buf = ex.wait(t)
    n = d_in[k % 3][2]
    got = torch.cat([buf[r, :shard_range(n, r, world)[1] - shard_range(n, r, world)[0]] for r in range(world)])
    async_ok &= torch.equal(got.view(torch.int32), want[k % 3].view(torch.int32))
for k in range(9):
    if len(pending) == 2:
        finish()
    t, n_loc, _ = d_in[k % 3]
    pending.append((k, ex.sweep_async(t.data_ptr(), n_loc)))
while pending:
    finish()
print(f"rank {rank}: async loop identical={async_ok} mode={ex.mode}", flush=True)
ex.close()
sw.close()
same = async_ok and torch.equal(out["nccl"].view(torch.int32), out["auto"].view(torch.int32))
hits = int((out["auto"][:, 9].contiguous().view(torch.int32) & 1).sum().item())
print(f"rank {rank}: identical={same} hits={hits} of {len(resps)}", flush=True)
dist.destroy_process_group()
sys.exit(0 if same else 1)
