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
same = torch.equal(out["nccl"].view(torch.int32), out["auto"].view(torch.int32))
hits = int((out["auto"][:, 9].contiguous().view(torch.int32) & 1).sum().item())
print(f"rank {rank}: identical={same} hits={hits} of {len(resps)}", flush=True)
dist.destroy_process_group()
sys.exit(0 if same else 1)
