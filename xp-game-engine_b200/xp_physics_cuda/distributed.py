"""Multi-GPU sweep: one process per GPU, spheres sharded, scene replicated, one all-gather.

SURVEY.md section 8(e): spheres are independent (collision_response reads only its own Response
and the read-only triangle set, xp_physics/src/lib.rs:29-62), so the batch is cut into contiguous
blocks, one per rank; every rank uploads the same triangles and builds the same LBVH; the only
exchange step is the all-gather of fixed-size 40 B contact records (xp_contact).  Each rank's sweep
writes its records straight into its slot of the gather buffer (xp_detect_contacts_dev), so there
is no packing pass between the kernel and the collective.

torch.distributed is plumbing only (process group + NCCL/gloo all_gather).
"""
from __future__ import annotations

import os

import numpy as np

from .api import CONTACT16_DTYPE, CONTACT_DTYPE, RESPONSE_DTYPE, Scene

CONTACT_FLOATS = 10      # xp_contact, 40 B
CONTACT16_FLOATS = 4     # xp_contact16 (Scene.set_options(contact_record="compact")), 16 B: (position, tri_index)


def shard_range(n: int, rank: int, world: int):
    """Contiguous block [lo, hi) of rank `rank`; blocks differ by at most one sphere."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_capacity(n: int, world: int) -> int:
    """Slot size of the gather buffer (the largest block)."""
    return (n + world - 1) // world


def gather_contacts(local, n_total: int, group=None):
    """All-gather per-rank contact blocks into the global list.

    local: torch tensor [capacity, 10] float32 (this rank's slot content, rows beyond its block are
    padding).  Returns a [n_total, 10] tensor holding every sphere's record in input order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    cap = shard_capacity(n_total, world)
    width = local.shape[1]
    assert local.shape[0] == cap and width in (CONTACT_FLOATS, CONTACT16_FLOATS), (local.shape, cap)
    buf = torch.empty((world, cap, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf.view(-1), local.reshape(-1).contiguous(), group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(buf[r, :hi - lo])
    return torch.cat(parts, dim=0)


def contacts_as_records(t) -> np.ndarray:
    """[n,10] (or [n,4]) float32 tensor -> structured CONTACT_DTYPE (CONTACT16_DTYPE) array (host)."""
    a = t.detach().cpu().numpy()
    return np.ascontiguousarray(a).view(CONTACT16_DTYPE if a.shape[1] == CONTACT16_FLOATS else CONTACT_DTYPE).reshape(-1)


class _DeviceArray:
    """Minimal __cuda_array_interface__ wrapper so torch can view library-allocated device memory."""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


class ContactExchange:
    """N_BUF gather buffers per rank (each [world, cap, 10] float32) that a kernel on any rank can store into;
    synchronous steps use the first, a loop of steps (sweep_async) goes through them round robin.

    Two transports (`mode` says which one is live):
      "p2p":       every rank maps every other rank's buffer through CUDA IPC and the kernel stores each
                   span into all of them (NVLink P2P).  The default.
      "multicast": XP_EXCHANGE_MULTICAST=1 -- the buffers are torch symmetric memory bound to one NVSwitch
                   multicast address; the record-building kernel stores every span ONCE (multimem.st) and
                   the switch replicates it into all ranks' buffers: 1/world of the NVLink egress.
    An all-gather is bound by what every GPU must RECEIVE (7/8 of the list), so the two measure the same
    on 8 B200s (exchange stage 0.43 vs 0.44 ms for 336 MB of records, step 2.51 vs 2.56 ms).
    `ok` is False when the transport could not be set up on every rank (callers fall back to NCCL)."""

    N_BUF = 4

    def __init__(self, scene: Scene, cap: int, device, group=None, n_buf: int = 0, record: str = "full"):
        import torch
        import torch.distributed as dist
        self.scene, self.group, self.device = scene, group, device
        # record = "compact": 16 B (position, triangle) instead of the 40 B xp_contact -- 2.5x fewer bytes over NVLink
        self.record = record
        self.width = CONTACT16_FLOATS if record == "compact" else CONTACT_FLOATS
        scene.set_options(contact_record=record)
        if n_buf:
            self.N_BUF = n_buf                      # 1 is enough for synchronous steps (sweep); sweep_async needs 4
        self.solo = not (dist.is_available() and dist.is_initialized())        # one process, no peers
        self.world = 1 if self.solo else dist.get_world_size(group)
        self.rank = 0 if self.solo else dist.get_rank(group)
        self.cap = cap + (cap & 1)                   # even slots: 16-byte aligned spans
        self.half = self.world * self.cap            # records per gather buffer; there are N_BUF (steps use them round robin)
        self.nbytes = self.N_BUF * self.half * 4 * self.width
        self.ptr, self.bufs, self.ok, self.mode = 0, [], True, "p2p"
        self.mc_ptr, self._symm = 0, None
        self._side, self._done, self._flag = None, {}, None
        shape = (self.N_BUF, self.world, self.cap, self.width)
        if (not self.solo and os.environ.get("XP_EXCHANGE_MULTICAST", "0") == "1"
                and self._try_multicast(torch, dist, device, shape)):
            self.mode = "multicast"
            return
        handle = b""
        try:
            self.ptr, handle = scene.exchange_alloc(self.nbytes)
        except Exception:
            self.ok = False
        if self.solo:
            if self.ok:
                self.bufs = [self.ptr]
                self.views = torch.as_tensor(_DeviceArray(self.ptr, shape), device=device)
                self.view = self.views[0]
            return
        handles = [None] * self.world
        dist.all_gather_object(handles, (self.ok, handle), group=group)
        if all(h[0] for h in handles):
            try:
                for r, (_, h) in enumerate(handles):
                    self.bufs.append(self.ptr if r == self.rank else scene.exchange_open(h))
            except Exception:
                self.ok = False
        else:
            self.ok = False
        flags = [None] * self.world
        dist.all_gather_object(flags, self.ok, group=group)
        if not all(flags):
            self.ok = False
            self.close()
        else:
            self.views = torch.as_tensor(_DeviceArray(self.ptr, shape), device=device)
            self.view = self.views[0]

    def _try_multicast(self, torch, dist, device, shape) -> bool:
        """Symmetric memory + multicast address through torch (plumbing only: allocation and rendezvous)."""
        mc, t, why = 0, None, ""
        try:
            import torch.distributed._symmetric_memory as symm_mem
            g = self.group if self.group is not None else dist.group.WORLD
            t = symm_mem.empty(shape, dtype=torch.float32, device=device)
            hdl = symm_mem.rendezvous(t, g)
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            if mc & 15:
                mc = 0
        except Exception as e:                       # noqa: BLE001 -- any failure means "not available here"
            mc, why = 0, repr(e)[:200]
        flags = [None] * self.world
        dist.all_gather_object(flags, mc != 0, group=self.group)
        self.multicast_note = why
        if not all(flags):
            return False
        t.zero_()
        self._symm, self.views, self.view, self.mc_ptr = t, t, t[0], mc
        torch.cuda.synchronize(device)
        dist.barrier(self.group)
        return True

    def sweep(self, d_in_ptr: int, n_local: int, horizon: str = "inf"):
        """Sweep this rank's spheres; their contact records land in slot `rank` of EVERY rank's buffer
        (`view`, the first of the gather buffers).  Synchronous; the caller adds the barrier."""
        if self.mode == "multicast":
            self.scene.detect_contacts_multicast_dev(d_in_ptr, n_local, self.mc_ptr, self.rank * self.cap, horizon)
        else:
            self.scene.detect_contacts_exchange_dev(d_in_ptr, n_local, self.bufs, self.rank * self.cap, horizon)

    # -- a loop of steps: step i's exchange runs under step i+1's sweep -----------------------------
    def sweep_async(self, d_in_ptr: int, n_local: int, horizon: str = "inf") -> int:
        """Queue one step: the sweep on the scene's stream, the record build + exchange on a second stream,
        the cross-rank barrier (a one-element NCCL all-reduce) behind it on that stream.  Steps go round robin
        through the gather buffers; `wait(ticket)` returns the buffer of that step."""
        import torch
        import torch.distributed as dist
        if self.mode != "p2p":
            raise RuntimeError("sweep_async uses the P2P transport")
        if self.N_BUF < 4:
            raise RuntimeError("sweep_async needs four gather buffers (see below)")
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
            self.scene.set_exchange_stream(self._side.cuda_stream)
            self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        # Four gather buffers, used round robin, make an explicit "buffer free" handshake unnecessary: step i+4
        # overwrites the buffer of step i on every rank, and a rank starts its exchange of step i+4 only after its
        # wait(i+2), i.e. after the barrier behind every rank's exchange of step i+2 -- which every rank queued after
        # the reads of buffer i it had put on its stream (wait(i) precedes submit(i+2): two steps in flight).  With
        # two buffers a fast rank's step i+2 could overwrite a buffer a slow rank was still reading (seen by
        # tools/check_exchange.py); an extra release all-reduce per step fixed that too but cost 0.25 ms on 8 GPUs.
        half = self._half_i = (getattr(self, "_half_i", self.N_BUF - 1) + 1) % self.N_BUF
        ticket = self.scene.detect_contacts_exchange_submit(
            d_in_ptr, n_local, self.bufs, half * self.half + self.rank * self.cap, horizon,
            self_index=self.rank if os.environ.get("XP_EXCHANGE_COPY", "1") != "0" else -1)
        if not self.solo:
            with torch.cuda.stream(self._side):
                dist.all_reduce(self._flag, group=self.group)        # every rank's exchange kernel has finished
        ev = torch.cuda.Event()
        ev.record(self._side)
        self._done[ticket] = (ev, half)
        return ticket

    def wait(self, ticket: int):
        """Complete a queued step; returns its gather buffer [world, cap, 10].  It stays valid until this rank
        queues the step after the next one: reads queued on the scene's stream (or done synchronously) before that
        call are safe on every rank (four buffers, see sweep_async)."""
        ev, half = self._done.pop(ticket)
        ok = self.scene.detect_contacts_exchange_wait(ticket)
        ev.synchronize()
        if not ok:
            raise RuntimeError("candidate buffer overflow in an asynchronous step: raise max_pairs or use sweep()")
        return self.views[half]

    def close(self):
        for r, b in enumerate(self.bufs):
            if r != self.rank and b:
                self.scene.exchange_close(b)
        self.bufs = []
        if self.ptr:
            self.scene.exchange_free(self.ptr)
            self.ptr = 0
        self._symm = None


class ShardedSweep:
    """Rank-local half of the multi-GPU sweep (needs a CUDA device per rank).

    exchange = "p2p":  every rank maps every other rank's gather buffer (CUDA IPC over NVLink) and the
                       kernel that builds the contact records stores them into all buffers -- compute and
                       all-gather are one kernel; the ranks only meet at a barrier.
    exchange = "nccl": records are written to the local slot and all_gather_into_tensor moves them.
    exchange = "auto": p2p when the peer mapping succeeds on all ranks, else nccl.
    """

    def __init__(self, triangles, device: int, exchange: str = "auto", record: str = "full", **scene_opts):
        import torch
        self.record = record
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.scene = Scene(device, **scene_opts)
        self.scene.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self.scene.set_triangles(triangles)          # replicated: every rank builds the same tree
        self.scene.build()
        self.exchange = exchange
        self._ex = None

    def _setup_p2p(self, n_total: int, group=None) -> bool:
        import torch.distributed as dist
        cap = shard_capacity(n_total, dist.get_world_size(group))
        if self._ex is not None and self._ex.ok and self._ex.cap >= cap:
            return True
        if self._ex is not None:
            self._ex.close()
        self._ex = ContactExchange(self.scene, cap, self.device, group, record=self.record)
        return self._ex.ok

    def sweep(self, responses: np.ndarray, group=None):
        """responses: the FULL batch (identical on every rank).  Returns the global contact list
        [n,10] on this rank's GPU."""
        import torch.distributed as dist
        torch = self.torch
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n = len(responses)
        lo, hi = shard_range(n, rank, world)
        mine = np.ascontiguousarray(responses[lo:hi], RESPONSE_DTYPE)
        d_in = torch.from_numpy(mine.view(np.float32).reshape(-1, 7)).to(self.device)
        use_p2p = self.exchange in ("auto", "p2p") and self._setup_p2p(n, group)
        if self.exchange == "p2p" and not use_p2p:
            raise RuntimeError("peer-memory exchange requested but CUDA IPC mapping failed")
        self.last_exchange = self._ex.mode if use_p2p else "nccl"
        if use_p2p:
            if hi > lo:
                self._ex.sweep(d_in.data_ptr(), hi - lo)
            torch.cuda.synchronize(self.device)
            dist.barrier(group)                       # every rank's stores have landed
            buf = self._ex.view
            parts = [buf[r, :shard_range(n, r, world)[1] - shard_range(n, r, world)[0]] for r in range(world)]
            out = torch.cat(parts, dim=0)
            dist.barrier(group)                       # nobody overwrites a buffer that is still being read
            return out
        cap = shard_capacity(n, world)
        self.scene.set_options(contact_record=self.record)
        slot = torch.zeros((cap, CONTACT16_FLOATS if self.record == "compact" else CONTACT_FLOATS), dtype=torch.float32, device=self.device)
        if hi > lo:
            self.scene.detect_contacts_dev(d_in.data_ptr(), hi - lo, slot.data_ptr())
        return gather_contacts(slot, n, group)

    def close(self):
        if self._ex is not None:
            self._ex.close()
        self.scene.close()
