#!/usr/bin/env python
"""bench.py -- swept-sphere vs triangle narrowphase tests/sec and spheres resolved/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 1..5]
                    [--arith exact|fast|...] [--scene heightmap|soup] [--no-cpu] [--no-extras]

Headline workload (BASELINE.json configs[1], "config 2"): 2^20 swept unit spheres vs a 1 001 112-triangle
synthetic heightmap terrain, single sweep, per GPU.  A "step" is one closest-collision sweep of the whole sphere
batch against the resident scene (the terrain is uploaded / indexed once, outside the timed region, exactly like
the reference's `&[Triangle]` argument exists before collision_response is called).

N > 1 is launched by torchrun (one rank per GPU): the sphere batch is sharded (weak scaling: 2^20 spheres per GPU,
scene replicated) and the global list of 40 B contact records is assembled in every rank's gather buffer over NVLink
peer memory by the copy engines (NCCL all_gather as the fallback); `exchange_verified` says whether every rank ended
up with every rank's records.

Prints ONE JSON line (rank 0):
  value        narrowphase tests/s with inputs resident in HBM, over a loop of steps with two in flight
  e2e          the same metric through the host-pointer C ABI with pinned HOST buffers, every step's H2D and D2H
               inside the timed region (compact results: 28 B in, 8 B out per sphere), next to the box's measured
               host-link ceiling for exactly those bytes
  roofline     the dominant kernel (narrowphase) against the fp32 FFMA peak measured in this run
  resolved     spheres resolved per second by the full collision_response loop (4 iterations), device-resident
  arith        exact (bit-identical, the default) and fast (tolerance mode) side by side; far_exact timing
  sustained    the same loop of steps run for >= 2 s with the clocks seen
  configs      BASELINE.json's other named configs (1, 3, 4, 5) on this run's GPUs, bounded to a few steps each
  cpu_baseline the CPU oracle (a port of the reference's algorithm: brute force over every triangle) on this
               box's host cores, bounded sample

`--config K` makes config K the line's headline instead (same JSON shape; config 2 is the default the driver runs).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "xp-game-engine_b200")]

NX, NZ = 708, 707                 # 2 * 708 * 707 = 1 001 112 triangles
N_SPHERES = 1 << 20
# algorithmic flops per narrowphase test by code path (full 245, full+PIT 298),
# minus the 14 flops of the two squared edge lengths that are stored with the triangle at upload (like n and
# the plane constant, they depend on the triangle only) and no longer computed per test
PATH_FLOPS = {"cull": 5, "far": 16, "face_hit": 86, "full": 231, "full_pit": 284}
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # 74.4


def build_workload(seed: int, n_spheres: int = N_SPHERES, with_heights: bool = False):
    from xp_physics_cuda import scenes
    tris, heights = scenes.sine_terrain(NX, NZ, seed=1)
    resps = scenes.terrain_spheres(heights, n_spheres, seed=seed)
    return (tris, resps, heights) if with_heights else (tris, resps)


def workload_config(n_gpus: int) -> dict:
    """The `config` of the JSON line: the SAME dict in both arms (the reference arm runs on this arm's config), so it
    names the workload only; how the B200 arm ran it (arithmetic mode, method, exchange transport ...) is the `arm` key."""
    return {"workload": "config2: 2^20 swept unit spheres x 1001112-triangle sine+noise heightmap terrain, "
                        "single sweep per step, horizon=inf (reference-faithful)",
            "spheres_per_gpu": N_SPHERES, "triangles": 2 * NX * NZ,
            "parallelism": f"spheres sharded over {n_gpus} GPU(s), scene replicated",
            "l2": ("B200 arm: inputs larger than L2 -- one step's inputs (scene records + 29 MB of spheres) exceed the 126 MB "
                   "L2 and a step streams ~0.5 GB through it; no flush inside the loop of steps (they overlap); "
                   "single_step is the same step with an explicit 256 MiB L2 flush between steps.  Reference arm: CPU, n/a")}


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed regions; rank 0 only)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True, period: float = 0.1):
        self.index, self.rows, self.stop, self.t, self.enabled, self.period = index, [], threading.Event(), None, enabled, period

    def __enter__(self):
        if not self.enabled:
            return self

        def run():
            while not self.stop.is_set():
                try:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    self.rows.append([x.strip() for x in out.strip().split(",")])
                except Exception:
                    pass
                self.stop.wait(self.period)
        self.t = threading.Thread(target=run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        if self.t is not None:
            self.stop.set()
            self.t.join(timeout=6)

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU legs (the only places the oracle is executed by bench.py)
# ---------------------------------------------------------------------------------------------
def cpu_sample(tris, resps, n_sample: int, threads: int = 0):
    """Oracle brute force (the reference's algorithm, lib.rs:33-35) on the first n_sample spheres."""
    from oracle import xp_oracle as o
    t0 = time.perf_counter()
    out = o.detect_batch(resps[:n_sample], tris, threads=threads)
    dt = time.perf_counter() - t0
    return out[4] / dt, dt, (threads or o.max_threads())


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (brute force over all triangles per
    sphere), all host threads, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import xp_oracle as o
    tris, resps = build_workload(seed=1, n_spheres=1 << 14)
    cores = o.max_threads()
    n_sample = 4 * cores                      # 4 spheres x 1M triangles per core per step: ~0.5 s
    for _ in range(args.warmup):
        o.detect_batch(resps[:n_sample], tris)
    t0 = time.perf_counter()
    tests = 0
    for k in range(args.steps):
        lo = (k * n_sample) % (len(resps) - n_sample)
        tests += o.detect_batch(resps[lo:lo + n_sample], tris)[4]
    dt = time.perf_counter() - t0
    v = tests / dt
    sample = f"{n_sample} spheres x {len(tris)} triangles brute force per step (all {cores} host threads)"
    line = {"impl": "reference", "metric": "narrowphase_tests_per_sec", "value": v, "unit": "tests/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "spheres_per_sec": n_sample * args.steps / dt,
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": v, "unit": "tests/s", "cores": cores, "kind": "port", "sample": sample,
                             "note": "C oracle = op-for-op port of xp_physics (Rust toolchain absent); no per-test "
                                     "heap allocations or println!, so faster than the Rust original"},
            "e2e": {"value": v, "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# helpers of the B200 arm
# ---------------------------------------------------------------------------------------------
class Ctx:
    """torch / distributed plumbing shared by the headline run and the config blocks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- xp_physics_cuda has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, *xs):
        if self.world == 1:
            return [float(x) for x in xs]
        t = self.torch.tensor(list(xs), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    def timed_region(self, fn):
        """fn() once between barrier + synchronize on both sides, CUDA events on the launching stream; max over ranks."""
        torch = self.torch
        self.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(self.stream)
        out = fn()
        b.record(self.stream)
        torch.cuda.synchronize()
        self.barrier()
        return self.max_over_ranks(a.elapsed_time(b)), out

    def dev_resps(self, r: np.ndarray):
        return self.torch.from_numpy(r.view(np.float32).reshape(-1, 7)).to(self.dev)


def dev_terrain_spheres(ctx: Ctx, heights_dev, n: int, seed: int):
    """scenes.terrain_spheres generated on the device (same distribution; big configs only -- 16M spheres take
    seconds in numpy): c.xz ~ U(grid), c.y = max corner height + U(1.05, 6), m = (U(-.3,.3), U(-.6,-.05), U(-.3,.3))."""
    torch = ctx.torch
    g = torch.Generator(device=ctx.dev)
    g.manual_seed(seed)
    nx, nz = heights_dev.shape[0] - 1, heights_dev.shape[1] - 1
    u = torch.rand((n, 6), generator=g, device=ctx.dev, dtype=torch.float32)
    x, z = u[:, 0] * nx, u[:, 1] * nz
    hx, hz = x.long().clamp_(max=nx - 1), z.long().clamp_(max=nz - 1)
    hmax = torch.maximum(torch.maximum(heights_dev[hx, hz], heights_dev[hx + 1, hz]),
                         torch.maximum(heights_dev[hx, hz + 1], heights_dev[hx + 1, hz + 1]))
    out = torch.empty((n, 7), dtype=torch.float32, device=ctx.dev)
    out[:, 0], out[:, 2] = x, z
    out[:, 1] = hmax + 1.05 + u[:, 2] * (6.0 - 1.05)
    out[:, 3] = 1.0
    out[:, 4] = -0.3 + 0.6 * u[:, 3]
    out[:, 5] = -0.6 + 0.55 * u[:, 4]
    out[:, 6] = -0.3 + 0.6 * u[:, 5]
    return out


def pcie_ceiling(ctx: Ctx, h2d_bytes: int, d2h_bytes: int, steps: int = 20) -> dict:
    """The box's host-link ceiling for one step's bytes: plain pinned copies of exactly h2d_bytes up and d2h_bytes
    down per step, the two directions on their own streams, every rank at once (max over ranks)."""
    torch = ctx.torch
    up_h = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    dn_h = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    up_d = torch.empty(h2d_bytes, dtype=torch.uint8, device=ctx.dev)
    dn_d = torch.empty(d2h_bytes, dtype=torch.uint8, device=ctx.dev)
    s_up, s_dn = torch.cuda.Stream(device=ctx.dev), torch.cuda.Stream(device=ctx.dev)

    def run(do_up, do_dn):
        ev = torch.cuda.Event()
        ev.record(ctx.stream)
        s_up.wait_event(ev); s_dn.wait_event(ev)
        for _ in range(steps):
            if do_up:
                with torch.cuda.stream(s_up):
                    up_d.copy_(up_h, non_blocking=True)
            if do_dn:
                with torch.cuda.stream(s_dn):
                    dn_h.copy_(dn_d, non_blocking=True)
        ctx.stream.wait_stream(s_up); ctx.stream.wait_stream(s_dn)

    out = {}
    for name, u, d in (("both", True, True), ("h2d_only", True, False), ("d2h_only", False, True)):
        run(u, d)                                                   # warm-up
        ms, _ = ctx.timed_region(lambda: run(u, d))
        out[name + "_ms_per_step"] = ms / steps
        moved = (h2d_bytes if u else 0) + (d2h_bytes if d else 0)
        out[name + "_gbs_per_gpu"] = moved / (ms / steps * 1e-3) / 1e9
    out["aggregate_gbs_all_gpus"] = out["both_gbs_per_gpu"] * ctx.world
    out["note"] = ("pinned cudaMemcpyAsync loops of one step's bytes per direction, both directions concurrently, all "
                   f"{ctx.world} rank(s) at once; max over ranks")
    return out


def measured_hbm_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# config 2: the headline
# ---------------------------------------------------------------------------------------------
def run_config2(ctx: Ctx, args) -> dict:
    import xp_physics_cuda as xp
    from xp_physics_cuda.distributed import ContactExchange
    torch, dist, world, rank, dev, stream = ctx.torch, ctx.dist, ctx.world, ctx.rank, ctx.dev, ctx.stream

    tris, resps, heights = build_workload(seed=1 + rank, with_heights=True)
    n = len(resps)
    scene = xp.Scene(ctx.local, arith=args.arith, method=args.method, prune=args.prune, timing=True)
    if args.far_exact:
        scene.set_options(far_exact=True)
    scene.set_stream(stream.cuda_stream)
    if args.scene == "heightmap":
        scene.set_heightmap(heights)              # the terrain as the reference holds it: a height grid (triangles emitted on the device)
    else:
        scene.set_triangles(tris)                 # the same terrain as an anonymous triangle soup: LBVH walk
    scene.build()
    build_stats = scene.stats()

    d_in = ctx.dev_resps(resps)
    contacts = torch.empty((world, n, 10), dtype=torch.float32, device=dev)   # 40 B records (NCCL fallback buffer)
    mine = contacts[rank]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    exchange = None
    if not os.environ.get("XP_BENCH_NCCL_EXCHANGE"):
        exchange = ContactExchange(scene, n, dev)
        if not exchange.ok:
            exchange = None
        else:
            mine = exchange.view[rank][:n]
    pipelined = exchange is not None and exchange.mode == "p2p" and not os.environ.get("XP_BENCH_SYNC_STEPS")

    def step_device():
        if exchange is not None:
            exchange.sweep(d_in.data_ptr(), n)       # synchronous step: the record-building kernel stores into every buffer
            if world > 1:
                dist.barrier()
            return
        scene.detect_contacts_dev(d_in.data_ptr(), n, mine.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(contacts.view(-1), mine.reshape(-1))

    def timed(fn, steps, warmup, flush_l2=True):
        for _ in range(warmup):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        stats = []
        ctx.barrier()
        torch.cuda.synchronize()
        for a, b in ev:
            if flush_l2:
                flush.fill_(1)                 # L2 flush, outside the event pair
            a.record(stream)
            fn()
            b.record(stream)
            stats.append(scene.stats())
        torch.cuda.synchronize()
        ctx.barrier()
        return ctx.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev)), stats

    last_view = [None]

    def loop_device(k, ex=None):
        """k steps as a loop: submit step i, then complete step i-1 -- step i's sweep runs while step i-1's records
        cross NVLink (copy engines, second stream) and its barrier completes.  Returns the per-step stats."""
        ex = ex or exchange
        pending, out = [], []
        for _ in range(k):
            if len(pending) == 2:
                last_view[0] = ex.wait(pending.pop(0))
                out.append(scene.stats())
            pending.append(ex.sweep_async(d_in.data_ptr(), n))
        while pending:
            last_view[0] = ex.wait(pending.pop(0))
            out.append(scene.stats())
        return out

    def timed_loop(steps, warmup, ex=None):
        loop_device(warmup, ex)
        ms, stats = ctx.timed_region(lambda: loop_device(steps, ex))
        return ms, stats

    W = max(args.warmup, 3)
    with ClockSampler(ctx.local, enabled=rank == 0) as clocks:      # nvidia-smi samples while both timed regions run
        sync_ms, sync_stats = timed(step_device, args.steps, W)     # one synchronous call per step, L2 flushed
        if pipelined:
            total_ms, stats = timed_loop(args.steps, W)
        else:
            total_ms, stats = sync_ms, sync_stats

    # ---- N > 1: does every rank hold every rank's records?  (a) per-slot checksums of this rank's gather buffer against
    # the checksum each owner computes over records it builds again, locally and synchronously; (b) all ranks agree.
    exchange_verified = None
    if world > 1:
        view = last_view[0] if pipelined else (exchange.view if exchange is not None else contacts)

        def csum(t):
            return int(t.contiguous().view(torch.int32).to(torch.int64).sum().item())
        own = torch.empty((n, 10), dtype=torch.float32, device=dev)
        scene.detect_contacts_dev(d_in.data_ptr(), n, own.data_ptr())
        torch.cuda.synchronize()
        mine_sum = torch.tensor([csum(own)], dtype=torch.int64, device=dev)
        all_own = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(all_own, mine_sum)
        slots_ok = all(csum(view[r][:n]) == int(all_own[r].item()) for r in range(world))
        same_as_local = bool(torch.equal(view[rank][:n].view(torch.int32), own.view(torch.int32)))
        flag = torch.tensor([1 if (slots_ok and same_as_local) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        exchange_verified = bool(flag.item() == 1)

    last_records = (last_view[0][rank][:n] if pipelined else mine).clone()
    # ---- N > 1: the same loop shipping 16 B records (XP_OPT_CONTACT_RECORD = 1: centre at the contact + triangle) instead
    # of the 40 B xp_contact: what the exchange's bytes cost the sweep they run under
    exchange_compact = None
    if world > 1 and pipelined and not args.no_extras:
        ex16 = ContactExchange(scene, n, dev, record="compact")
        if ex16.ok:
            keep = last_view[0]
            ms16, st16 = timed_loop(args.steps, W, ex16)
            (t16,) = ctx.sum_over_ranks(sum(x["narrowphase_tests"] for x in st16))
            exchange_compact = {"record": "xp_contact16: (position, tri_index), 16 B instead of 40 B per sphere",
                                "ms_per_step": ms16 / args.steps, "value": t16 / (ms16 * 1e-3), "unit": "tests/s"}
            last_view[0] = keep
        ex16.close()
        scene.set_options(contact_record="full")

    tests_local = sum(s["narrowphase_tests"] for s in stats)
    launches = sum(s["kernel_launches"] for s in stats)
    (tests_all,) = ctx.sum_over_ranks(tests_local)
    sec = total_ms * 1e-3
    value = tests_all / sec
    spheres_per_sec = world * n * args.steps / sec

    # ---- rooflines, from the CUDA-event stage times the library records on its own stream around each kernel
    st_ms = {k: float(np.mean([s["stage_ms"][k] for s in stats])) for k in stats[0]["stage_ms"]}
    two_stage = args.method in ("lbvh_pairs", "wide_pairs", "q4_pairs")
    # path mix of the narrowphase tests: one untimed diagnostic sweep with the fused kernel, which counts the code
    # path of every test (the two-stage kernels carry no counters)
    scene.set_options(method="lbvh_fused", prune=False)
    scene.detect_contacts_dev(d_in.data_ptr(), n, mine.data_ptr())
    paths = {k: float(v) for k, v in scene.stats()["path_tests"].items()}
    scene.set_options(method=args.method, prune=args.prune)
    flops = sum(paths[k] * PATH_FLOPS[k] for k in PATH_FLOPS)
    try:
        peak = xp.probe_fp32_peak(ctx.local)
        peak_src = "measured: FFMA micro-benchmark xp_probe_fp32_peak in this run (nominal 74.4; MEASURED_PEAKS.json has no fp32 figure)"
    except Exception:
        peak, peak_src = NOMINAL_FP32_TFLOPS, "nominal 148 SM x 128 lanes x 2 x 1.965 GHz"
    np_ms = st_ms["narrowphase"] if two_stage else st_ms["traverse"] + st_ms["narrowphase"]
    achieved = flops / (np_ms * 1e-3) / 1e12 if np_ms > 0 else 0.0
    kname = {"wide_fused": "k_sweep", "wide_pairs": "k_narrow_pairs", "lbvh_fused": "k_traverse",
             "lbvh_pairs": "k_narrow_pairs", "q4_pairs": "k_narrow_pairs<FILTER>", "brute": "k_brute"}[args.method]
    grid_walk = args.scene == "heightmap" and args.method == "lbvh_pairs"
    roofline = {"bound": "fp32", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak else None, "peak_source": peak_src, "kernel_ms": np_ms,
                "algorithmic_flops_per_launch": flops, "flops_per_test_by_path": PATH_FLOPS, "path_mix": paths,
                "traffic": TRAFFIC_NARROW, "traffic_source": TRAFFIC_SOURCE,
                "note": "231 algorithmic flops per full test; the exact kernel issues ~423 thread instructions per test "
                        "(round 1: 577), the fast one ~230; ncu (profiles/r02zz_ncu_full.txt): issue slots 85% busy, "
                        "FMA pipe 54%, 30 of 32 lanes active; exact arithmetic may not contract (the reference does not), so "
                        "one flop costs one issue slot where the FFMA peak counts two"}
    tests_step = float(np.mean([s["narrowphase_tests"] for s in stats]))
    hbm_peak, hbm_src = measured_hbm_gbs()
    bp_bytes = 64.0 * n + 8.0 * tests_step                       # ray records in + candidate pairs out
    bp_ms = st_ms["traverse"]
    roofline_hbm = {"bound": "hbm", "kernel": "k_broadphase_grid" if grid_walk else {"lbvh_pairs": "k_broadphase", "q4_pairs": "k_broadphase4"}.get(args.method, "k_sweep<EMIT>"),
                    "achieved": bp_bytes / (bp_ms * 1e-3) / 1e9 if (two_stage and bp_ms > 0) else None,
                    "peak": hbm_peak, "peak_source": hbm_src, "unit": "GB/s", "kernel_ms": bp_ms,
                    "frac": (bp_bytes / (bp_ms * 1e-3) / 1e9) / hbm_peak if (two_stage and bp_ms > 0) else None,
                    "algorithmic_bytes_per_launch": bp_bytes, "traffic": TRAFFIC_BROADPHASE, "traffic_source": TRAFFIC_SOURCE,
                    "note": "64 B ray record per sphere + 8 B per candidate pair; the walk itself reads L1/L2-resident heights "
                            "(grid) or nodes (LBVH) and is bound by instruction issue (grid: 29 of 32 lanes active, issue slots 80% "
                            "busy, DRAM 11%) or by L1 wavefronts (LBVH), not by HBM"}

    # ---- e2e: host pointers, pinned buffers, every step's H2D + D2H inside the timed region, two batches in flight
    L = xp._lib.lib()
    h_in = torch.from_numpy(resps.view(np.float32).reshape(n, 7)).pin_memory()

    def host_bufs():
        return dict(i=h_in.clone().pin_memory(), c=torch.empty((n, 8), dtype=torch.float32).pin_memory(),
                    h=torch.empty(n, dtype=torch.uint8).pin_memory(), t=torch.empty(n, dtype=torch.int32).pin_memory(),
                    s=torch.empty(n, dtype=torch.int32).pin_memory(), k=torch.empty((n, 2), dtype=torch.int32).pin_memory())
    hb = [host_bufs(), host_bufs()]

    def stream_steps(k, compact):
        tests, tickets = 0, []
        for i in range(k):
            if len(tickets) == 2:
                scene.detect_batch_wait(tickets.pop(0))
                tests += scene.stats()["narrowphase_tests"]
            b = hb[i % 2]
            if compact:
                tickets.append(scene.detect_batch_compact_submit(b["i"].data_ptr(), n, b["k"].data_ptr()))
            else:
                tickets.append(scene.detect_batch_submit(b["i"].data_ptr(), n, b["c"].data_ptr(), b["h"].data_ptr(),
                                                         b["t"].data_ptr(), b["s"].data_ptr()))
        for t in tickets:
            scene.detect_batch_wait(t)
            tests += scene.stats()["narrowphase_tests"]
        return tests

    def e2e_leg(compact):
        stream_steps(W, compact)
        ms, tests = ctx.timed_region(lambda: stream_steps(args.steps, compact))
        (tests,) = ctx.sum_over_ranks(tests)
        return ms, tests

    c_ms, c_tests = e2e_leg(True)
    f_ms, f_tests = e2e_leg(False)
    # the compact results say the same as the full records
    hits_c = hb[0]["k"][:, 1] != -1
    consistent = bool(torch.equal(hits_c, hb[0]["h"] != 0) and torch.equal(hb[0]["k"][:, 1], hb[0]["t"]) and
                      torch.equal(hb[0]["k"][:, 0][hits_c], hb[0]["c"][:, 0].view(torch.int32)[hits_c]))

    def step_single_call():
        b = hb[0]
        rc = L.xp_detect_batch(scene._h, C.c_void_p(b["i"].data_ptr()), n, 0, C.c_void_p(b["c"].data_ptr()),
                               C.c_void_p(b["h"].data_ptr()), C.c_void_p(b["t"].data_ptr()), C.c_void_p(b["s"].data_ptr()))
        xp._lib.check(rc, "xp_detect_batch")
    s_ms, s_stats = timed(step_single_call, max(3, args.steps // 2), 2)
    ceiling = pcie_ceiling(ctx, n * 28, n * 8)
    e2e = {"value": c_tests / (c_ms * 1e-3), "unit": "tests/s", "h2d_bytes_per_step": n * 28, "d2h_bytes_per_step": n * 8,
           "ms_per_step": c_ms / args.steps, "spheres_per_sec": world * n * args.steps / (c_ms * 1e-3),
           "api": "xp_detect_batch_compact_submit / xp_detect_batch_wait (host pointers, pinned), 2 batches in flight: 28 B "
                  "Response in, 8 B (time_to, tri_index) out per sphere; every step's H2D and D2H copies are inside the timed region",
           "hits": int(hits_c.sum().item()), "compact_equals_full_records": consistent,
           "pcie_ceiling": ceiling, "frac_of_pcie_ceiling": ceiling["both_ms_per_step"] / (c_ms / args.steps),
           "full_records": {"value": f_tests / (f_ms * 1e-3), "unit": "tests/s", "ms_per_step": f_ms / args.steps,
                            "h2d_bytes_per_step": n * 28, "d2h_bytes_per_step": n * 41,
                            "api": "xp_detect_batch_submit / _wait: the reference's 32 B Collision + hit + tri_index + status out"},
           "single_call": {"value": sum(s["narrowphase_tests"] for s in s_stats) * world / (s_ms * 1e-3), "unit": "tests/s",
                           "ms_per_step": s_ms / len(s_stats),
                           "api": "xp_detect_batch (one synchronous call per step: H2D, sweep, D2H back to back; L2 flushed between steps)"}}

    # ---- second half of BASELINE's metric: spheres RESOLVED per second = the full collision_response loop
    # (lib.rs:29-62) with up to 4 sweep-and-slide iterations per sphere, device-resident
    d_out = torch.empty_like(d_in)
    d_it = torch.empty(n, dtype=torch.int32, device=dev)
    d_st = torch.empty(n, dtype=torch.int32, device=dev)

    def step_response():
        scene.collision_response_batch_dev(d_in.data_ptr(), n, d_out.data_ptr(), max_iter=4,
                                           d_iter=d_it.data_ptr(), d_status=d_st.data_ptr())

    r_steps = max(3, args.steps // 4)
    r_ms, r_stats = timed(step_response, r_steps, 3)
    resolved = {"metric": "spheres_resolved_per_sec", "value": world * n * r_steps / (r_ms * 1e-3), "unit": "spheres/s",
                "max_iter": 4, "ms_per_step": r_ms / r_steps, "steps": r_steps,
                "loop": "device-resident: no host read-back between iterations (live counts, next rays and lists stay on the GPU)",
                "mean_iterations": float(d_it.float().mean().item()),
                "iter_cap_fraction": float(((d_st & 4) != 0).float().mean().item()),
                "tests_per_step": float(np.mean([s["narrowphase_tests"] for s in r_stats])),
                "kernel_launches_per_step": float(np.mean([s["kernel_launches"] for s in r_stats])),
                "stage_ms": {k: float(np.mean([s["stage_ms"][k] for s in r_stats])) for k in r_stats[0]["stage_ms"]
                             if r_stats[0]["stage_ms"][k] > 0}}

    extras = {}
    if not args.no_extras:
        # ---- both arithmetic modes side by side, and the price of far_exact (synchronous steps, L2 flushed)
        arith = {}
        for mode in ("exact", "fast"):
            scene.set_options(arith=mode)
            ms, sts = timed(step_device, max(5, args.steps // 2), 2)
            t = sum(s["narrowphase_tests"] for s in sts)
            npm = float(np.mean([s["stage_ms"]["narrowphase"] for s in sts]))
            arith[mode] = {"ms_per_step": ms / len(sts), "tests_per_sec": t * world / (ms * 1e-3), "narrowphase_ms": npm,
                           "roofline_frac": (flops / (npm * 1e-3) / 1e12) / peak if npm > 0 else None}
        arith["exact"]["note"] = "bit-identical to the reference (default)"
        arith["fast"]["note"] = "<= 1e-5 relative on well-conditioned contacts; grazing pairs may flip (DESIGN.md section 2)"
        scene.set_options(arith=args.arith, far_exact=True)
        ms, sts = timed(step_device, max(5, args.steps // 2), 2)
        arith["far_exact"] = {"ms_per_step": ms / len(sts), "tests_per_step": float(np.mean([s["narrowphase_tests"] for s in sts])),
                              "note": "XP_OPT_FAR_EXACT=1: second pass that also reproduces the reference's far-range grazing noise"}
        scene.set_options(far_exact=bool(args.far_exact))
        extras["arith"] = arith
        # ---- sustained: the loop of steps for >= 2 s
        if pipelined:
            per = max(total_ms / args.steps, 0.2)
            k = int(2200.0 / per) + 1
            with ClockSampler(ctx.local, enabled=rank == 0, period=0.2) as sc:
                s_ms2, s_st = timed_loop(k, 2)
            (s_tests,) = ctx.sum_over_ranks(sum(s["narrowphase_tests"] for s in s_st))
            extras["sustained"] = {"seconds": s_ms2 * 1e-3, "steps": k, "ms_per_step": s_ms2 / k,
                                   "value": s_tests / (s_ms2 * 1e-3), "unit": "tests/s", "clocks": sc.summary()}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        n_sample = min(n, 96 * cores)
        v, dt, used = cpu_sample(tris, resps, n_sample)
        cpu = {"value": v, "unit": "tests/s", "cores": used, "kind": "port",
               "sample": f"first {n_sample} spheres of the batch x all {len(tris)} triangles, brute force "
                         f"(the reference's algorithm, lib.rs:33-35), {dt:.1f} s",
               "spheres_per_sec": n_sample / dt}
        v1, dt1, _ = cpu_sample(tris, resps, 8, threads=1)
        cpu["single_thread"] = {"value": v1, "unit": "tests/s", "cores": 1, "sample": f"8 spheres, {dt1:.1f} s"}

    exch = ("none (one GPU)" if world == 1 else "nccl all_gather" if exchange is None else
            "one multimem.st per record span into an NVSwitch multicast address (fused with the record build)"
            if exchange.mode == "multicast" else
            ("copy engines: the record-building kernel fills this rank's slot of its own gather buffer, then one "
             "cudaMemcpyAsync (DMA over NVLink) per peer carries the 40 B x 2^20 span into that peer's buffer, on a second "
             "stream under the next step's sweep; barrier = one-element NCCL all-reduce") if pipelined else
            "p2p stores into mapped peer buffers by the record-building kernel (synchronous steps)")
    line = {"metric": "narrowphase_tests_per_sec", "value": value, "unit": "tests/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "arm": dict(arith=args.arith, method=args.method, scene=args.scene,
                        broadphase="height-grid walk (k_broadphase_grid)" if grid_walk else "LBVH walk (leaf runs)",
                        prune=args.prune, far_exact=bool(args.far_exact),
                        l2=("not flushed inside the loop (steps overlap): inputs larger than L2" if pipelined
                            else "flushed (256 MiB write) between timed steps"),
                        step=("loop of steps, two in flight: step i's record build + exchange runs on a second stream under "
                              "step i+1's sweep" if pipelined else "one synchronous call per step"),
                        exchange=exch),
            "spheres_per_sec": spheres_per_sec, "tests_per_step": tests_all / args.steps,
            "candidates_per_sphere": tests_all / args.steps / (world * n),
            "walk_visits_per_step": float(np.mean([s["node_visits"] for s in stats])),
            "stage_ms": st_ms,
            "hit_fraction": float((last_records[:, 9].contiguous().view(torch.int32) & 1).float().mean().item()),
            "exchange_verified": exchange_verified, "exchange_compact": exchange_compact,
            "single_step": {"ms_per_step": sync_ms / args.steps,
                            "value": sum(s_["narrowphase_tests"] for s_ in sync_stats) * world / (sync_ms * 1e-3),
                            "note": "one synchronous call per step (host waits for the result), L2 flushed between steps"},
            "clocks": clocks.summary() if rank == 0 else None, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "roofline_hbm": roofline_hbm, "resolved": resolved, "cpu_baseline": cpu,
            "build": {"ms": sum(build_stats["stage_ms"].values()), "stage_ms": build_stats["stage_ms"],
                      "launches": build_stats["kernel_launches"],
                      "note": "first build of the process (includes lazy kernel loading); steady-state rebuilds: configs.3"}}
    line.update(extras)
    if exchange is not None:
        exchange.close()
    scene.close()
    del d_in, contacts, flush
    torch.cuda.empty_cache()
    return line


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of the
# same command (bench.py --steps 2 --warmup 3 --no-cpu), profiles/r02zz_ncu_full.txt
TRAFFIC_NARROW = 482.9e6         # k_narrow_pairs: 473 MB read (315 MB of it the pair list) + 10 MB written (round 1: 610 MB; 956 MB before the spans were claimed dynamically)
TRAFFIC_BROADPHASE = 332.8e6     # k_broadphase_grid: 74 MB read + 259 MB written (the pair list)
TRAFFIC_SOURCE = "profiles/r02zz_ncu_full.txt"


# ---------------------------------------------------------------------------------------------
# the other named configs (bounded: a few steps each)
# ---------------------------------------------------------------------------------------------
def ev_ms(ctx: Ctx, fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    ms, _ = ctx.timed_region(lambda: [fn() for _ in range(steps)])
    return ms / steps


def run_config1(ctx: Ctx, args) -> dict:
    """single swept player sphere vs the res/map heightmap crop (80 000 triangles), one frame, host API"""
    import xp_physics_cuda as xp
    from xp_physics_cuda import scenes
    red = np.load(os.path.join(ROOT, "tests", "golden", "heightmap_crop_red.npy"))
    h = red.astype(np.float32) / np.float32(10.0)
    w = red.shape[0] - 1
    r = xp.make_responses([[0, red[100, 100] / 10.0 + 1.5, 0]], [[0.05, -0.05, 0]])
    out = {"workload": "config1: 1 swept player sphere x 80000-triangle res/map heightmap crop, collision_response, one frame, host API"}
    with xp.Scene(ctx.local) as s:
        s.set_heightmap(h, x0=-(w // 2), z0=-(w // 2), spacing=1.0)
        s.build()
        s.collision_response_batch(r)
        t0 = time.perf_counter()
        for _ in range(100):
            res, it, st, ln = s.collision_response_batch(r)
        out["gpu_us_per_call"] = (time.perf_counter() - t0) / 100 * 1e6
        out["iterations"] = int(it[0])
        out["kernel_launches_per_call"] = s.stats()["kernel_launches"]
        # the same call straight on the C ABI with the arrays allocated once: what a compiled host (the game) pays,
        # without the Python wrapper's four numpy allocations per call
        import ctypes as C
        L = xp._lib.lib()
        o_res, o_it, o_st = np.zeros(1, xp.RESPONSE_DTYPE), np.zeros(1, np.uint32), np.zeros(1, np.uint32)
        o_ln = np.zeros((1, 3), np.float32)
        ptr = [a.ctypes.data_as(C.c_void_p) for a in (r, o_res, o_it, o_st, o_ln)]
        for _ in range(20):
            L.xp_collision_response_batch(s._h, ptr[0], 1, 0, 0, ptr[1], ptr[2], ptr[3], ptr[4])
        t0 = time.perf_counter()
        for _ in range(200):
            rc = L.xp_collision_response_batch(s._h, ptr[0], 1, 0, 0, ptr[1], ptr[2], ptr[3], ptr[4])
        out["c_abi_us_per_call"] = (time.perf_counter() - t0) / 200 * 1e6
        out["c_abi_matches_wrapper"] = bool(rc == 0 and np.array_equal(o_res.view(np.uint8), np.ascontiguousarray(res).view(np.uint8)))
    if ctx.rank == 0 and not args.no_cpu:
        from oracle import xp_oracle as o
        tris = scenes.heightmap_from_pixels(red)
        t0 = time.perf_counter()
        for _ in range(10):
            w_out = o.collision_response(r, tris)
        out["cpu_reference_us_per_call"] = (time.perf_counter() - t0) / 10 * 1e6
        out["matches_cpu_reference_bitwise"] = bool(np.array_equal(res["c"].view(np.uint32)[0], np.asarray(w_out[0]["c"], np.float32).view(np.uint32)))
        out["note"] = ("a handful of spheres takes the small-batch path: the whole response loop in ONE launch (a warp per sphere "
                       "walks the LBVH 32 nodes at a time), one copy in, one copy out; round 1 went through the batch kernels "
                       "(33 launches, 295 us).  The CPU path brute-forces 80 000 triangles per iteration")
    return out


def run_config3(ctx: Ctx, args, scale: float = 1.0) -> dict:
    """16M spheres vs a 10M-triangle cube + icosphere field, LBVH rebuilt every frame"""
    import xp_physics_cuda as xp
    from xp_physics_cuda import scenes
    torch, dev = ctx.torch, ctx.dev
    T, S = int(10_000_000 * scale), int((1 << 24) * scale)
    tris, lo, hi = scenes.mesh_field(T, seed=2)
    d_tris = torch.from_numpy(tris.reshape(-1, 9)).to(dev)
    g = torch.Generator(device=dev)
    g.manual_seed(2 + ctx.rank)
    u = torch.rand((S, 7), generator=g, device=dev, dtype=torch.float32)
    d_in = torch.empty((S, 7), dtype=torch.float32, device=dev)
    d_in[:, 0:3] = torch.tensor(lo, device=dev) + u[:, 0:3] * torch.tensor(hi - lo, device=dev)
    d_in[:, 3] = 1.0
    dirs = torch.nn.functional.normalize(u[:, 3:6] - 0.5, dim=1)
    d_in[:, 4:7] = dirs * u[:, 6:7].pow(1.0 / 3.0)                      # uniform in the unit ball
    del u, dirs
    hits = torch.empty((S, 2), dtype=torch.int32, device=dev)
    s = xp.Scene(ctx.local, timing=True, arith=args.arith)              # prune = auto: distance windows once a sweep saw > 96 candidates per sphere
    s.set_stream(ctx.stream.cuda_stream)
    shift = torch.tensor([0.01, 0.0, 0.01] * 3, device=dev)
    build_ms, sweep_ms, tests, stage, sweep_stage = [], [], [], [], []

    def frame():
        d_tris.add_(shift)                                   # every mesh moves: the tree is stale
        s.set_triangles_dev(d_tris.data_ptr(), len(tris))
        s.build()
        b = s.stats()["stage_ms"]
        build_ms.append(sum(b.values())); stage.append(b)
        s.detect_batch_compact_dev(d_in.data_ptr(), S, hits.data_ptr())
        st = s.stats()
        sweep_ms.append(sum(st["stage_ms"].values())); tests.append(st["narrowphase_tests"]); sweep_stage.append(st["stage_ms"])
    ms = ev_ms(ctx, frame, 3, warmup=2)
    hbm, hbm_src = measured_hbm_gbs()
    bms = float(np.mean(build_ms[2:]))
    (tests_all,) = ctx.sum_over_ranks(float(np.mean(tests[2:])))
    out = {"workload": f"config3: {S} spheres x {len(tris)}-triangle cube+icosphere field per GPU, LBVH rebuilt every frame",
           "ms_per_frame": ms, "build_ms": bms, "sweep_ms": float(np.mean(sweep_ms[2:])),
           "build_stage_ms": {k: float(np.mean([x[k] for x in stage[2:]])) for k in stage[0] if stage[0][k] > 0},
           "sweep_stage_ms": {k: float(np.mean([x[k] for x in sweep_stage[2:]])) for k in sweep_stage[0] if sweep_stage[0][k] > 0},
           "tests_per_sec": tests_all / (ms * 1e-3), "spheres_per_sec": ctx.world * S / (ms * 1e-3),
           "candidates_per_sphere": float(np.mean(tests[2:])) / S,
           "build_hbm": {"algorithmic_bytes": 224.0 * len(tris), "achieved_gbs": 224.0 * len(tris) / (bms * 1e-3) / 1e9, "peak_gbs": hbm,
                         "peak_source": hbm_src, "frac": 224.0 * len(tris) / (bms * 1e-3) / 1e9 / hbm,
                         "note": "SURVEY 8(d): ~224 algorithmic B per triangle for the LBVH build"},
           "hit_fraction": float((hits[:, 1] != -1).float().mean().item()), "prune": "auto (distance windows)"}
    s.close()
    del d_tris, d_in, hits
    torch.cuda.empty_cache()
    return out


def run_config4(ctx: Ctx, args, scale: float = 1.0) -> dict:
    """4M spheres, 4-iteration sweep-and-slide vs clipmap terrain + voxel chunks"""
    import xp_physics_cuda as xp
    from xp_physics_cuda import scenes
    torch, dev = ctx.torch, ctx.dev
    S = int((1 << 22) * scale)
    tris = scenes.config4_scene()
    resps = scenes.scene_spheres(tris, S, seed=3 + ctx.rank, extent=900.0)
    d_in = ctx.dev_resps(resps)
    d_out = torch.empty_like(d_in)
    d_it = torch.empty(S, dtype=torch.int32, device=dev)
    d_st = torch.empty(S, dtype=torch.int32, device=dev)
    s = xp.Scene(ctx.local, tris, timing=True, arith=args.arith)
    s.set_stream(ctx.stream.cuda_stream)
    tests = []

    def step():
        s.collision_response_batch_dev(d_in.data_ptr(), S, d_out.data_ptr(), max_iter=4, d_iter=d_it.data_ptr(), d_status=d_st.data_ptr())
        tests.append(s.stats()["narrowphase_tests"])
    ms = ev_ms(ctx, step, 5)
    (tests_all,) = ctx.sum_over_ranks(float(np.mean(tests[2:])))
    out = {"workload": f"config4: {S} spheres per GPU, 4-iteration sweep-and-slide response, {len(tris)}-triangle clipmap terrain + voxel-model meshes",
           "ms_per_step": ms, "spheres_resolved_per_sec": ctx.world * S / (ms * 1e-3), "tests_per_sec": tests_all / (ms * 1e-3),
           "mean_iterations": float(d_it.float().mean().item()), "iter_cap_fraction": float(((d_st & 4) != 0).float().mean().item())}
    s.close()
    del d_in, d_out
    torch.cuda.empty_cache()
    return out


def run_config5(ctx: Ctx, args, scale: float = 1.0) -> dict:
    """128M spheres vs a 50M-triangle terrain sharded across the GPUs: 2^24 spheres per GPU (the full 2^27 at 8 GPUs),
    scene replicated, the global list of 40 B contact records gathered on every GPU."""
    import xp_physics_cuda as xp
    from xp_physics_cuda import scenes
    from xp_physics_cuda.distributed import ContactExchange
    torch, dist, dev, world, rank = ctx.torch, ctx.dist, ctx.dev, ctx.world, ctx.rank
    nx = int(5000 * np.sqrt(scale))
    S = int((1 << 24) * scale)
    h = scenes.sine_terrain_heights(nx, nx, seed=1)
    h_dev = torch.from_numpy(h).to(dev)
    d_in = dev_terrain_spheres(ctx, h_dev, S, seed=50 + rank)
    del h_dev
    s = xp.Scene(ctx.local, timing=True, arith=args.arith)
    s.set_stream(ctx.stream.cuda_stream)
    t0 = time.perf_counter()
    s.set_heightmap(h)
    s.build()
    up_s = time.perf_counter() - t0
    s.build()
    build = s.stats()
    ex = None
    gathered = None
    if world > 1:
        ex = ContactExchange(s, S, dev, n_buf=1) if not os.environ.get("XP_BENCH_NCCL_EXCHANGE") else None
        if ex is not None and not ex.ok:
            ex = None
        if ex is None:
            gathered = torch.empty((world, S, 10), dtype=torch.float32, device=dev)
    local = torch.empty((S, 10), dtype=torch.float32, device=dev)
    tests, stage = [], []

    def step():
        if ex is not None:
            ex.sweep(d_in.data_ptr(), S)
            dist.barrier()
        else:
            s.detect_contacts_dev(d_in.data_ptr(), S, local.data_ptr())
            if world > 1:
                dist.all_gather_into_tensor(gathered.view(-1), local.view(-1))
        st = s.stats()
        tests.append(st["narrowphase_tests"]); stage.append(st["stage_ms"])
    ms = ev_ms(ctx, step, 3, warmup=1)
    verified = None
    if world > 1:
        view = ex.view if ex is not None else gathered
        s.detect_contacts_dev(d_in.data_ptr(), S, local.data_ptr())
        torch.cuda.synchronize()
        ok = bool(torch.equal(view[rank][:S].view(torch.int32), local.view(torch.int32)))
        sums = torch.stack([view[r][:S].contiguous().view(torch.int32).to(torch.int64).sum() for r in range(world)])
        lo_, hi_ = sums.clone(), sums.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        flag = torch.tensor([1 if (ok and bool(torch.equal(lo_, hi_))) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        verified = bool(flag.item() == 1)
    (tests_all,) = ctx.sum_over_ranks(float(np.mean(tests[1:])))
    free, total = torch.cuda.mem_get_info()
    out = {"workload": f"config5: {world * S} spheres ({S} per GPU) x {2 * nx * nx}-triangle heightmap terrain, scene replicated per GPU, "
                       f"global contact list ({world * S * 40 / 1e9:.2f} GB) gathered on every GPU",
           "full_size": bool(world * S == (1 << 27) and nx == 5000),
           "ms_per_step": ms, "tests_per_sec": tests_all / (ms * 1e-3), "spheres_per_sec": world * S / (ms * 1e-3),
           "candidates_per_sphere": float(np.mean(tests[1:])) / S,
           "stage_ms": {k: float(np.mean([x[k] for x in stage[1:]])) for k in stage[0] if stage[0][k] > 0},
           "build_ms": sum(build["stage_ms"].values()), "upload_and_first_build_s": up_s,
           "exchange": ("none (one GPU)" if world == 1 else "nccl all_gather" if ex is None else
                        "p2p stores into mapped peer buffers by the record-building kernel, then a barrier"),
           "exchange_verified": verified, "gathered_bytes_per_gpu": world * S * 40,
           "device_mem_gb": (total - free) / 1e9}
    if ex is not None:
        ex.close()
        # the same steps with 16 B records (XP_OPT_CONTACT_RECORD = 1: centre at the contact + triangle): the exchange of the
        # 40 B records runs at NVLink line rate, so fewer bytes is the only way to make it shorter
        ex16 = ContactExchange(s, S, dev, n_buf=1, record="compact")
        if ex16.ok:
            tests.clear(); stage.clear()

            def step16():
                ex16.sweep(d_in.data_ptr(), S)
                dist.barrier()
                st = s.stats()
                tests.append(st["narrowphase_tests"]); stage.append(st["stage_ms"])
            ms16 = ev_ms(ctx, step16, 3, warmup=1)
            local16 = torch.empty((S, 4), dtype=torch.float32, device=dev)
            s.detect_contacts_dev(d_in.data_ptr(), S, local16.data_ptr())
            torch.cuda.synchronize()
            ok = bool(torch.equal(ex16.view[rank][:S].view(torch.int32), local16.view(torch.int32)))
            sums = torch.stack([ex16.view[r][:S].contiguous().view(torch.int32).to(torch.int64).sum() for r in range(world)])
            lo_, hi_ = sums.clone(), sums.clone()
            dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
            flag = torch.tensor([1 if (ok and bool(torch.equal(lo_, hi_))) else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            out["compact_records"] = {"record": "xp_contact16: (position, tri_index), 16 B", "ms_per_step": ms16,
                                      "spheres_per_sec": world * S / (ms16 * 1e-3), "gathered_bytes_per_gpu": world * S * 16,
                                      "finalize_ms": float(np.mean([x["finalize"] for x in stage[1:]])),
                                      "exchange_verified": bool(flag.item() == 1)}
            del local16
        ex16.close()
        s.set_options(contact_record="full")
    s.close()
    del d_in, local, gathered
    torch.cuda.empty_cache()
    return out


CONFIG_RUNNERS = {1: run_config1, 3: run_config3, 4: run_config4, 5: run_config5}


def run_b200(args):
    ctx = Ctx()
    if args.config == 2:
        line = run_config2(ctx, args)
        if not args.no_extras and not args.no_configs:
            cfg = {}
            for k in (1, 3, 4, 5):
                try:
                    cfg[str(k)] = CONFIG_RUNNERS[k](ctx, args)
                except Exception as e:                     # noqa: BLE001 -- a config that cannot run here says so instead of killing the headline
                    cfg[str(k)] = {"error": repr(e)[:300]}
            line["configs"] = cfg
    else:
        c = CONFIG_RUNNERS[args.config](ctx, args)
        metric = {1: ("collision_response_latency_us", c.get("c_abi_us_per_call", c.get("gpu_us_per_call")), "us", False),
                  3: ("narrowphase_tests_per_sec", c.get("tests_per_sec"), "tests/s", True),
                  4: ("spheres_resolved_per_sec", c.get("spheres_resolved_per_sec"), "spheres/s", True),
                  5: ("narrowphase_tests_per_sec", c.get("tests_per_sec"), "tests/s", True)}[args.config]
        line = {"metric": metric[0], "value": metric[1], "unit": metric[2], "n_gpus": ctx.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": c.get("ms_per_step", c.get("ms_per_frame")), "higher_is_better": metric[3],
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": c["workload"], "arith": args.arith}, "detail": c}
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json config to make the headline (default 2)")
    ap.add_argument("--arith", default="exact", choices=["exact", "fast", "exact_literal", "fast_literal"])
    ap.add_argument("--method", default="lbvh_pairs", choices=["wide_fused", "wide_pairs", "lbvh_fused", "lbvh_pairs", "brute", "q4_pairs"])
    ap.add_argument("--prune", type=int, default=2, help="XP_OPT_PRUNE (2 = auto, the library default)")
    ap.add_argument("--scene", default="heightmap", choices=["heightmap", "soup"],
                    help="how the config-2 terrain is handed over: xp_scene_set_heightmap (grid walk) or xp_scene_set_triangles (LBVH walk)")
    ap.add_argument("--far-exact", type=int, default=0, help="1: XP_OPT_FAR_EXACT (also reproduce the reference's far-range grazing noise)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-extras", action="store_true", help="headline only: no arith / sustained / configs blocks")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs block")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
