#!/usr/bin/env python
"""bench.py -- swept-sphere vs triangle narrowphase tests/sec (and spheres resolved/sec).

Workload (BASELINE.json configs[1]): 2^20 swept unit spheres vs a 1 001 112-triangle synthetic
heightmap terrain, single sweep, per GPU.  A "step" is one closest-collision sweep of the whole
sphere batch against the resident scene (triangle records + LBVH are uploaded / built once,
outside the timed region, exactly like the reference's `&[Triangle]` argument exists before
collision_response is called).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 is launched by torchrun (one rank per GPU): the sphere batch is sharded (weak scaling:
2^20 spheres per GPU, scene + LBVH replicated) and the global list of 40 B contact records is
assembled in every rank's gather buffer over NVLink peer memory (NCCL all_gather as the fallback).

Prints ONE JSON line (rank 0).  `value` = narrowphase tests/s with inputs resident in HBM, measured
over a loop of steps with two in flight (step i's contact build + exchange on a second stream under
step i+1's sweep; `single_step` = one synchronous step with an L2 flush in between);
`e2e` = the same metric through the host-pointer C ABI for a stream of batches (pinned host buffers,
every step's H2D + D2H inside the timed region, two batches in flight; `e2e.single_call` = one
synchronous call per step); `roofline` = the dominant kernel against the measured fp32 FFMA peak;
`cpu_baseline` = the CPU oracle (a port of the reference's algorithm: brute force over every
triangle) timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
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
    return {"workload": "config2: 2^20 swept unit spheres x 1001112-triangle sine+noise heightmap terrain, "
                        "single sweep per step, horizon=inf (reference-faithful)",
            "spheres_per_gpu": N_SPHERES, "triangles": 2 * NX * NZ,
            "parallelism": f"spheres sharded over {n_gpus} GPU(s), scene+LBVH replicated",
            "l2": "flushed (256 MiB write) between timed steps"}


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.stop, self.t = index, [], threading.Event(), None

    def __enter__(self):
        def run():
            while not self.stop.is_set():
                try:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    self.rows.append([x.strip() for x in out.strip().split(",")])
                except Exception:
                    pass
                self.stop.wait(0.05)
        self.t = threading.Thread(target=run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
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
# the B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import xp_physics_cuda as xp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- xp_physics_cuda has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    tris, resps, heights = build_workload(seed=1 + rank, with_heights=True)
    n = len(resps)
    scene = xp.Scene(local, arith=args.arith, method=args.method, prune=args.prune, timing=True)
    if args.far_exact:
        scene.set_options(far_exact=True)
    stream = torch.cuda.current_stream()
    scene.set_stream(stream.cuda_stream)
    if args.scene == "heightmap":
        scene.set_heightmap(heights)              # the terrain as the reference holds it: a height grid (emitted on the device)
    else:
        scene.set_triangles(tris)                 # the same terrain as an anonymous triangle soup: LBVH walk
    scene.build()
    build_stats = scene.stats()

    # device-resident inputs / outputs
    d_in = torch.from_numpy(resps.view(np.float32).reshape(n, 7)).to(dev)
    contacts = torch.empty((world, n, 10), dtype=torch.float32, device=dev)   # 40 B records, gather buffer
    mine = contacts[rank]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # The kernel that builds the contact records stores them into every rank's gather buffer over NVLink peer
    # memory (CUDA IPC); with one GPU the only "peer" is the local buffer.  If the mapping is unavailable the
    # steps fall back to the synchronous path + an NCCL all-gather.
    exchange = None
    if not os.environ.get("XP_BENCH_NCCL_EXCHANGE"):
        from xp_physics_cuda.distributed import ContactExchange
        exchange = ContactExchange(scene, n, dev)
        if not exchange.ok:
            exchange = None
        else:
            mine = exchange.view[rank][:n]
    pipelined = exchange is not None and exchange.mode == "p2p" and not os.environ.get("XP_BENCH_SYNC_STEPS")

    def step_device():
        if exchange is not None:
            exchange.sweep(d_in.data_ptr(), n)       # compute + all-gather in one kernel (P2P stores)
            if world > 1:
                dist.barrier()                       # every rank's records have landed everywhere
            return
        # writes this rank's contact records straight into its slot of the gather buffer
        scene.detect_contacts_dev(d_in.data_ptr(), n, mine.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(contacts.view(-1), mine.reshape(-1))

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        stats = []
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for a, b in ev:
            if not os.environ.get("XP_BENCH_NO_FLUSH"):       # experiment knob; reported numbers always flush
                flush.fill_(1)                 # L2 flush, outside the event pair
            a.record(stream)
            fn()
            b.record(stream)
            stats.append(scene.stats())
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in ev)
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, stats

    def loop_device(k):
        """k steps as a loop: submit step i, then complete step i-1 -- step i's sweep runs while step i-1's
        records cross NVLink (second stream) and its barrier completes.  Returns the per-step stats."""
        pending, out = [], []
        for _ in range(k):
            if len(pending) == 2:
                exchange.wait(pending.pop(0))
                out.append(scene.stats())
            pending.append(exchange.sweep_async(d_in.data_ptr(), n))
        while pending:
            exchange.wait(pending.pop(0))
            out.append(scene.stats())
        return out

    with ClockSampler(local) as clocks:                      # nvidia-smi samples while both timed regions run
        sync_ms, sync_stats = timed(step_device, args.steps, max(args.warmup, 3))     # one synchronous call per step, L2 flushed
        if pipelined:
            loop_device(max(args.warmup, 3))
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            stats = loop_device(args.steps)
            ev1.record(stream)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            total_ms = ev0.elapsed_time(ev1)
            if world > 1:
                t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                total_ms = float(t.item())
        else:
            total_ms, stats = sync_ms, sync_stats
    last_records = (exchange.views[exchange._half_i][rank][:n] if pipelined else mine).clone()
    tests_local = sum(s["narrowphase_tests"] for s in stats)
    launches = sum(s["kernel_launches"] for s in stats)
    tests_all = tests_local
    if world > 1:
        t = torch.tensor([tests_local], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        tests_all = float(t.item())
    sec = total_ms * 1e-3
    value = tests_all / sec
    spheres_per_sec = world * n * args.steps / sec

    # ---- rooflines, from the CUDA-event stage times the library records on its own stream around
    # each kernel inside the timed steps (XP_OPT_TIMING; no extra synchronisation)
    st_ms = {k: float(np.mean([s["stage_ms"][k] for s in stats])) for k in stats[0]["stage_ms"]}
    two_stage = args.method in ("lbvh_pairs", "wide_pairs", "q4_pairs")
    # path mix of the narrowphase tests: one untimed diagnostic sweep with the fused kernel, which
    # counts the code path of every test (the two-stage kernels carry no counters)
    scene.set_options(method="lbvh_fused", prune=False)
    scene.detect_contacts_dev(d_in.data_ptr(), n, mine.data_ptr())
    paths = {k: float(v) for k, v in scene.stats()["path_tests"].items()}
    scene.set_options(method=args.method, prune=args.prune)
    flops = sum(paths[k] * PATH_FLOPS[k] for k in PATH_FLOPS)
    try:
        peak = xp.probe_fp32_peak(local)
        peak_src = "measured: FFMA micro-benchmark xp_probe_fp32_peak in this run (nominal 74.4)"
    except Exception:
        peak, peak_src = NOMINAL_FP32_TFLOPS, "nominal 148 SM x 128 lanes x 2 x 1.965 GHz"
    np_ms = st_ms["narrowphase"] if two_stage else st_ms["traverse"] + st_ms["narrowphase"]
    achieved = flops / (np_ms * 1e-3) / 1e12 if np_ms > 0 else 0.0
    kname = {"wide_fused": "k_sweep (32-wide tree traversal + narrowphase, fused)",
             "wide_pairs": "k_narrow_pairs", "lbvh_fused": "k_traverse (binary LBVH + narrowphase, fused)",
             "lbvh_pairs": "k_narrow_pairs", "q4_pairs": "k_narrow_pairs<FILTER>", "brute": "k_brute"}[args.method]
    roofline = {"bound": "fp32", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak else None, "peak_source": peak_src, "kernel_ms": np_ms,
                "algorithmic_flops_per_launch": flops, "flops_per_test_by_path": PATH_FLOPS, "path_mix": paths,
                "traffic": TRAFFIC_NARROW, "traffic_source": TRAFFIC_SOURCE,
                "note": "231 algorithmic flops are ~580 issued instructions (11 IEEE div, 10 IEEE sqrt); "
                        "ncu (profiles/r01h_ncu_full.txt): issue slots 77% busy, L1 data pipe 65% busy, FMA pipe 42% busy, 30.1/32 lanes active"}
    # the broadphase is the longest kernel of the step and is memory-latency / L1-bound, not compute
    tests_step = float(np.mean([s["narrowphase_tests"] for s in stats]))
    bp_bytes = 64.0 * n + 8.0 * tests_step                       # ray records in + candidate pairs out
    bp_ms = st_ms["traverse"]
    roofline_hbm = {"bound": "hbm", "kernel": {"lbvh_pairs": "k_broadphase", "q4_pairs": "k_broadphase4"}.get(args.method, "k_sweep<EMIT>"),
                    "achieved": bp_bytes / (bp_ms * 1e-3) / 1e9 if (two_stage and bp_ms > 0) else None,
                    "peak": measured_hbm_gbs(), "unit": "GB/s", "kernel_ms": bp_ms,
                    "frac": (bp_bytes / (bp_ms * 1e-3) / 1e9) / measured_hbm_gbs() if (two_stage and bp_ms > 0) else None,
                    "algorithmic_bytes_per_launch": bp_bytes, "traffic": TRAFFIC_BROADPHASE,
                    "traffic_source": TRAFFIC_SOURCE,
                    "note": "64 B ray record per sphere + 8 B per candidate pair; the 64 B nodes it chases "
                            "(83 M visits) are L1/L2 hits (scene is L2-resident), so the kernel is bound by L1 "
                            "throughput and the latency of one dependent L2 access per step (ncu: L1 data pipe 87% busy, issue slots 55%, 28 of 32 lanes active, DRAM 4%), not by HBM"}

    # e2e: the host-pointer C-ABI call, pinned host buffers, H2D + D2H inside the timed region
    h_in = torch.from_numpy(resps.view(np.float32).reshape(n, 7)).pin_memory()
    h_col = torch.empty((n, 8), dtype=torch.float32).pin_memory()
    h_hit = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_tri = torch.empty(n, dtype=torch.int32).pin_memory()
    h_st = torch.empty(n, dtype=torch.int32).pin_memory()
    L = xp._lib.lib()
    import ctypes as C

    def step_e2e():
        rc = L.xp_detect_batch(scene._h, C.c_void_p(h_in.data_ptr()), n, 0, C.c_void_p(h_col.data_ptr()),
                               C.c_void_p(h_hit.data_ptr()), C.c_void_p(h_tri.data_ptr()), C.c_void_p(h_st.data_ptr()))
        xp._lib.check(rc, "xp_detect_batch")

    e2e_ms, e2e_stats = timed(step_e2e, args.steps, max(args.warmup, 3))
    e2e_tests = sum(s["narrowphase_tests"] for s in e2e_stats)

    # the same metric for a STREAM of batches: xp_detect_batch_submit / _wait with two batches in flight
    # (double-buffered pinned host arrays), so one batch's PCIe copies run under the other's kernels.
    # Every step still copies its 28 B/sphere in and its 41 B/sphere out inside the timed region.  No explicit
    # L2 flush here (steps overlap): one step streams ~0.5 GB of rays, pairs and results through the 126 MB L2.
    hb = [dict(i=h_in, c=h_col, h=h_hit, t=h_tri, s=h_st),
          dict(i=h_in.clone().pin_memory(), c=torch.empty_like(h_col).pin_memory(), h=torch.empty_like(h_hit).pin_memory(),
               t=torch.empty_like(h_tri).pin_memory(), s=torch.empty_like(h_st).pin_memory())]

    def stream_steps(k):
        tests, tickets = 0, []
        for i in range(k):
            if len(tickets) == 2:
                scene.detect_batch_wait(tickets.pop(0))
                tests += scene.stats()["narrowphase_tests"]
            b = hb[i % 2]
            tickets.append(scene.detect_batch_submit(b["i"].data_ptr(), n, b["c"].data_ptr(), b["h"].data_ptr(),
                                                     b["t"].data_ptr(), b["s"].data_ptr()))
        for t in tickets:
            scene.detect_batch_wait(t)
            tests += scene.stats()["narrowphase_tests"]
        return tests

    stream_steps(max(args.warmup, 3))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    p_tests = stream_steps(args.steps)
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    p_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([p_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        p_ms = float(t.item())
        t = torch.tensor([e2e_tests, p_tests], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        e2e_tests, p_tests = float(t[0].item()), float(t[1].item())
    same = bool(torch.equal(hb[0]["c"], hb[1]["c"]) and torch.equal(hb[0]["h"], hb[1]["h"]) and torch.equal(hb[0]["t"], hb[1]["t"]))
    e2e = {"value": p_tests / (p_ms * 1e-3), "unit": "tests/s", "h2d_bytes_per_step": n * 28,
           "d2h_bytes_per_step": n * (32 + 1 + 4 + 4), "ms_per_step": p_ms / args.steps,
           "spheres_per_sec": world * n * args.steps / (p_ms * 1e-3),
           "api": "xp_detect_batch_submit / xp_detect_batch_wait (host pointers, pinned), 2 batches in flight; "
                  "every step's H2D and D2H copies are inside the timed region",
           "hits": int(h_hit.sum().item()), "both_buffers_identical": same,
           "single_call": {"value": e2e_tests / (e2e_ms * 1e-3), "unit": "tests/s", "ms_per_step": e2e_ms / args.steps,
                           "api": "xp_detect_batch (one synchronous call per step: H2D, sweep, D2H back to back; "
                                  "L2 flushed between steps)",
                           "stage_ms": {k: float(np.mean([s["stage_ms"][k] for s in e2e_stats]))
                                        for k in e2e_stats[0]["stage_ms"] if e2e_stats[0]["stage_ms"][k] > 0}}}

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
                "mean_iterations": float(d_it.float().mean().item()),
                "iter_cap_fraction": float(((d_st & 4) != 0).float().mean().item()),
                "tests_per_step": float(np.mean([s["narrowphase_tests"] for s in r_stats])),
                "stage_ms": {k: float(np.mean([s["stage_ms"][k] for s in r_stats])) for k in r_stats[0]["stage_ms"]
                             if r_stats[0]["stage_ms"][k] > 0}}

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

    if rank == 0:
        line = {"metric": "narrowphase_tests_per_sec", "value": value, "unit": "tests/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": dict(workload_config(world), arith=args.arith, method=args.method, scene=args.scene,
                                                    prune=bool(args.prune), far_exact=bool(args.far_exact),
                                                    l2=("not flushed inside the loop (steps overlap): one step's inputs -- 128 MB of "
                                                        "scene records + 29 MB of spheres -- exceed the 126 MB L2 and a step streams "
                                                        "~0.5 GB through it; single_step is the same step with an explicit L2 flush"
                                                        if pipelined else "flushed (256 MiB write) between timed steps"),
                                                    step=("loop of steps, two in flight: step i's record build + exchange runs on a "
                                                          "second stream under step i+1's sweep" if pipelined else
                                                          "one synchronous call per step"),
                                                    exchange=("none" if world == 1 else "nccl all_gather" if exchange is None else
                                                              "one multimem.st per record span into an NVSwitch multicast address bound to "
                                                              "every rank's gather buffer (fused with the record build)"
                                                              if exchange.mode == "multicast" else
                                                              "p2p stores into mapped peer buffers (fused with the record build)")),
                "spheres_per_sec": spheres_per_sec, "tests_per_step": tests_all / args.steps,
                "node_visits_per_step": float(np.mean([s["node_visits"] for s in stats])),
                "stage_ms": {k: float(np.mean([s["stage_ms"][k] for s in stats])) for k in stats[0]["stage_ms"]},
                "hit_fraction": float((last_records[:, 9].contiguous().view(torch.int32) & 1).float().mean().item()),
                "single_step": {"ms_per_step": sync_ms / args.steps,
                                "value": sum(s_["narrowphase_tests"] for s_ in sync_stats) * world / (sync_ms * 1e-3),
                                "note": "one synchronous call per step (host waits for the result), L2 flushed between steps"
                                        + ("" if world == 1 else "; value assumes every rank ran the same number of tests")},
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "roofline_hbm": roofline_hbm, "resolved": resolved,
                "cpu_baseline": cpu,
                "build": {"ms": sum(build_stats["stage_ms"].values()), "stage_ms": build_stats["stage_ms"],
                          "launches": build_stats["kernel_launches"],
                          "note": "first build of the process: includes lazy kernel loading; steady-state rebuilds "
                                  "(config 3) are timed by tools/bench_configs.py"}}
        print(json.dumps(line), flush=True)
    scene.close()
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of the
# same command (bench.py --steps 2 --warmup 3 --no-cpu), profiles/r01h_ncu_full.txt
TRAFFIC_NARROW = 628.5e6         # k_narrow_pairs: ~600 MB read + ~20 MB written (315 MB of the reads is the pair list)
TRAFFIC_BROADPHASE = 414.7e6     # k_broadphase: ~128 MB read + ~277 MB written (the pair list)
TRAFFIC_SOURCE = "profiles/r01h_ncu_full.txt"


def measured_hbm_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arith", default="exact", choices=["exact", "fast", "exact_literal", "fast_literal"])
    ap.add_argument("--method", default="lbvh_pairs", choices=["wide_fused", "wide_pairs", "lbvh_fused", "lbvh_pairs", "brute", "q4_pairs"])
    ap.add_argument("--prune", type=int, default=0)
    ap.add_argument("--scene", default="heightmap", choices=["heightmap", "soup"],
                    help="how the config-2 terrain is handed over: xp_scene_set_heightmap (grid walk) or xp_scene_set_triangles (LBVH walk)")
    ap.add_argument("--far-exact", type=int, default=0, help="1: XP_OPT_FAR_EXACT (also reproduce the reference's far-range grazing noise)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
