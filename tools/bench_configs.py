#!/usr/bin/env python
"""Secondary measurements: the BASELINE.json configs that are not the bench.py headline
(config 1 latency, config 3 rebuild-per-frame, config 4 four-iteration response on the clipmap+voxel
scene, config 5 one GPU's shard at full size).  One JSON line per config on stdout.

    python tools/bench_configs.py [--configs 1,3,4,5] [--scale 1.0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "xp-game-engine_b200")]


def ev_time(torch, fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,4,5")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink sphere / triangle counts (smoke runs)")
    ap.add_argument("--method", default="lbvh_pairs")
    ap.add_argument("--prune", type=int, default=0)
    args = ap.parse_args()
    import torch

    import xp_physics_cuda as xp
    from xp_physics_cuda import scenes
    todo = [int(x) for x in args.configs.split(",")]
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()

    def dev_resps(r):
        return torch.from_numpy(r.view(np.float32).reshape(-1, 7)).to(dev)

    if 1 in todo:   # one swept player sphere vs the res/map heightmap crop, one frame, host API
        red = np.load(os.path.join(ROOT, "tests", "golden", "heightmap_crop_red.npy"))
        tris = scenes.heightmap_from_pixels(red)
        r = xp.make_responses([[0, red[100, 100] / 10.0 + 1.5, 0]], [[0.05, -0.05, 0]])
        with xp.Scene(0, tris) as s:
            s.collision_response_batch(r)
            t0 = time.perf_counter()
            for _ in range(200):
                out, it, st, ln = s.collision_response_batch(r)
            us = (time.perf_counter() - t0) / 200 * 1e6
        from oracle import xp_oracle as o
        t0 = time.perf_counter()
        for _ in range(20):
            o.collision_response(r, tris)
        cpu_us = (time.perf_counter() - t0) / 20 * 1e6
        print(json.dumps({"config": 1, "what": "1 sphere x 80000-triangle heightmap crop, collision_response, host API",
                          "gpu_us_per_call": us, "cpu_oracle_us_per_call": cpu_us, "iterations": int(it[0])}), flush=True)

    if 3 in todo:   # 16M spheres vs 10M-triangle mesh field, LBVH rebuilt every frame
        T, S = int(10_000_000 * args.scale), int((1 << 24) * args.scale)
        tris, lo, hi = scenes.mesh_field(T, seed=2)
        resps = scenes.field_spheres(lo, hi, S, seed=2)
        d_tris = torch.from_numpy(tris.reshape(-1, 9)).to(dev)
        d_in = dev_resps(resps)
        contacts = torch.empty((S, 10), dtype=torch.float32, device=dev)
        s = xp.Scene(0, timing=True, method=args.method, prune=bool(args.prune))
        s.set_stream(stream.cuda_stream)
        shift = torch.tensor([0.01, 0.0, 0.01] * 3, device=dev)
        build_ms, sweep_ms, tests = [], [], []

        def frame():
            d_tris.add_(shift)                                   # every mesh moves: the tree is stale
            s.set_triangles_dev(d_tris.data_ptr(), len(tris))
            s.build()
            build_ms.append(sum(s.stats()["stage_ms"].values()))
            s.detect_contacts_dev(d_in.data_ptr(), S, contacts.data_ptr())
            st = s.stats()
            sweep_ms.append(sum(st["stage_ms"].values())); tests.append(st["narrowphase_tests"])
        ms = ev_time(torch, frame, 5)
        print(json.dumps({"config": 3, "method": args.method, "prune": args.prune,
                          "what": f"{S} spheres x {len(tris)}-triangle cube+icosphere field, LBVH rebuild per frame",
                          "ms_per_frame": ms, "build_ms": float(np.mean(build_ms[2:])), "sweep_ms": float(np.mean(sweep_ms[2:])),
                          "tests_per_frame": float(np.mean(tests[2:])), "tests_per_sec": float(np.mean(tests[2:])) / (ms * 1e-3),
                          "spheres_per_sec": S / (ms * 1e-3), "build_triangles_per_sec": len(tris) / (np.mean(build_ms[2:]) * 1e-3),
                          "hit_fraction": float((contacts[:, 9].contiguous().view(torch.int32) & 1).float().mean().item())}), flush=True)
        s.close()
        del d_tris, d_in, contacts

    if 4 in todo:   # 4M spheres, 4-iteration sweep-and-slide vs clipmap terrain + voxel chunks
        S = int((1 << 22) * args.scale)
        tris = scenes.config4_scene()
        resps = scenes.scene_spheres(tris, S, seed=3, extent=900.0)
        d_in = dev_resps(resps)
        d_out = torch.empty_like(d_in)
        d_it = torch.empty(S, dtype=torch.int32, device=dev)
        d_st = torch.empty(S, dtype=torch.int32, device=dev)
        s = xp.Scene(0, tris, timing=True, method=args.method, prune=bool(args.prune))
        s.set_stream(stream.cuda_stream)
        tests = []

        def step():
            s.collision_response_batch_dev(d_in.data_ptr(), S, d_out.data_ptr(), max_iter=4, d_iter=d_it.data_ptr(),
                                           d_status=d_st.data_ptr())
            tests.append(s.stats()["narrowphase_tests"])
        ms = ev_time(torch, step, 5)
        print(json.dumps({"config": 4, "what": f"{S} spheres, max_iter=4, {len(tris)}-triangle clipmap + voxel scene",
                          "ms_per_step": ms, "spheres_resolved_per_sec": S / (ms * 1e-3), "tests_per_sec": float(np.mean(tests[2:])) / (ms * 1e-3),
                          "mean_iterations": float(d_it.float().mean().item()),
                          "iter_cap_fraction": float(((d_st & 4) != 0).float().mean().item())}), flush=True)
        s.close()
        del d_in, d_out

    if 5 in todo:   # one GPU's shard of config 5: 2^27 / 8 spheres vs the full 50M-triangle terrain
        nx = int(5000 * np.sqrt(args.scale))
        S = int((1 << 24) * args.scale)
        tris, h = scenes.sine_terrain(nx, nx, seed=1)
        resps = scenes.terrain_spheres(h, S, seed=5)
        d_in = dev_resps(resps)
        contacts = torch.empty((S, 10), dtype=torch.float32, device=dev)
        s = xp.Scene(0, timing=True)
        s.set_stream(stream.cuda_stream)
        t0 = time.perf_counter()
        s.set_triangles(tris)
        s.build()
        s.build()
        build = s.stats()
        up_s = time.perf_counter() - t0
        tests = []

        def step():
            s.detect_contacts_dev(d_in.data_ptr(), S, contacts.data_ptr())
            tests.append(s.stats()["narrowphase_tests"])
        ms = ev_time(torch, step, 3, warmup=1)
        st = s.stats()
        print(json.dumps({"config": 5, "what": f"one GPU's shard: {S} spheres x {len(tris)}-triangle terrain (scene replicated per GPU)",
                          "ms_per_step": ms, "tests_per_sec": float(np.mean(tests)) / (ms * 1e-3), "spheres_per_sec": S / (ms * 1e-3),
                          "build_ms": sum(build["stage_ms"].values()), "upload_and_2_builds_s": up_s,
                          "stage_ms": {k: v for k, v in st["stage_ms"].items() if v > 0},
                          "device_mem_gb": torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9}), flush=True)
        s.close()


if __name__ == "__main__":
    main()
