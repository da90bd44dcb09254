#!/usr/bin/env python
"""BASELINE config 5: worst-case geometry sweep, walls per world x rays per robot, at 262,144
worlds over 8 GPUs = 32,768 per GPU (16,384 for >= 1000 walls, to bound host-side scene
generation).  Prints one roofline point per cell.

    python tools/sweep_config5.py [--out gpurun_out/sweep_config5.json]            # one GPU's shard
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_config5.py
                                          # the whole batch: every rank its shard, max over ranks
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import aitk_robots_b200 as rb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/sweep_config5.json")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--segments", default="10,30,100,300,1000,2000")
ap.add_argument("--rays", default="1,4,16,64,256")
ap.add_argument("--paths", default="auto", help="comma list of auto,fused,solo,mini[2|4|8|16]: launch plans to time per cell")
args = ap.parse_args()

rank, world_size = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world_size > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))


def sync():
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize()


peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
hbm_peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
ffma, _ = rb.measure_fp32_peak(local_rank, 4096)
rows = []
stream = torch.cuda.current_stream()
for S in [int(v) for v in args.segments.split(",")]:
    B = 32768 if S < 1000 else 16384
    t0 = time.time()
    base = rb.scenes.config5(n_worlds=B, n_segments=S, n_rays=256, first_world=rank * B)
    gen_s = time.time() - t0
    for n in [int(v) for v in args.rays.split(",")]:
        base.robots[0].devices = [rb.RangeSpec(position=(0.0, 0.0), a=360.0 * i / n, max=200.0, width=0.0,
                                               name="ray%d" % i) for i in range(n)]
        for path in args.paths.split(","):
            for k in ("BRS_SOLO", "BRS_MINI", "BRS_MINI_G"):
                os.environ.pop(k, None)
            if path.startswith("mini"):  # mini, mini2, mini4, mini8, mini16: the team-per-world plan (team size)
                if n > 16:
                    continue
                os.environ["BRS_MINI"] = "1"
                if path[4:]:
                    os.environ["BRS_MINI_G"] = path[4:]
            elif path != "auto":
                os.environ["BRS_SOLO"] = "1" if path == "solo" else "0"
                os.environ["BRS_MINI"] = "0"
            eng = rb.BatchEngine(base, device=local_rank)
            tests = eng.count_tests()
            for _ in range(3):
                eng.step(stream=stream.cuda_stream)
            sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                eng.step(stream=stream.cuda_stream)
            e1.record(stream)
            sync()
            ms = e0.elapsed_time(e1) / args.steps
            if world_size > 1:  # max over ranks; tests summed over the shards
                t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t[0])
                t = torch.tensor([float(tests)], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                tests = float(t[0])
            alg_bytes = B * (20 * eng.seg_cap + 12 + 72 + 48 + 1 + 4 * n)
            row = {"segments": S, "rays": n, "worlds": B * world_size, "n_gpus": world_size, "path_requested": path,
                   "ms_per_step": ms, "world_steps_per_sec": B * world_size / (ms * 1e-3),
                   "tests_per_sec": tests / (ms * 1e-3), "fp32_frac": tests * 13 / (ms * 1e-3) / 1e12 / (ffma * world_size),
                   "hbm_frac": alg_bytes / (ms * 1e-3) / 1e9 / hbm_peak, "launch": eng.launch_info()}
            rows.append(row)
            if rank == 0:
              print("S=%5d rays=%4d worlds=%6d %-5s ms=%8.4f tests/s=%.3g fp32=%.3f hbm=%.3f | %s" % (
                S, n, B, path, ms, row["tests_per_sec"], row["fp32_frac"], row["hbm_frac"], row["launch"]["path"][:10]), flush=True)
            eng.close()
    if rank == 0:
        print("  (scene generation %.1f s)" % gen_s, flush=True)
if rank == 0:
    json.dump({"ffma_peak_tflops": ffma, "hbm_peak_gbs": hbm_peak, "n_gpus": world_size, "cells": rows}, open(args.out, "w"), indent=1)
if world_size > 1:
    dist.barrier()
    dist.destroy_process_group()
