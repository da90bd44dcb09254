#!/usr/bin/env python
"""brs_steps (n steps in one launch) against n launches of brs_step, CUDA events, per config.

    python tools/steps_timing.py [--steps 20] [--out gpurun_out/steps_timing.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import aitk_robots_b200 as rb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--out", default="gpurun_out/steps_timing.json")
ap.add_argument("--configs", default="config2,config2x4,config4,config3,config1,config1x1,config5-100-16,config5-2000-256")
args = ap.parse_args()
makers = {
    "config2": lambda: rb.scenes.config2(n_worlds=4096),
    "config2x4": lambda: rb.scenes.config2(n_worlds=16384),
    "config4": lambda: rb.scenes.config4(n_worlds=131072),
    "config3": lambda: rb.scenes.config3(n_worlds=65536),
    "config1": lambda: rb.scenes.config1(n_worlds=4096, seed=1),
    "config1x1": lambda: rb.scenes.config1(n_worlds=1, seed=1),
    "config5-100-16": lambda: rb.scenes.config5(n_worlds=32768, n_segments=100, n_rays=16),
    "config5-2000-256": lambda: rb.scenes.config5(n_worlds=8192, n_segments=2000, n_rays=256),
}
stream = torch.cuda.current_stream()
rows = {}
for name in args.configs.split(","):
    scene = makers[name]()
    K = 1000 if name == "config1x1" else args.steps
    res = {}
    # a fresh engine per mode, the same number of steps in the same order: both modes time the
    # same world states (the cost of a step drifts as the robots move)
    for mode in ("single", "steps"):
        eng = rb.BatchEngine(scene)
        for _ in range(3):
            eng.step(stream=stream.cuda_stream)
        reps = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            if mode == "single":
                for _ in range(K):
                    eng.step(stream=stream.cuda_stream)
            else:
                eng.steps(K, stream=stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            reps.append(e0.elapsed_time(e1) / K)
        res[mode] = reps
        path = eng.launch_info()["path"]
        eng.close()
    a, b = sorted(res["single"])[2], sorted(res["steps"])[2]
    rows[name] = {"worlds": scene.n_worlds, "steps": K, "ms_per_step_single_launches": a, "ms_per_step_one_launch": b,
                  "reps_single": res["single"], "reps_one_launch": res["steps"], "path": path}
    print("%-18s worlds=%7d K=%4d  single %.5f ms/step   steps() %.5f ms/step   x%.3f  %s | first/last rep single %.4f %.4f steps %.4f %.4f" % (
        name, scene.n_worlds, K, a, b, a / b, path[:12], res["single"][0], res["single"][-1], res["steps"][0], res["steps"][-1]), flush=True)
json.dump(rows, open(args.out, "w"), indent=1)
