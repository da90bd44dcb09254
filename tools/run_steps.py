#!/usr/bin/env python
"""Profiling target: 3 single brs_step launches, a brs_steps(2) launch, then ONE brs_steps(K) launch of a config.

    python tools/run_steps.py config2 [K=10] [worlds]
    ncu --set full --clock-control none --import-source on -k regex:brs_ -s 4 -c 1 -o out python tools/run_steps.py config2 5
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import aitk_robots_b200 as rb  # noqa: E402
import bench  # noqa: E402

cfg = sys.argv[1]
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
worlds = int(sys.argv[3]) if len(sys.argv) > 3 else bench.DEFAULT_WORLDS.get(cfg, 32768)
scene = bench.make_scene(rb, cfg, worlds, 0)
eng = rb.BatchEngine(scene)
for _ in range(3):
    eng.step()
eng.steps(2)  # (first call allocates the epoch words)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.steps(K)
e1.record()
torch.cuda.synchronize()
print("%s worlds=%d K=%d launches=%d %.5f ms/step %s" % (cfg, worlds, K, eng.steps_launches(K), e0.elapsed_time(e1) / K,
                                                       eng.launch_info()["path"]))
eng.close()
