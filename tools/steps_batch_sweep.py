#!/usr/bin/env python
"""config 2 under brs_steps: ns per world-step against batch size and steps per launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import aitk_robots_b200 as rb  # noqa: E402

stream = torch.cuda.current_stream()
for B in (1024, 2048, 4096, 8192, 16384, 32768):
    scene = rb.scenes.config2(n_worlds=B)
    for K in (20, 100):
        eng = rb.BatchEngine(scene)
        eng.steps(3, stream=stream.cuda_stream)
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            eng.steps(K, stream=stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / K)
        ms = sorted(ts)[1]
        print("B=%6d K=%4d ms/step %.5f ns/world %.2f GB/s %.0f" % (B, K, ms, ms * 1e6 / B, B * 131584 / (ms * 1e-3) / 1e9), flush=True)
        eng.close()
