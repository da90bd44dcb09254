"""One-off soak: N random scenes (tests/test_parity_gpu.py::_random_scene) against the oracle.
usage: python tools/soak_random_scenes.py [first_seed] [count]"""
import importlib.util
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
spec = importlib.util.spec_from_file_location("tp", os.path.join(ROOT, "tests", "test_parity_gpu.py"))
tp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tp)
import parity  # noqa: E402

first, count = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 500
t0, ok, skipped = time.time(), 0, 0
tot = {"ranges": 0, "range_hit_flips": 0, "pixels": 0, "pixel_mismatch": 0, "lights": 0, "light_flips": 0}
for seed in range(first, first + count):
    try:
        scene = tp._random_scene(seed)
    except RuntimeError:
        skipped += 1
        continue
    try:
        st = parity.compare(scene, n_steps=4)
    except AssertionError as e:
        print("PARITY VIOLATION at seed", seed, e)
        raise
    for k in tot:
        tot[k] += st[k]
    ok += 1
print("scenes ok %d skipped %d in %.0f s" % (ok, skipped, time.time() - t0), tot)
