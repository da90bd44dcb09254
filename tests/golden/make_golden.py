# -*- coding: utf-8 -*-
"""Generates the golden fixtures in this directory from the oracle restatement.

NOTE (parity unpinned): the real `aitk.robots` is not importable in the build
container, so these vectors pin the *restatement*, not the
reference.  If `import aitk.robots` ever succeeds, regenerate with
`--engine aitk` (tests/test_real_aitk.py does the comparison live).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import aitk_robots_b200 as rb  # noqa: E402
from oracle import scene_adapter as sa  # noqa: E402


def make(name, generator, kwargs, steps, worlds):
    scene = getattr(rb.scenes, generator)(**kwargs)
    out = {"generator": generator, "kwargs": kwargs, "steps": steps, "worlds": []}
    for i in worlds:
        obs = sa.step_and_observe(scene, i, steps, camera=False)
        obs["world"] = i
        out["worlds"].append(obs)
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    make("config1_1000steps.json", "config1", {"n_worlds": 4, "seed": 1}, 1000, [0, 1, 2, 3])
    make("demo_3steps.json", "demo_scene", {"n_worlds": 8, "seed": 7}, 3, [0, 3, 7])
