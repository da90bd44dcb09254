"""Small multi-feature run for compute-sanitizer: demo scene (3 robots, every device kind) + config4 slice."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import aitk_robots_b200 as rb

for scene in (rb.scenes.demo_scene(40), rb.scenes.config4(n_worlds=64), rb.scenes.config3(n_worlds=48),
              rb.scenes.config5(n_worlds=8, n_segments=300, n_rays=16)):
    eng = rb.BatchEngine(scene)
    for _ in range(3):
        eng.step()
    eng.synchronize()
    print(scene.name, "range sum", float(np.nansum(eng.download("range"))), eng.launch_info())
    eng.close()
print("done")
