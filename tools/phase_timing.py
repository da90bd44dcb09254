"""Phase shares of the step kernel (needs a -DBRS_PHASE_TIMING build selected with BRS_LIB).
usage: BRS_LIB=build/var/libbrs_phase.so python tools/phase_timing.py config4 131072"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aitk_robots_b200 as rb

cfg, B = sys.argv[1], int(sys.argv[2])
eng = rb.BatchEngine(getattr(rb.scenes, cfg)(n_worlds=B))
for _ in range(10):
    eng.step()
eng.synchronize()
print(cfg, B, eng.launch_info(), flush=True)
eng.close()  # prints the shares
