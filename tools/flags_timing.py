import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aitk_robots_b200 as rb
cfg = sys.argv[1] if len(sys.argv) > 1 else "config2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
scene = getattr(rb.scenes, cfg)(n_worlds=B)
eng = rb.BatchEngine(scene)
st = torch.cuda.current_stream()
for name, flags in (("all", 7), ("move+camera", 5), ("camera", 4), ("move+sense", 3), ("move", 1), ("all", 7)):
    for _ in range(5):
        eng.step(flags=flags, stream=st.cuda_stream)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(20):
        eng.step(flags=flags, stream=st.cuda_stream)
    b.record(st)
    torch.cuda.synchronize()
    print("%-12s %.4f ms" % (name, a.elapsed_time(b) / 20))
