import sys; sys.path.insert(0,'.')
import aitk_robots_b200 as rb
for nb in (128<<20, 512<<20, 2048<<20):
    for mode in (0,1):
        print(nb>>20, "MB mode", mode, "%.0f GB/s" % rb.measure_write_bw(0, nb, mode, 5))
