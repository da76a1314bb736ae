"""Dev: time the CUDA UnevenMap builder against the host builder on the full 200 x 200 x 64 grid (synthetic 200k-point cloud)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_map_build import _cloud
from uneven_planner_b200 import _lib, api, maps
pts = _cloud(200000, 5.0, 0, hole=False)
geom = _lib.map_geometry()
opt = api.BatchALMTrajOpt()
for rep in range(2):
    t0 = time.perf_counter(); dev, ms = opt.build_map(pts, geom); t1 = time.perf_counter()
print("device: kernel %.1f ms, call %.1f ms (preprocessing on the host + copies)" % (ms, (t1 - t0) * 1e3))
t0 = time.perf_counter(); host = maps.build_from_cloud(pts, geom); t1 = time.perf_counter()
print("host  : %.1f ms on %d threads" % ((t1 - t0) * 1e3, os.cpu_count()))
print("bit-identical:", bool(np.array_equal(host.cells.view(np.uint32), dev.cells.view(np.uint32))))
cells = 200 * 200 * 64
print("cells/s device %.3g, host %.3g" % (cells / (ms * 1e-3), cells / (t1 - t0)))
