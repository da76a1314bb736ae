import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_map_build import _cloud
from uneven_planner_b200 import _lib, api, maps
geom = _lib.map_geometry(4.0, 4.0, 0.05, 0.1)
opt = api.BatchALMTrajOpt()
for seed, n, hole in [(0, 60000, True), (0, 60000, False), (1, 4000, True)]:
    pts = _cloud(n, 2.3, seed, hole)
    host = maps.build_from_cloud(pts, geom)
    dev, ms = opt.build_map(pts, geom)
    neq = host.cells.view(np.uint32) != dev.cells.view(np.uint32)
    print("seed", seed, "n", n, "hole", hole, "mismatching values", int(neq.sum()), "cells", int(neq.any(axis=-1).sum()), "per field", neq.reshape(-1, 4).sum(axis=0),
          "nan host/dev", int(np.isnan(host.cells).sum()), int(np.isnan(dev.cells).sum()))
    idx = np.argwhere(neq.any(axis=-1))[:6]
    for i in idx:
        print("  cell", tuple(i), "host", host.cells[tuple(i)], "dev", dev.cells[tuple(i)])
