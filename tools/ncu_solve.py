"""One solve of a small batch, for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uneven_planner_b200 import maps, problems, _lib, api
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
m = maps.get_terrain("hill") or maps.synthetic_terrain("bumps")
pb = problems.generate(m, B, seed=1)
opt = api.BatchALMTrajOpt().init(_lib.default_params()).set_environment(m)
opt.upload(pb)
opt.solve_resident(); opt.sync()
ms, _ = opt.last_solve_ms()
print("B", B, "ms", ms)
