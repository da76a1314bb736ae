"""The bench workload (1024 problems, seed 0, hill map) once through the resident path + the penalty kernel, for ncu traffic captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uneven_planner_b200 import maps, problems, _lib, api
m = maps.get_terrain("hill")
pb = problems.generate(m, 1024, seed=0)
opt = api.BatchALMTrajOpt().init(_lib.default_params()).set_environment(m)
opt.upload(pb)
opt.solve_resident(); opt.sync()
print("solve ms", opt.last_solve_ms())
print("penalty", opt.time_penalty_kernel(1))
