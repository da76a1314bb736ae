"""Penalty-phase kernel alone (one evaluation per trajectory), for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uneven_planner_b200 import maps, problems, _lib, api
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = maps.get_terrain("hill") or maps.synthetic_terrain("bumps")
pb = problems.generate(m, B, seed=0)
opt = api.BatchALMTrajOpt().init(_lib.default_params()).set_environment(m)
opt.upload(pb)
print(opt.time_penalty_kernel(1))
