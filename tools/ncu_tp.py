"""Workload for ncu (GPU box): one batch through the throughput path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uneven_planner_b200 import maps, problems, _lib, api
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
m = maps.get_terrain("hill")
pb = problems.generate(m, B, seed=0)
opt = api.BatchALMTrajOpt(precision=prec).init(_lib.default_params()).set_environment(m)
res, _, _ = opt.optimize(pb)
print("converged", sum(1 for r in res if r.ret_code == 0))
opt.close()
