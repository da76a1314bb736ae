"""Dev script (GPU box): in-kernel phase profile of one batch solve."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uneven_planner_b200 import maps, problems, _lib, api
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = maps.get_terrain("hill") or maps.synthetic_terrain("bumps")
pb = problems.generate(m, B, seed=0)
opt = api.BatchALMTrajOpt().init(_lib.default_params()).set_environment(m)
opt.upload(pb)
opt.profile(True)
for rep in range(2):
    opt.solve_resident(); opt.sync()
ms, _ = opt.last_solve_ms()
pr = opt.profile(True, read=True)
res, _, _ = opt.download()
ev = sum(r.n_evals for r in res); it = sum(r.n_lbfgs_iters for r in res)
tot = pr["total"]
print(f"B={B} kernel {ms:.1f} ms; evals {ev} iters {it}; mean cycles/traj {tot/B:.3e}; cycles/eval {tot/ev:.0f}")
for k, v in pr.items():
    print(f"  {k:12s} {100.0*v/tot:6.2f}%   {v/ev:10.0f} cyc/eval")
