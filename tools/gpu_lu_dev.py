"""Dev: per-phase cycles of single trajectories alone on the GPU (no contention) vs in a full batch."""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uneven_planner_b200 import maps, problems, _lib, api
m = maps.get_terrain("hill")
pb = problems.generate(m, 64, seed=1)
opt = api.BatchALMTrajOpt().init(_lib.default_params()).set_environment(m)
names = ["fill", "lu", "solve", "jerk", "tables", "samples", "accumulate", "combine", "adjoint", "tail", "twoloop", "linesearch", "scaling", "dual", "other", "total"]
for idx in [0, 20, 40, 63]:
    one = pb.select([idx])
    opt.upload(one); opt.profile(True)
    for rep in range(2):
        opt.solve_resident(); opt.sync()
    raw = (C.c_longlong * 16)()
    opt.L.ualm_profile(opt.h, 2, raw)
    pr = np.array(raw)
    res, _, _ = opt.download()
    ev = res[0].n_evals
    N, M = int(one.N[0]), int(one.M[0])
    print(f"N={N} M={M} evals={ev} total {pr[15]/ev:.0f} cyc/eval; " + " ".join(f"{n}={pr[i]/ev:.0f}" for i, n in enumerate(names[:15]) if pr[i] / ev > 500),
          f"| lu per pivot {pr[1]/ev/(6*M):.0f} cyc; sweeps per block {(pr[2]+pr[8])/ev/(4*M):.0f}")
