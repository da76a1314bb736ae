import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uneven_planner_b200 import maps, problems, _lib, api
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = maps.get_terrain("hill")
pb = problems.generate(m, B, seed=0)
opt = api.BatchALMTrajOpt().init(_lib.default_params()).set_environment(m)
opt.upload(pb); opt.profile(True)
opt.solve_resident(); opt.sync()
ms, _ = opt.last_solve_ms()
raw = (C.c_longlong * (B * 16))()
opt.L.ualm_profile(opt.h, 2, raw)
pr = np.array(raw).reshape(B, 16)
res, _, _ = opt.download()
tot = pr[:, 15] / 1.965e6  # ms, rows in problem order
N = pb.N; ev = np.array([res[i].n_evals for i in range(B)]); ret = np.array([res[i].ret_code for i in range(B)])
print("kernel ms", ms, "mean traj ms", tot.mean(), "max", tot.max(), "p50 p90 p99", np.percentile(tot, [50, 90, 99]))
idx = np.argsort(-tot)[:12]
for i in idx: print("slot", i, "N", N[i], "evals", ev[i], "ret", ret[i], "ms %.1f" % tot[i], "ms/eval %.3f" % (tot[i] / ev[i]))
A = np.column_stack([ev, ev * N]); coef, *_ = np.linalg.lstsq(A, tot, rcond=None)
print("fit ms = evals*(%.4f + %.5f*N)" % tuple(coef))
print("evals: mean", ev.mean(), "p50 p90 p99 max", np.percentile(ev, [50, 90, 99, 100]))
for lo, hi in [(0, 12), (12, 18), (18, 24), (24, 30), (30, 60)]:
    s = (N >= lo) & (N < hi)
    print("N in [%d,%d): count %d mean ms %.1f max ms %.1f mean evals %.0f" % (lo, hi, s.sum(), tot[s].mean() if s.any() else 0, tot[s].max() if s.any() else 0, ev[s].mean() if s.any() else 0))
