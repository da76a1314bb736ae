"""Dev script (GPU box): compare the CUDA path with the oracle bit by bit, then time a batch."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from uneven_planner_b200 import maps, problems, _lib, api
import pyoracle as po

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
BT = int(sys.argv[2]) if len(sys.argv) > 2 else 256
m = maps.get_terrain("hill") or maps.synthetic_terrain("bumps")
print("map", m.name)
pb = problems.generate(m, B, seed=0)
params = _lib.default_params()
op = po.params_from(params); om = po.OracleMap(m)
opt = api.BatchALMTrajOpt().init(params).set_environment(m)
opt.upload(pb)
K = params.int_K
offx = np.concatenate([[0], np.cumsum(pb.nvar())]); offs = np.concatenate([[0], np.cumsum(pb.nsamples(K))])
_, _, ocx, ocy = pb.offsets()

def ulpdiff(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return int((a != b).sum()), float(np.max(np.abs(a - b) / (np.abs(b) + 1e-300))) if a.size else 0.0

# 1. eval at x0 with zero duals
t = time.time(); ev = opt.eval_batch(); print("gpu eval s", time.time() - t)
bad = 0
for i in range(B):
    o = po.eval_one(op, om, pb, i, pb.x0(i))
    d = [ulpdiff(ev["f"][i], o["f"]), ulpdiff(ev["grad"][offx[i]:offx[i+1]], o["grad"]), ulpdiff(ev["hx"][offs[i]:offs[i+1]], o["hx"]),
         ulpdiff(ev["gx"][6*offs[i]:6*offs[i+1]], o["gx"]), ulpdiff(ev["c_xy"][ocx[i]:ocx[i+1]], o["c_xy"]), ulpdiff(ev["c_yaw"][ocy[i]:ocy[i+1]], o["c_yaw"])]
    if any(x[0] for x in d):
        bad += 1
        if bad <= 5: print("eval mismatch prob", i, "N", pb.N[i], "M", pb.M[i], "[f,grad,hx,gx,cxy,cyaw] (#diff, maxrel):", d)
print("EVAL x0: problems with any bit difference:", bad, "/", B)

# 2. eval with random duals / scales
rng = np.random.default_rng(0)
S = int(offs[-1])
lam = rng.standard_normal(S) * 0.1; mu = np.abs(rng.standard_normal(6 * S)) * 0.1 * (rng.random(6 * S) < 0.5); scx = rng.uniform(0.01, 1.0, 7 * S)
sfx = rng.uniform(1e-6, 1e-3, B)
xs = np.concatenate([pb.x0(i) * (1 + 1e-3 * rng.standard_normal(pb.nvar()[i])) for i in range(B)])
ev = opt.eval_batch(xs, lam, mu, scx, sfx, rho=8.0)
bad = 0
for i in range(B):
    o = po.eval_one(op, om, pb, i, xs[offx[i]:offx[i+1]], lam[offs[i]:offs[i+1]], mu[6*offs[i]:6*offs[i+1]], scx[7*offs[i]:7*offs[i+1]], sfx[i], 8.0)
    d = [ulpdiff(ev["f"][i], o["f"]), ulpdiff(ev["grad"][offx[i]:offx[i+1]], o["grad"]), ulpdiff(ev["hx"][offs[i]:offs[i+1]], o["hx"]), ulpdiff(ev["gx"][6*offs[i]:6*offs[i+1]], o["gx"])]
    if any(x[0] for x in d):
        bad += 1
        if bad <= 5: print("eval2 mismatch prob", i, d)
print("EVAL duals: problems with any bit difference:", bad, "/", B)

# 3. initScaling
opt.upload(pb)
t = time.time(); gsfx, gscx = opt.init_scaling_batch(); print("gpu scaling s", time.time() - t)
bad = 0
for i in range(min(B, 16)):
    osfx, oscx = po.init_scaling(op, om, pb, i)
    d = [ulpdiff(gsfx[i], osfx), ulpdiff(gscx[7*offs[i]:7*offs[i+1]], oscx)]
    if any(x[0] for x in d):
        bad += 1
        if bad <= 5: print("scaling mismatch prob", i, d)
print("SCALING: problems with any bit difference:", bad, "/", min(B, 16))

# 4. full solve
t = time.time(); res, cxy, cyaw = opt.optimize(pb); tg = time.time() - t
ms, _ = opt.last_solve_ms()
print("gpu solve wall s", tg, "kernel ms", ms)
t = time.time(); ores = po.solve_batch(op, om, pb, threads=os.cpu_count()); tc = time.time() - t
print("oracle wall s", tc, "threads", os.cpu_count())
bad = 0; worst = 0
for i in range(B):
    r, ocxy, ocyaw, ox = ores[i]
    g = res[i]
    same = (g.ret_code == r.ret_code and g.n_evals == r.n_evals and g.outer_iters == r.outer_iters and g.inner_cost == r.inner_cost)
    dc = ulpdiff(cxy[ocx[i]:ocx[i+1]], ocxy); dy = ulpdiff(cyaw[ocy[i]:ocy[i+1]], ocyaw)
    worst = max(worst, dc[1], dy[1])
    if not same or dc[0] or dy[0]:
        bad += 1
        if bad <= 8: print("solve mismatch", i, "gpu", g.ret_code, g.outer_iters, g.n_evals, g.inner_cost, "orc", r.ret_code, r.outer_iters, r.n_evals, r.inner_cost, dc, dy)
print("SOLVE: problems with any difference:", bad, "/", B, "worst rel coef diff", worst)
conv = sum(1 for i in range(B) if res[i].ret_code == 0)
print("converged", conv, "/", B)

# 5. timing at larger batch
pb2 = problems.generate(m, BT, seed=1)
opt.upload(pb2)
for rep in range(3):
    opt.solve_resident(); opt.sync(); ms, _ = opt.last_solve_ms()
    res2, _, _ = opt.download()
    conv2 = sum(1 for i in range(BT) if res2[i].ret_code == 0)
    ev2 = sum(res2[i].n_evals for i in range(BT))
    print(f"B={BT} kernel ms {ms:.1f} -> {BT/ms*1e3:.0f} traj/s ({conv2/ms*1e3:.0f} converged/s), evals {ev2}, {ms*1e3/ev2*BT:.1f} us/eval/traj-slot")
pms, pbytes = opt.time_penalty_kernel(5)
print("penalty kernel ms/launch", pms, "alg bytes", pbytes, "GB/s", pbytes / pms / 1e6)
