"""Dev script (GPU box): throughput path (precision 65 / 32) against the oracle: one evaluation, initScaling, full solves, throughput."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as po
from uneven_planner_b200 import maps, problems, _lib, api

def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 65
what = sys.argv[2] if len(sys.argv) > 2 else "all"
m = maps.get_terrain("hill")
params = _lib.default_params()
K = params.int_K
op, om = po.params_from(params), po.OracleMap(m)
opt = api.BatchALMTrajOpt(precision=prec).init(params).set_environment(m)

if what in ("all", "eval"):
    pb = problems.generate(m, 24, seed=3)
    offx = np.concatenate([[0], np.cumsum(pb.nvar())]); offs = np.concatenate([[0], np.cumsum(pb.nsamples(K))])
    _, _, ocx, ocy = pb.offsets()
    rng = np.random.default_rng(0)
    S = int(offs[-1])
    lam = rng.standard_normal(S) * 0.1
    mu = np.abs(rng.standard_normal(6 * S)) * 0.1 * (rng.random(6 * S) < 0.5)
    scx = rng.uniform(0.01, 1.0, 7 * S)
    sfx = rng.uniform(1e-6, 1e-3, pb.B)
    xs = np.concatenate([pb.x0(i) * (1 + 1e-3 * rng.standard_normal(pb.nvar()[i])) for i in range(pb.B)])
    opt.upload(pb)
    ev = opt.eval_batch(xs, lam, mu, scx, sfx, rho=8.0)
    worst = dict(f=0, grad=0, hx=0, gx=0, c=0)
    for i in range(pb.B):
        o = po.eval_one(op, om, pb, i, xs[offx[i]:offx[i + 1]], lam[offs[i]:offs[i + 1]], mu[6 * offs[i]:6 * offs[i + 1]], scx[7 * offs[i]:7 * offs[i + 1]], sfx[i], 8.0)
        worst["f"] = max(worst["f"], abs(ev["f"][i] - o["f"]) / abs(o["f"]))
        worst["grad"] = max(worst["grad"], rel(ev["grad"][offx[i]:offx[i + 1]], o["grad"]))
        worst["hx"] = max(worst["hx"], rel(ev["hx"][offs[i]:offs[i + 1]], o["hx"]))
        worst["gx"] = max(worst["gx"], rel(ev["gx"][6 * offs[i]:6 * offs[i + 1]], o["gx"]))
        worst["c"] = max(worst["c"], rel(ev["c_xy"][ocx[i]:ocx[i + 1]], o["c_xy"]), rel(ev["c_yaw"][ocy[i]:ocy[i + 1]], o["c_yaw"]))
    print("EVAL prec", prec, "worst relative errors:", worst, flush=True)

if what in ("all", "scale"):
    pb = problems.generate(m, 8, seed=4)
    opt.upload(pb)
    sfx, scx = opt.init_scaling_batch()
    offs = np.concatenate([[0], np.cumsum(pb.nsamples(K))])
    w1 = w2 = 0
    for i in range(pb.B):
        osfx, oscx = po.init_scaling(op, om, pb, i)
        w1 = max(w1, abs(sfx[i] - osfx) / osfx); w2 = max(w2, float(np.max(np.abs(scx[7 * offs[i]:7 * offs[i + 1]] - oscx) / oscx)))
    print("SCALE prec", prec, "worst rel err scale_fx", w1, "scale_cx", w2, flush=True)

if what in ("all", "solve"):
    pb = problems.generate(m, 256, seed=0)
    t = time.time(); res, cxy, cyaw = opt.optimize(pb); dt = time.time() - t
    ores = po.solve_batch(op, om, pb, threads=len(os.sched_getaffinity(0)))
    rc = np.array([r.ret_code for r in res]); orc = np.array([r[0].ret_code for r in ores])
    ev = np.array([r.n_evals for r in res]); oev = np.array([r[0].n_evals for r in ores])
    cost = np.array([r.inner_cost for r in res]); ocost = np.array([r[0].inner_cost for r in ores])
    print("SOLVE prec", prec, "B=256 wall %.3f s; converged gpu %d oracle %d; evals mean gpu %.1f oracle %.1f max gpu %d oracle %d" % (dt, (rc == 0).sum(), (orc == 0).sum(), ev.mean(), oev.mean(), ev.max(), oev.max()))
    both = (rc == 0) & (orc == 0)
    rel_c = np.abs(cost[both] - ocost[both]) / np.abs(ocost[both])
    print("   cost rel diff percentiles 10/50/90/99:", np.percentile(rel_c, [10, 50, 90, 99]), "within 1e-5: %.3f 1e-4: %.3f 1e-3: %.3f" % ((rel_c < 1e-5).mean(), (rel_c < 1e-4).mean(), (rel_c < 1e-3).mean()), flush=True)
    feas = opt.feasibility(0.01)
    print("   feasibility sample", feas[0], flush=True)

if what in ("all", "perf"):
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    pb = problems.generate(m, B, seed=0)
    for depth in (1, 2, 3, 4):
        for l in range(depth):
            opt.select_lane(l); opt.upload(pb)
        def run(n):
            for s in range(n):
                l = s % depth
                opt.select_lane(l)
                if s >= depth: opt.sync()
                if s == 0: opt.mark_begin()
                opt.solve_resident()
            for l in range(depth):
                opt.select_lane(l); opt.sync()
            return opt.mark_end()
        run(depth)
        steps = 8
        ms = run(steps)
        opt.select_lane(0)
        res, _, _ = opt.download()
        conv = sum(1 for r in res if r.ret_code == 0)
        print("PERF prec %d B=%d depth %d: %.1f ms/step, %.0f conv/s, %.0f solved/s (conv %d)" % (prec, B, depth, ms / steps, conv * steps / ms * 1e3, B * steps / ms * 1e3, conv), flush=True)
    opt.select_lane(0); opt.upload(pb)
    for tma in (1, 0):
        os.environ["UALM_TP_TMA"] = "1" if tma else "0"
        pms, pbytes = opt.time_penalty_kernel(10)
        print("PENALTY prec %d B=%d tma=%d: %.4f ms/launch, %.1f MB algorithmic, %.1f GB/s" % (prec, B, tma, pms, pbytes / 1e6, pbytes / pms / 1e6), flush=True)
opt.close()
