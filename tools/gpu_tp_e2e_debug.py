"""Dev script (GPU box): submit / wait pipeline of the throughput path on a terrain, with timing per call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uneven_planner_b200 import configs, maps, problems, api
terrain, B, depth, nsteps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
m = maps.get_terrain(terrain)
params = configs.params_for(terrain)
pb = problems.generate(m, B, seed=0, **configs.gen_kwargs(terrain))
opt = api.BatchALMTrajOpt(precision=32).init(params).set_environment(m)
t0 = time.perf_counter()
def T(): return "%.2f s" % (time.perf_counter() - t0)
tickets = []
for s in range(nsteps):
    if s >= depth:
        r, _, _ = opt.wait(tickets[s - depth]); print(T(), "waited", s - depth, "conv", sum(1 for q in r if q.ret_code == 0), "max evals", max(q.n_evals for q in r), flush=True)
    tickets.append(opt.submit(pb, depth=depth)); print(T(), "submitted", s, flush=True)
for s in range(max(0, nsteps - depth), nsteps):
    r, _, _ = opt.wait(tickets[s]); print(T(), "waited", s, "max evals", max(q.n_evals for q in r), flush=True)
opt.close()
