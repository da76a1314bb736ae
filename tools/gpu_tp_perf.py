"""Dev script (GPU box): throughput path, pipelined: python tools/gpu_tp_perf.py PREC B DEPTH STEPS"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uneven_planner_b200 import maps, problems, _lib, api
prec, B, depth, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
m = maps.get_terrain("hill")
pb = problems.generate(m, B, seed=0)
opt = api.BatchALMTrajOpt(precision=prec).init(_lib.default_params()).set_environment(m)
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
ms = run(steps)
opt.select_lane(0)
res, _, _ = opt.download()
conv = sum(1 for r in res if r.ret_code == 0)
print("PERF prec %d B=%d depth %d: %.1f ms/step, %.0f conv/s, %.0f solved/s (conv %d)" % (prec, B, depth, ms / steps, conv * steps / ms * 1e3, B * steps / ms * 1e3, conv), flush=True)
opt.close()
