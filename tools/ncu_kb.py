"""Workload for ncu (GPU box): the penalty kernel of the throughput path alone over one batch (python tools/ncu_kb.py PREC B [tma])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uneven_planner_b200 import maps, problems, _lib, api
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
if len(sys.argv) > 3 and sys.argv[3] == "tma":
    os.environ["UALM_TP_TMA"] = "1"
m = maps.get_terrain("hill")
pb = problems.generate(m, B, seed=0)
opt = api.BatchALMTrajOpt(precision=prec).init(_lib.default_params()).set_environment(m)
opt.upload(pb)
ms, by = opt.time_penalty_kernel(5)
print("kb_kernel B=%d prec=%d: %.4f ms/launch, %.1f MB algorithmic -> %.1f GB/s" % (B, prec, ms, by / 1e6, by / ms / 1e6))
opt.close()
