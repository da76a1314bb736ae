"""Dev script (GPU box): pipelined throughput of the resident bench batch for warp-group policies x batches in flight."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "one":
    sys.path.insert(0, ROOT)
    from uneven_planner_b200 import maps, problems, _lib, api
    B, depth, steps = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    m = maps.get_terrain("hill")
    pb = problems.generate(m, B, seed=0)
    opt = api.BatchALMTrajOpt().init(_lib.default_params()).set_environment(m)
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
    print(json.dumps({"ms_per_step": ms / steps, "conv_per_s": conv * steps / ms * 1e3, "solved_per_s": B * steps / ms * 1e3}))
    sys.exit(0)
B = sys.argv[1] if len(sys.argv) > 1 else "1024"
steps = sys.argv[2] if len(sys.argv) > 2 else "8"
combos = [({}, d) for d in (1, 2, 3, 4)] + \
         [({"UALM_GROUPS": "0"}, d) for d in (2, 3, 4)] + \
         [({"UALM_F4": "0", "UALM_F2": "1"}, d) for d in (2, 3, 4)] + \
         [({"UALM_F4": "0", "UALM_F2": "0.5"}, 3), ({"UALM_F4": "0.05", "UALM_F2": "0.95"}, 3), ({"UALM_F4": "0", "UALM_F2": "1", "UALM_NOADOPT": "1"}, 3),
          ({"UALM_F4": "0.25", "UALM_F2": "0.75"}, 3)]
for envx, depth in combos:
    env = dict(os.environ); env.update(envx)
    r = subprocess.run([sys.executable, __file__, "one", B, str(depth), steps], env=env, capture_output=True, text=True)
    print(envx, "depth", depth, "->", r.stdout.strip() or r.stderr[-400:], flush=True)
