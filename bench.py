#!/usr/bin/env python
"""bench.py -- converged trajectories/sec of the batched MINCO/ALM optimizer (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores (oracle)

A "step" is one pass of the hot path (B independent optimizeSE2Traj solves) over one batch of synthetic problems:
BASELINE.json configs[1], B = 1024 random SE(2) start/goal pairs per GPU on the hill UnevenMap (weak scaling: config 3's
8192 problems on 8 GPUs is the same 1024 per GPU).  `value` is measured with the problems resident in HBM; `e2e` goes
through the host-buffer C-ABI call (ualm_solve_batch) with pinned host inputs, H2D and D2H inside the timed region.
One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "converged trajectories/sec (batch MINCO)"
UNIT = "traj/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this round."""
    p = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(p):
        return json.load(open(p))
    return {}


def get_map(name):
    from uneven_planner_b200 import maps
    m = maps.get_terrain(name)
    if m is not None:
        return m, name
    # no .umap travelled: clearly labelled analytic stand-in (not one of the reference's terrains)
    return maps.synthetic_terrain("bumps", seed=0), "synthetic-bumps (maps_built/%s.umap missing)" % name


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def algorithmic_bytes(pb, res, K, e=8):
    """SURVEY 8d / BASELINE.md section 2: bytes the launch must move, from the work it actually did."""
    N = pb.N.astype(np.float64); M = pb.M.astype(np.float64)
    n = 1 + 2 * (N - 1) + (M - 1); S = N * (K + 1)
    evals = np.array([r.n_evals for r in res], dtype=np.float64)
    iters = np.array([r.n_lbfgs_iters for r in res], dtype=np.float64)
    sumb = np.array([r.sum_bound for r in res], dtype=np.float64)
    pen = evals * (S * 45 * e + (25 * N + 13 * M) * e)
    lb = (4 * n * sumb + 8 * n * iters) * e
    minco = evals * (n + 2 * (12 * N + 6 * M) + (N + M)) * e
    return float(pen.sum()), float(lb.sum()), float(minco.sum())


def time_reference_build(m, params, pb, threads, nprob=16):
    """oracle/_ref/libref.so = the reference's own alm_traj_opt.cpp compiled against oracle/shim: timed on a few
    problems for the record.  It is bit-identical to the oracle port but several times slower (the shim evaluates every Eigen
    expression into heap temporaries), so the port stays the quoted CPU baseline."""
    import ctypes as C
    path = os.path.join(ROOT, "oracle", "_ref", "libref.so")
    if not os.path.exists(path):
        return {"unavailable": "oracle/_ref/libref.so not built (needs the reference sources at build time)"}
    from concurrent.futures import ThreadPoolExecutor
    L = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    L.ref_map_create.restype = C.c_void_p
    L.ref_alm_solve_h.restype = C.c_int
    L.ref_map_destroy.argtypes = [C.c_void_p]

    class RefParams(C.Structure):
        _fields_ = [(n, C.c_double) for n in ("rho_T", "rho_ter", "max_vel", "max_acc_lon", "max_acc_lat", "max_kap", "min_cxi", "max_sig")] + \
                   [("use_scaling", C.c_int)] + \
                   [(n, C.c_double) for n in ("rho", "beta", "gamma", "epsilon_con", "max_iter", "g_epsilon", "min_step", "inner_max_iter", "delta")] + \
                   [("mem_size", C.c_int), ("past", C.c_int), ("int_K", C.c_int), ("gravity", C.c_double)]
    rp = RefParams()
    for n, _ in RefParams._fields_:
        setattr(rp, n, getattr(params, n))
    g = m.geom
    cells = np.ascontiguousarray(m.cells, dtype=np.float64)
    vn = (C.c_int * 3)(*g.voxel_num); org = (C.c_double * 3)(*g.origin); mxb = (C.c_double * 3)(*g.max_boundary)
    nprob = min(nprob, pb.B)
    oxy, oyaw, _, _ = pb.offsets()
    # the reference prints from initScaling / the ALM loop: keep this process's stdout (one JSON line) clean
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        h = C.c_void_p(L.ref_map_create(cells.ctypes.data_as(dp), vn, org, mxb, C.c_double(g.xy_resolution), C.c_double(g.yaw_resolution),
                                        C.c_double(params.gravity)))

        def one(i):
            N, M = int(pb.N[i]), int(pb.M[i])
            S = N * (params.int_K + 1)
            ixy = np.ascontiguousarray(pb.inner_xy[oxy[i]:oxy[i + 1]]); iyaw = np.ascontiguousarray(pb.inner_yaw[oyaw[i]:oyaw[i + 1]])
            bufs = [np.zeros(k) for k in (12 * N, 6 * M, 2, S, 6 * S, S, 6 * S, 1, 7 * S, 1, 7)]
            bnd = np.ascontiguousarray(pb.bnd[i])
            return L.ref_alm_solve_h(C.byref(rp), h, N, M, bnd.ctypes.data_as(dp), C.c_double(float(pb.total_time[i])),
                                     (ixy if ixy.size else np.zeros(1)).ctypes.data_as(dp), (iyaw if iyaw.size else np.zeros(1)).ctypes.data_as(dp),
                                     *[b.ctypes.data_as(dp) for b in bufs])
        t0 = time.perf_counter()
        with ThreadPoolExecutor(min(threads, nprob)) as ex:
            rets = list(ex.map(one, range(nprob)))
        dt = time.perf_counter() - t0
        L.ref_map_destroy(h)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved); os.close(devnull)
    return {"kind": "reference sources (alm_traj_opt.cpp, unmodified) against oracle/shim", "problems": nprob, "threads": min(threads, nprob),
            "converged_per_s": sum(1 for r in rets if r == 0) / dt, "seconds_per_trajectory_per_thread": dt / max(1, -(-nprob // min(threads, nprob))),
            "note": "bit-identical to the oracle port (tests/test_ref_pin.py); slower only because the Eigen stand-in is eager"}


def run_reference(args):
    """--impl reference: the reference algorithm (CPU oracle, oracle/oracle.cpp) on all host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    from uneven_planner_b200 import _lib, problems
    po.build()
    m, mname = get_map(args.map)
    params = _lib.default_params()
    threads = os.cpu_count() or 1
    sample = min(args.ref_sample, args.batch * max(args.gpus, 1))
    pb = problems.generate(m, args.batch * max(args.gpus, 1), seed=args.seed).select(np.arange(sample))
    op, om = po.params_from(params), po.OracleMap(m)
    for _ in range(min(args.warmup, 1)):
        po.solve_batch(op, om, pb.select(np.arange(min(threads, sample))), threads=threads)
    t0 = time.perf_counter()
    conv = 0
    for _ in range(args.steps):
        out = po.solve_batch(op, om, pb, threads=threads)
        conv += sum(1 for r in out if r[0].ret_code == 0)
    dt = time.perf_counter() - t0
    val = conv / dt
    ref_build = time_reference_build(m, params, pb, threads)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "configs[1]: random SE(2) start/goal pairs on %s UnevenMap, run_hill.yaml parameters" % mname,
                       "batch_per_step": sample, "note": "bounded sample of the batch per step, all host threads"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d of the %d problems of the workload per step, %d host threads (one optimizer per thread)" % (sample, args.batch * max(args.gpus, 1), threads)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "reference_build": ref_build}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=1024, help="problems per GPU per step")
    ap.add_argument("--map", default="hill")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--ref-sample", type=int, default=512, dest="ref_sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--big-batch", type=int, default=4096, dest="big_batch", help="extra single-launch throughput datapoint (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from uneven_planner_b200 import _lib, api, problems, distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    m, mname = get_map(args.map)
    params = _lib.default_params()
    K = params.int_K
    Btot = args.batch * world
    pb_all = problems.generate(m, Btot, seed=args.seed)          # identical on every rank (counter-based RNG)
    shards = D.shard_indices(pb_all.nsamples(K), world)
    pb = pb_all.select(shards[rank])
    stride = D.record_stride(pb_all.N.max(), pb_all.M.max())

    opt = api.BatchALMTrajOpt(device=local_rank).init(params).set_environment(m)
    opt.set_stream(torch.cuda.current_stream().cuda_stream)
    records = torch.zeros((pb.B, stride), dtype=torch.float64, device=dev)

    def step_resident():
        opt.solve_resident()
        opt.pack_records(records.data_ptr(), stride)
        return D.all_gather_records(records, shards, rank, world) if world > 1 else records

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM ----------------
    opt.upload(pb)
    for _ in range(args.warmup):
        full = step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    e0.record()
    for _ in range(args.steps):
        full = step_resident()
        kernel_ms.append(None)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    solve_ms, launches = opt.last_solve_ms()                       # CUDA events around the last solve kernel on its stream
    res, cxy, cyaw = opt.download()
    conv_local = sum(1 for r in res if r.ret_code == 0)
    # independent quality check of the solved batch (outside the timed region): the reference's post-solve scan
    # (getMaxVxAxAyCurAttSig + getNonHolError, 0.01 s sampling) on the GPU, rank-local
    feas = opt.feasibility(0.01)
    okc = np.array([r.ret_code == 0 for r in res])
    tol = 1.05
    within = (np.abs(feas[:, 0]) <= params.max_vel * tol) & (np.abs(feas[:, 1]) <= params.max_acc_lon * tol) & \
             (np.abs(feas[:, 2]) <= params.max_acc_lat * tol) & (np.abs(feas[:, 3]) <= params.max_kap * tol) & \
             (-feas[:, 4] >= params.min_cxi / tol) & (feas[:, 5] <= params.max_sig * tol)
    quality = {"converged": int(okc.sum()), "converged_and_within_limits": int((okc & within).sum()),
               "limits": "max |vx|, |ax|, |ay|, |curvature|, sigma <= 1.05 x limit and min cos(xi) >= limit / 1.05 over 0.01 s samples "
                         "(ualm_feasibility_batch; rank 0's shard)",
               "median_nonholonomic_error_per_sample": float(np.median(feas[okc, 6] / np.maximum(feas[okc, 7], 1.0))) if okc.any() else None}
    full_h = full.cpu().numpy()
    conv_total = int((full_h[:, 0] == 0).sum()) if world > 1 else conv_local
    value = conv_total * args.steps / (ms * 1e-3)

    # ---------------- e2e: host buffers through the C-ABI call, H2D + D2H inside the timed region ----------------
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    keep = [pinned(a) for a in (pb.N.astype(np.int32), pb.M.astype(np.int32), pb.bnd, pb.total_time, pb.inner_xy, pb.inner_yaw)]
    hN, hM, hbnd, hT, hxy, hyaw = [k[1] for k in keep]
    out_res = (api.Result * pb.B)()
    o_cxy_t, o_cxy = pinned(np.zeros(int(12 * pb.N.astype(np.int64).sum())))
    o_cyaw_t, o_cyaw = pinned(np.zeros(int(6 * pb.M.astype(np.int64).sum())))
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)

    def step_e2e():
        rc = opt.L.ualm_solve_batch(opt.h, pb.B, hN.ctypes.data_as(ip), hM.ctypes.data_as(ip), hbnd.ctypes.data_as(dp), hT.ctypes.data_as(dp),
                                    hxy.ctypes.data_as(dp), hyaw.ctypes.data_as(dp), out_res, o_cxy.ctypes.data_as(dp), o_cyaw.ctypes.data_as(dp))
        if rc != 0:
            raise RuntimeError(opt.L.ualm_last_error())
    e2e_steps = max(1, min(args.steps, 3))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = conv_total * e2e_steps / float(te.item())
    h2d = int(sum(a.nbytes for a in (hN, hM, hbnd, hT, hxy, hyaw)))
    d2h = int(C.sizeof(api.Result) * pb.B + o_cxy.nbytes + o_cyaw.nbytes)

    # ---------------- roofline of the dominant kernel (solve_kernel) ----------------
    peak, peak_src = load_peaks()
    pen_b, lb_b, mc_b = algorithmic_bytes(pb, res, K)
    alg = pen_b + lb_b + mc_b
    achieved = alg / (solve_ms * 1e-3) / 1e9
    traffic = load_traffic()
    roof = {"kernel": "ualm::solve_kernel (whole ALM/L-BFGS solve of the batch: a warp group per trajectory, one launch per size class on concurrent streams)", "bound": "hbm", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic.get("solve_kernel_bytes_per_launch_b1024") if args.batch == 1024 else None,
            "peak_source": peak_src, "kernel_ms": solve_ms,
            "algorithmic_bytes": {"penalty": pen_b, "lbfgs": lb_b, "minco_io": mc_b},
            "note": "latency-bound: bit-reproducible fp64 dependent chains, a warp group per trajectory (DESIGN.md section 4)"}
    pms, pbytes = opt.time_penalty_kernel(5)
    roof_pen = {"kernel": "ualm::penalty_only_kernel (calConstrainCostGrad samples + accumulation, 1 evaluation per trajectory)",
                "bound": "hbm", "achieved": pbytes / (pms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": pbytes / (pms * 1e-3) / 1e9 / peak,
                "traffic": traffic.get("penalty_only_kernel_bytes_per_launch_b1024") if args.batch == 1024 else None, "kernel_ms": pms}

    # ---------------- throughput with the GPU kept full: 4x the batch in one launch (single GPU only, outside the timed region)
    big = None
    if world == 1 and args.big_batch > 0:
        pbb = problems.generate(m, args.big_batch, seed=args.seed + 1)
        opt.upload(pbb)
        opt.solve_resident(); opt.sync()
        bms, _ = opt.last_solve_ms()
        rb, _, _ = opt.download()
        big = {"batch": args.big_batch, "kernel_ms": bms, "solved_per_s": args.big_batch / bms * 1e3,
               "converged_per_s": sum(1 for r in rb if r.ret_code == 0) / bms * 1e3}
        opt.upload(pb)

    # ---------------- CPU baseline on this box's host cores (rank 0, bounded sample) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle as po
        po.build()
        threads = os.cpu_count() or 1
        sample = min(args.ref_sample, pb_all.B)
        sub = pb_all.select(np.arange(sample))
        t0 = time.perf_counter()
        out = po.solve_batch(po.params_from(params), po.OracleMap(m), sub, threads=threads)
        dt = time.perf_counter() - t0
        cpu_conv = sum(1 for r in out if r[0].ret_code == 0)
        one = np.mean([r[0].t_total for r in out])
        # the reference's own operating mode: one optimizer on one otherwise idle core ("1 thread, per-trajectory ms")
        n1 = min(8, sample)
        t1 = time.perf_counter()
        po.solve_batch(po.params_from(params), po.OracleMap(m), pb_all.select(np.arange(n1)), threads=1)
        one_idle = (time.perf_counter() - t1) / n1
        cpu = {"value": cpu_conv / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "single_thread_ms_per_trajectory": one_idle * 1e3,
               "sample": "first %d problems of the workload, %d host threads (one optimizer instance per thread); mean %.1f ms/trajectory inside a "
                         "thread under that load, %.1f ms/trajectory for the first %d problems on one thread of the idle host" % (sample, threads, one * 1e3, one_idle * 1e3, n1)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "configs[1]: batch=%d random SE(2) start/goal pairs per GPU on %s UnevenMap, run_hill.yaml parameters" % (args.batch, mname),
                           "global_batch": Btot, "parallelism": "dp%d (independent shards, final NCCL all-gather of result records)" % world,
                           "converged_per_step": conv_total, "solved_per_step": Btot,
                           "l2": "per-step working set (L-BFGS history + sample scratch, %.0f MB) exceeds the 126 MB L2; the 41 MB map is reused within a step" % ((lb_b and (8.0 * 2 * params.mem_size * float(pb.nvar().sum())) / 1e6))},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
                "gpu_launches": int((launches + 1) * args.steps),
                "clocks": clocks, "roofline": roof, "roofline_penalty": roof_pen, "cpu_baseline": cpu, "large_batch": big, "quality": quality,
                "work": {"evals_per_step": int(sum(r.n_evals for r in res)), "lbfgs_iters_per_step": int(sum(r.n_lbfgs_iters for r in res)), "rank0_batch": pb.B}}
        print(json.dumps(line))
    opt.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
