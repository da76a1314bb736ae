"""ctypes wrapper of oracle/liboracle.so -- TEST INFRASTRUCTURE (see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


class OParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("rho_T", "rho_ter", "max_vel", "max_acc_lon", "max_acc_lat", "max_kap", "min_cxi", "max_sig")] + \
               [("use_scaling", C.c_int)] + \
               [(n, C.c_double) for n in ("rho", "beta", "gamma", "epsilon_con", "max_iter",
                                          "g_epsilon", "min_step", "inner_max_iter", "delta")] + \
               [("mem_size", C.c_int), ("past", C.c_int), ("int_K", C.c_int), ("gravity", C.c_double)]


class OMap(C.Structure):
    _fields_ = [("cells", C.POINTER(C.c_double)), ("voxel_num", C.c_int * 3), ("origin", C.c_double * 3),
                ("max_boundary", C.c_double * 3), ("xy_resolution", C.c_double), ("yaw_resolution", C.c_double)]


class OProblem(C.Structure):
    _fields_ = [("N", C.c_int), ("M", C.c_int)] + [(n, C.POINTER(C.c_double)) for n in
                                                   ("init_xy", "end_xy", "inner_xy", "init_yaw", "end_yaw", "inner_yaw")] + \
               [("total_time", C.c_double)]


class OResult(C.Structure):
    _fields_ = [("ret_code", C.c_int), ("outer_iters", C.c_int), ("n_evals", C.c_int), ("n_lbfgs_iters", C.c_int),
                ("last_lbfgs_ret", C.c_int), ("max_bound", C.c_int)] + \
               [(n, C.c_double) for n in ("inner_cost", "jerk_cost", "total_T", "res_h", "res_g", "scale_fx", "rho_final",
                                          "t_total", "t_minco", "t_penalty", "t_adjoint", "t_lbfgs", "t_scaling")]


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("oracle.cpp", "oracle.h", "Makefile")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.run(["make", "-C", HERE, "-s"] + (["-B"] if force else []), check=True)
    return LIB


_lib = None
dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.orc_solve.argtypes = [C.POINTER(OParams), C.POINTER(OMap), C.POINTER(OProblem), C.POINTER(OResult), dp, dp, dp, dp, dp, dp]
        L.orc_solve_f32.argtypes = [C.POINTER(OParams), C.POINTER(OMap), C.POINTER(OProblem), C.POINTER(OResult), dp, dp, dp]
        L.orc_eval.argtypes = [C.POINTER(OParams), C.POINTER(OMap), C.POINTER(OProblem), dp, dp, dp, dp, C.c_double, C.c_double] + [dp] * 11
        L.orc_init_scaling.argtypes = [C.POINTER(OParams), C.POINTER(OMap), C.POINTER(OProblem), dp, dp, dp]
        L.orc_map_query.argtypes = [C.POINTER(OMap), dp, dp, dp]
        L.orc_map_query.restype = None
        L.orc_minco_generate.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp]
        L.orc_minco_jerk.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp]
        L.orc_minco_jerk.restype = C.c_double
        L.orc_minco_grad_ct_to_qt.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp]
        L.orc_lbfgs_rosenbrock.argtypes = [C.c_int, dp, dp, C.c_int, C.c_double, C.c_int, C.c_double, C.POINTER(C.c_int)]
        for n in ("orc_expC2", "orc_logC2", "orc_dTdtau"):
            getattr(L, n).argtypes = [C.c_double]
            getattr(L, n).restype = C.c_double
        _lib = L
    return _lib


def P(a):
    return None if a is None else a.ctypes.data_as(dp)


def params_from(p):
    """Copy the fields of a product-side params struct (same field names) into the oracle's own struct."""
    o = OParams()
    for name, _ in OParams._fields_:
        setattr(o, name, getattr(p, name))
    return o


class OracleMap:
    """Holds the double grid the oracle reads: the map's own double cells (cells64) when it has them, else the float32 cells widened."""

    def __init__(self, mapdata):
        c64 = getattr(mapdata, "cells64", None)
        self.cells = np.ascontiguousarray(mapdata.cells if c64 is None else c64, dtype=np.float64)
        g = mapdata.geom
        self.c = OMap()
        self.c.cells = self.cells.ctypes.data_as(dp)
        for k in range(3):
            self.c.voxel_num[k] = g.voxel_num[k]
            self.c.origin[k] = g.origin[k]
            self.c.max_boundary[k] = g.max_boundary[k]
        self.c.xy_resolution = g.xy_resolution
        self.c.yaw_resolution = g.yaw_resolution

    def query(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        v = np.zeros(7)
        g = np.zeros((7, 3))
        lib().orc_map_query(C.byref(self.c), P(pos), P(v), P(g))
        return v, g


def _problem(pb, i, keep):
    oxy, oyaw, _, _ = pb.offsets()
    bnd = np.ascontiguousarray(pb.bnd[i])
    ixy = np.ascontiguousarray(pb.inner_xy[oxy[i]:oxy[i + 1]])
    iyaw = np.ascontiguousarray(pb.inner_yaw[oyaw[i]:oyaw[i + 1]])
    keep += [bnd, ixy, iyaw]
    pr = OProblem()
    pr.N = int(pb.N[i]); pr.M = int(pb.M[i])
    base = bnd.ctypes.data
    pr.init_xy = C.cast(base, dp); pr.end_xy = C.cast(base + 6 * 8, dp)
    pr.init_yaw = C.cast(base + 12 * 8, dp); pr.end_yaw = C.cast(base + 15 * 8, dp)
    pr.inner_xy = P(ixy) if ixy.size else C.cast(base, dp)
    pr.inner_yaw = P(iyaw) if iyaw.size else C.cast(base, dp)
    pr.total_time = float(pb.total_time[i])
    return pr


def solve_one(params, omap, pb, i, f32=False, want_duals=False):
    keep = []
    pr = _problem(pb, i, keep)
    N, M = int(pb.N[i]), int(pb.M[i])
    S = N * (params.int_K + 1)
    res = OResult()
    cxy = np.zeros(12 * N); cyaw = np.zeros(6 * M); x = np.zeros(1 + 2 * (N - 1) + (M - 1))
    if f32:
        lib().orc_solve_f32(C.byref(params), C.byref(omap.c), C.byref(pr), C.byref(res), P(cxy), P(cyaw), P(x))
        return res, cxy, cyaw, x
    lam = np.zeros(S) if want_duals else None
    mu = np.zeros(6 * S) if want_duals else None
    sc = np.zeros(7 * S) if want_duals else None
    lib().orc_solve(C.byref(params), C.byref(omap.c), C.byref(pr), C.byref(res), P(cxy), P(cyaw), P(x), P(lam), P(mu), P(sc))
    if want_duals:
        return res, cxy, cyaw, x, lam, mu, sc
    return res, cxy, cyaw, x


def solve_batch(params, omap, pb, threads=1, idx=None, f32=False):
    """Solve problems idx (default all) with `threads` host threads (ctypes releases the GIL).
    Returns list of (OResult, c_xy, c_yaw, x)."""
    idx = list(range(pb.B)) if idx is None else list(idx)
    if threads <= 1:
        return [solve_one(params, omap, pb, i, f32) for i in idx]
    with ThreadPoolExecutor(threads) as ex:
        return list(ex.map(lambda i: solve_one(params, omap, pb, i, f32), idx))


def eval_one(params, omap, pb, i, x, lam=None, mu=None, scale_cx=None, scale_fx=1.0, rho=None):
    keep = []
    pr = _problem(pb, i, keep)
    N, M = int(pb.N[i]), int(pb.M[i])
    S = N * (params.int_K + 1)
    n = 1 + 2 * (N - 1) + (M - 1)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = dict(f=np.zeros(1), grad=np.zeros(n), hx=np.zeros(S), gx=np.zeros(6 * S), parts=np.zeros(3), c_xy=np.zeros(12 * N),
               c_yaw=np.zeros(6 * M), gdCxy=np.zeros(12 * N), gdTxy=np.zeros(N), gdCyaw=np.zeros(6 * M), gdTyaw=np.zeros(M))
    lam = None if lam is None else np.ascontiguousarray(lam, dtype=np.float64)
    mu = None if mu is None else np.ascontiguousarray(mu, dtype=np.float64)
    scale_cx = None if scale_cx is None else np.ascontiguousarray(scale_cx, dtype=np.float64)
    lib().orc_eval(C.byref(params), C.byref(omap.c), C.byref(pr), P(x), P(lam), P(mu), P(scale_cx), float(scale_fx),
                   float(params.rho if rho is None else rho), *[P(out[k]) for k in
                                                                ("f", "grad", "hx", "gx", "parts", "c_xy", "c_yaw", "gdCxy", "gdTxy", "gdCyaw", "gdTyaw")])
    out["f"] = float(out["f"][0])
    return out


def init_scaling(params, omap, pb, i, x0=None):
    keep = []
    pr = _problem(pb, i, keep)
    S = int(pb.N[i]) * (params.int_K + 1)
    x0 = np.ascontiguousarray(pb.x0(i) if x0 is None else x0, dtype=np.float64)
    sfx = np.zeros(1); scx = np.zeros(7 * S)
    lib().orc_init_scaling(C.byref(params), C.byref(omap.c), C.byref(pr), P(x0), P(sfx), P(scx))
    return float(sfx[0]), scx


def feasibility(omap, gravity, N, M, c_xy, c_yaw, T_xy, T_yaw, dt=0.01):
    """orc_feasibility: [max_vx, max_ax, max_ay, max_cur, max_att, max_sig, nonhol_error, samples]."""
    out = np.zeros(8)
    c_xy = np.ascontiguousarray(c_xy, dtype=np.float64); c_yaw = np.ascontiguousarray(c_yaw, dtype=np.float64)
    f = lib().orc_feasibility
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, dp, dp, C.c_double, C.c_double, C.c_double, dp]
    f(C.addressof(omap.c), float(gravity), int(N), int(M), c_xy.ctypes.data_as(dp), c_yaw.ctypes.data_as(dp), float(T_xy), float(T_yaw), float(dt),
      out.ctypes.data_as(dp))
    return out
