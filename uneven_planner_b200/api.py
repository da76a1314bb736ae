"""Python face of the C ABI (include/ualm.h) used by tests and bench.py.

`BatchALMTrajOpt` mirrors the reference's ALMTrajOpt wiring for a batch of problems:
    init(params)            <- ALMTrajOpt::init(nh)                 (alm_traj_opt.cpp:5-45)
    set_environment(map)    <- ALMTrajOpt::setEnvironment(map)      (alm_traj_opt.h:127-130)
    optimize(problems)      <- B x ALMTrajOpt::optimizeSE2Traj(...) (alm_traj_opt.h:92-98), then getTraj()
All compute happens inside libualm.so on the GPU; this file only marshals numpy arrays.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Params, Result

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


def _p(a, t=dp):
    return None if a is None else a.ctypes.data_as(t)


class UalmError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        msg = _lib.lib().ualm_last_error()
        raise UalmError(f"libualm error {rc}: {msg.decode() if msg else ''}")


class BatchALMTrajOpt:
    def __init__(self, device=0, precision=64):
        L = _lib.lib()
        if not hasattr(L, "ualm_create"):
            raise ImportError("libualm.so was built without the CUDA translation unit")
        self.L = L
        self.h = C.c_void_p()
        _check(L.ualm_create(C.byref(self.h), device, precision))
        self.params = None
        self.pb = None

    def close(self):
        if self.h:
            self.L.ualm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self, params=None):
        self.params = params or _lib.default_params()
        _check(self.L.ualm_set_params(self.h, C.byref(self.params)))
        return self

    def set_stream(self, cuda_stream_ptr):
        _check(self.L.ualm_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))
        return self

    def set_environment(self, mapdata, repack_to_float=False):
        """Binds the map grid.  A map that carries the reference's own double grid (`cells64`, RXS2 = 4 doubles per cell) goes
        through ualm_set_map_f64; repack_to_float rounds it to the float4 grid on the way."""
        self.map = mapdata
        c64 = getattr(mapdata, "cells64", None)
        if c64 is not None:
            _check(self.L.ualm_set_map_f64(self.h, C.byref(mapdata.geom), c64.ctypes.data_as(dp), 1 if repack_to_float else 0))
        else:
            _check(self.L.ualm_set_map(self.h, C.byref(mapdata.geom), mapdata.cells.ctypes.data_as(C.POINTER(C.c_float))))
        return self

    def reset_stream(self):
        _check(self.L.ualm_reset_stream(self.h))
        return self

    # ---- lanes: several resident batches in flight on this context ----------------------
    def select_lane(self, lane):
        _check(self.L.ualm_select_lane(self.h, int(lane)))
        self.pb = self._lane_pb.get(int(lane), self.pb) if hasattr(self, "_lane_pb") else self.pb
        self._lane = int(lane)
        return self

    def mark_begin(self):
        _check(self.L.ualm_mark_begin(self.h))

    def mark_end(self):
        ms = C.c_float()
        _check(self.L.ualm_mark_end(self.h, C.byref(ms)))
        return ms.value

    def submit(self, pb, depth=2, host=None):
        """ualm_submit_batch: returns a ticket; `host` = optional pre-made (N, M, bnd, T, ixy, iyaw) host arrays (e.g. pinned)."""
        if host is None:
            host = [np.ascontiguousarray(a) for a in (pb.N.astype(np.int32), pb.M.astype(np.int32), pb.bnd.astype(np.float64),
                                                        pb.total_time.astype(np.float64), pb.inner_xy.astype(np.float64),
                                                        pb.inner_yaw.astype(np.float64))]
        N, M, bnd, T, ixy, iyaw = host
        t = C.c_int(-1)
        _check(self.L.ualm_submit_batch(self.h, pb.B, _p(N, ip), _p(M, ip), _p(bnd), _p(T), _p(ixy), _p(iyaw), int(depth), C.byref(t)))
        if not hasattr(self, "_tickets"):
            self._tickets = {}
        self._tickets[t.value] = (pb, host)
        return t.value

    def wait(self, ticket, out=None):
        """ualm_wait_batch: (results, c_xy, c_yaw) of the ticket's batch; `out` = optional preallocated (res, cxy, cyaw)."""
        pb, _ = self._tickets.pop(ticket)
        if out is None:
            out = ((Result * pb.B)(), np.zeros(int(12 * pb.N.astype(np.int64).sum())), np.zeros(int(6 * pb.M.astype(np.int64).sum())))
        res, cxy, cyaw = out
        _check(self.L.ualm_wait_batch(self.h, int(ticket), res, _p(cxy), _p(cyaw)))
        return res, cxy, cyaw

    # ---- three-step path -------------------------------------------------------------
    def upload(self, pb):
        self.pb = pb
        if not hasattr(self, "_lane_pb"):
            self._lane_pb = {}
        self._lane_pb[getattr(self, "_lane", 0)] = pb
        self._keep = [np.ascontiguousarray(a) for a in (pb.N.astype(np.int32), pb.M.astype(np.int32), pb.bnd.astype(np.float64),
                                                        pb.total_time.astype(np.float64), pb.inner_xy.astype(np.float64),
                                                        pb.inner_yaw.astype(np.float64))]
        N, M, bnd, T, ixy, iyaw = self._keep
        _check(self.L.ualm_upload(self.h, pb.B, _p(N, ip), _p(M, ip), _p(bnd), _p(T), _p(ixy), _p(iyaw)))

    def solve_resident(self):
        _check(self.L.ualm_solve_resident(self.h))

    def sync(self):
        _check(self.L.ualm_sync(self.h))

    def last_solve_ms(self):
        ms, n = C.c_float(), C.c_int()
        _check(self.L.ualm_last_solve_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def download(self):
        pb = self.pb
        res = (Result * pb.B)()
        cxy = np.zeros(int(12 * pb.N.astype(np.int64).sum()))
        cyaw = np.zeros(int(6 * pb.M.astype(np.int64).sum()))
        _check(self.L.ualm_download(self.h, res, _p(cxy), _p(cyaw)))
        return res, cxy, cyaw

    # ---- one call, host buffers in / host buffers out (the drop-in call) ---------------
    def optimize(self, pb):
        self.pb = pb
        N = np.ascontiguousarray(pb.N, np.int32); M = np.ascontiguousarray(pb.M, np.int32)
        bnd = np.ascontiguousarray(pb.bnd, np.float64); T = np.ascontiguousarray(pb.total_time, np.float64)
        ixy = np.ascontiguousarray(pb.inner_xy, np.float64); iyaw = np.ascontiguousarray(pb.inner_yaw, np.float64)
        res = (Result * pb.B)()
        cxy = np.zeros(int(12 * N.astype(np.int64).sum()))
        cyaw = np.zeros(int(6 * M.astype(np.int64).sum()))
        _check(self.L.ualm_solve_batch(self.h, pb.B, _p(N, ip), _p(M, ip), _p(bnd), _p(T), _p(ixy), _p(iyaw), res, _p(cxy), _p(cyaw)))
        return res, cxy, cyaw

    def pack_records(self, dev_ptr, stride, wait=True):
        f = self.L.ualm_pack_records_device if wait else self.L.ualm_pack_records_device_async
        _check(f(self.h, C.c_void_p(dev_ptr), stride))

    # ---- phase entry points ----------------------------------------------------------
    def eval_batch(self, x=None, lam=None, mu=None, scale_cx=None, scale_fx=None, rho=None):
        pb = self.pb
        K = self.params.int_K
        nx = int(pb.nvar().sum()); S = int(pb.nsamples(K).sum())
        f = np.zeros(pb.B); grad = np.zeros(nx); hx = np.zeros(S); gx = np.zeros(6 * S)
        cxy = np.zeros(int(12 * pb.N.astype(np.int64).sum())); cyaw = np.zeros(int(6 * pb.M.astype(np.int64).sum()))
        arrs = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (x, lam, mu, scale_cx, scale_fx)]
        _check(self.L.ualm_eval_batch(self.h, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), _p(arrs[4]),
                                      float(self.params.rho if rho is None else rho), _p(f), _p(grad), _p(hx), _p(gx), _p(cxy), _p(cyaw)))
        return dict(f=f, grad=grad, hx=hx, gx=gx, c_xy=cxy, c_yaw=cyaw)

    def init_scaling_batch(self):
        pb = self.pb
        S = int(pb.nsamples(self.params.int_K).sum())
        sfx = np.zeros(pb.B); scx = np.zeros(7 * S)
        _check(self.L.ualm_init_scaling_batch(self.h, _p(sfx), _p(scx)))
        return sfx, scx

    def time_penalty_kernel(self, reps=10):
        ms, by = C.c_float(), C.c_double()
        _check(self.L.ualm_time_penalty_kernel(self.h, reps, C.byref(ms), C.byref(by)))
        return ms.value, by.value

    def mpc_export(self, N, M, dt=0.01, init_v=None, init_a=None):
        """SE2Traj message arrays, the MPC side's re-solved coefficients and the planned-vs-tracked deviation of the resident solved
        batch (see ualm_mpc_export_batch).  N, M: the piece counts of the batch."""
        N = np.asarray(N, dtype=np.int64); M = np.asarray(M, dtype=np.int64)
        B, sN, sM = len(N), int(N.sum()), int(M.sum())
        out = dict(pos_pts=np.zeros(2 * (sN + B)), posT_pts=np.zeros(sN), angle_pts=np.zeros(sM + B), angleT_pts=np.zeros(sM),
                   c_mpc_xy=np.zeros(12 * sN), c_mpc_yaw=np.zeros(6 * sM), dev=np.zeros((B, 4)))
        dp = C.POINTER(C.c_double)
        P = lambda a: a.ctypes.data_as(dp)
        iv = None if init_v is None else np.ascontiguousarray(init_v, dtype=np.float64)
        ia = None if init_a is None else np.ascontiguousarray(init_a, dtype=np.float64)
        _check(self.L.ualm_mpc_export_batch(self.h, float(dt), None if iv is None else P(iv), None if ia is None else P(ia), P(out["pos_pts"]), P(out["posT_pts"]),
                                            P(out["angle_pts"]), P(out["angleT_pts"]), P(out["c_mpc_xy"]), P(out["c_mpc_yaw"]), P(out["dev"])))
        return out

    def feasibility(self, dt=0.01):
        """Post-solve scan of the resident batch: array [B, 10] (see ualm_feasibility_batch)."""
        out = np.zeros((self.pb.B, 10))
        _check(self.L.ualm_feasibility_batch(self.h, float(dt), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def build_map(self, pts, geom=None, ellipsoid=(0.2, 0.1, 0.1), iter_num=2, name="map"):
        """UnevenMap grid from a point cloud on the GPU (ualm_map_build_device): (UnevenMapData, kernel ms)."""
        from . import maps
        geom = geom or _lib.map_geometry()
        X, Y, W = geom.voxel_num
        cells = np.zeros((X, Y, W, 4), np.float32)
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        ms = C.c_float()
        fp = C.POINTER(C.c_float)
        _check(self.L.ualm_map_build_device(self.h, pts.ctypes.data_as(fp), pts.shape[0], C.byref(geom), ellipsoid[0], ellipsoid[1], ellipsoid[2],
                                            iter_num, cells.ctypes.data_as(fp), C.byref(ms)))
        return maps.UnevenMapData(geom, cells, name), ms.value

    PHASES = ("fill", "lu", "solve", "jerk", "tables", "samples", "accumulate", "combine", "adjoint", "tail", "twoloop", "linesearch",
              "scaling", "dual", "other", "total")

    def profile(self, enable=True, read=False):
        out = (C.c_longlong * 16)()
        _check(self.L.ualm_profile(self.h, 1 if enable else 0, out if read else None))
        return dict(zip(self.PHASES, list(out))) if read else None


def solve_batch_multi(opts, pb):
    """ualm_solve_batch_multi: one host process, one BatchALMTrajOpt per device (same params and map bound on each)."""
    L = opts[0].L
    hs = (C.c_void_p * len(opts))(*[o.h for o in opts])
    N = np.ascontiguousarray(pb.N, np.int32); M = np.ascontiguousarray(pb.M, np.int32)
    bnd = np.ascontiguousarray(pb.bnd, np.float64); T = np.ascontiguousarray(pb.total_time, np.float64)
    ixy = np.ascontiguousarray(pb.inner_xy, np.float64); iyaw = np.ascontiguousarray(pb.inner_yaw, np.float64)
    res = (Result * pb.B)()
    cxy = np.zeros(int(12 * N.astype(np.int64).sum())); cyaw = np.zeros(int(6 * M.astype(np.int64).sum()))
    _check(L.ualm_solve_batch_multi(hs, len(opts), pb.B, _p(N, ip), _p(M, ip), _p(bnd), _p(T), _p(ixy), _p(iyaw), res, _p(cxy), _p(cyaw)))
    return res, cxy, cyaw
