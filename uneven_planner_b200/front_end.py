"""ctypes plumbing for the front-end entry points of libualm.so: KinoAstar::plan for one (start, goal) pair and the
batch form (search + PlanManager's resampler) that produces the optimizer's inputs."""
import ctypes as C

import numpy as np

from . import _lib
from .problems import ProblemBatch

dp = C.POINTER(C.c_double)


def default_params():
    p = _lib.AstarParams()
    _lib.lib().ualm_astar_default_params(C.byref(p))
    return p


def resample_params(piece_len=0.3, yaw_piece_times=2.0, mean_vel=0.5, init_time_times=1.2, init_sig_vel=0.05):
    """plan_manager/* of run_hill.yaml:57-62"""
    return _lib.ResampleParams(piece_len, yaw_piece_times, mean_vel, init_time_times, init_sig_vel)


class MapView:
    """the ualm_astar_map_t of a UnevenMapData (keeps the arrays alive)"""

    def __init__(self, mapdata, min_cnormal=0.8, max_rho=0.05, use_cells64=True):
        self.mapdata = mapdata
        self.occ3, self.occ2 = mapdata.occupancy(min_cnormal, max_rho)
        self.geom = mapdata.geom
        u8 = C.POINTER(C.c_uint8)
        c64 = mapdata.cells64 if (use_cells64 and mapdata.cells64 is not None) else None
        self.c = _lib.AstarMap(C.pointer(self.geom), mapdata.cells.ctypes.data_as(C.POINTER(C.c_float)),
                               c64.ctypes.data_as(dp) if c64 is not None else None, self.occ3.ctypes.data_as(u8), self.occ2.ctypes.data_as(u8))


def plan(view, start, goal, params=None, max_pts=1 << 16):
    """KinoAstar::plan: (n, 3) polyline, empty when there is no path; also the number of closed nodes"""
    params = params or default_params()
    s = np.ascontiguousarray(start, dtype=np.float64); g = np.ascontiguousarray(goal, dtype=np.float64)
    buf = np.zeros((max_pts, 3))
    ex = C.c_int(0)
    n = _lib.lib().ualm_kino_astar_plan(C.byref(view.c), C.byref(params), s.ctypes.data_as(dp), g.ctypes.data_as(dp), buf.ctypes.data_as(dp), max_pts, C.byref(ex))
    if n < 0:
        raise RuntimeError(f"ualm_kino_astar_plan: {n}")
    return buf[:n].copy(), ex.value


def plan_batch(view, starts, goals, params=None, rparams=None, nthreads=0):
    """(start, goal) pairs -> ProblemBatch of the pairs that have a path, the index of every pair in it (or -1), closed-node counts"""
    params = params or default_params()
    rparams = rparams or resample_params()
    starts = np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, 3); goals = np.ascontiguousarray(goals, dtype=np.float64).reshape(-1, 3)
    B = starts.shape[0]
    ip = C.POINTER(C.c_int32)
    N = np.zeros(B, np.int32); M = np.zeros(B, np.int32); bnd = np.zeros((B, 18)); T = np.zeros(B)
    ixy = np.zeros(2 * 63 * max(B, 1)); iyaw = np.zeros(127 * max(B, 1))
    packed = np.zeros(B, np.int32); nexp = np.zeros(B, np.int32)
    k = _lib.lib().ualm_front_end_batch(C.byref(view.c), C.byref(params), C.byref(rparams), B, starts.ctypes.data_as(dp), goals.ctypes.data_as(dp), nthreads,
                                        N.ctypes.data_as(ip), M.ctypes.data_as(ip), bnd.ctypes.data_as(dp), T.ctypes.data_as(dp), ixy.ctypes.data_as(dp), ixy.size,
                                        iyaw.ctypes.data_as(dp), iyaw.size, packed.ctypes.data_as(ip), nexp.ctypes.data_as(ip))
    if k < 0:
        raise RuntimeError(f"ualm_front_end_batch: {k}")
    nxy = int((2 * (N[:k].astype(np.int64) - 1)).sum()); nyaw = int((M[:k].astype(np.int64) - 1).sum())
    pb = ProblemBatch(N[:k].copy(), M[:k].copy(), bnd[:k].copy(), T[:k].copy(), ixy[:nxy].copy(), iyaw[:nyaw].copy())
    keep = packed >= 0
    pb.starts = starts[keep]; pb.goals = goals[keep]
    return pb, packed, nexp
