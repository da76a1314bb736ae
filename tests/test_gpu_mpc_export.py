"""SURVEY 8f-3 on the GPU: the SE2Traj message, the MPC side's MINCO re-solve and the planned-vs-tracked deviation of a solved batch
(ualm_mpc_export_batch, csrc/ualm_kernels.cuh: mpc_export_kernel).

Parity chain: the MPC's own header (mpc_controller/include/utils/minco_traj.hpp:365-444, compiled unmodified) equals orc_minco_generate bit
for bit on such messages (tests/test_ref_pin.py::test_mpc_side_minco_resolve_matches_oracle_bitwise, CPU); here the CUDA kernel equals
orc_minco_generate bit for bit on the messages of solved batches, so the trajectory the MPC tracks behind this back-end is the one it tracks
behind the reference's.  The message itself is compared with the values the reference's publish loop reads off getTraj()
(plan_manager.cpp:159-184): constant coefficients and PolyTrajectory::getValue(total)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
dp = C.POINTER(C.c_double)
P = lambda a: a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def gpu(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from uneven_planner_b200 import api
    return api


@pytest.fixture(scope="module")
def orc(built):
    import pyoracle as po
    L = po.lib()
    L.orc_minco_generate.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp]
    return L


def _end_value(c, Pn, dur):
    """PolyTrajectory::getValue(getTotalDuration()) (se2traj.hpp:291-300, 343-368) for uniform durations"""
    total = 0.0
    for _ in range(Pn):
        total += dur
    t, idx = total, 0
    while idx < Pn and t > dur:
        t -= dur; idx += 1
    if idx == Pn:
        idx -= 1; t += dur
    v, tn = 0.0, 1.0
    for k in range(6):
        v += tn * c[6 * idx + k]; tn *= t
    return v


def _eval(c, Pn, dur, t):
    idx = 0
    while idx < Pn and t > dur:
        t -= dur; idx += 1
    if idx == Pn:
        idx -= 1; t += dur
    v, tn = 0.0, 1.0
    for k in range(6):
        v += tn * c[6 * idx + k]; tn *= t
    return v


@pytest.mark.parametrize("prec", [64, 32])
def test_message_and_mpc_resolve_of_a_solved_batch(gpu, orc, bumps_map, prec):
    from uneven_planner_b200 import _lib, problems
    params = _lib.default_params()
    pb = problems.generate(bumps_map, 24, seed=12)
    opt = gpu.BatchALMTrajOpt(precision=prec).init(params).set_environment(bumps_map)
    res, cxy, cyaw = opt.optimize(pb)
    ex = opt.mpc_export(pb.N, pb.M, dt=0.01)
    opt.close()
    _, _, ocx, ocy = pb.offsets()
    sN = np.concatenate([[0], np.cumsum(pb.N)]); sM = np.concatenate([[0], np.cumsum(pb.M)])
    worst = 0.0
    for i in range(pb.B):
        N, M = int(pb.N[i]), int(pb.M[i])
        c2 = cxy[ocx[i]:ocx[i + 1]]; cy = cyaw[ocy[i]:ocy[i + 1]]
        Tx, Ty = res[i].piece_T_xy, res[i].piece_T_yaw
        pp = ex["pos_pts"][2 * (sN[i] + i):2 * (sN[i + 1] + i + 1)].reshape(N + 1, 2)
        ap = ex["angle_pts"][sM[i] + i:sM[i + 1] + i + 1]
        # (1) the message, bit for bit what the publish loop reads
        assert np.array_equal(pp[:N, 0], c2[0:6 * N:6]) and np.array_equal(pp[:N, 1], c2[6 * N::6])
        assert pp[N, 0] == _end_value(c2[:6 * N], N, Tx) and pp[N, 1] == _end_value(c2[6 * N:], N, Tx)
        assert np.array_equal(ap[:M], cy[0::6]) and ap[M] == _end_value(cy, M, Ty)
        assert np.all(ex["posT_pts"][sN[i]:sN[i + 1]] == Tx) and np.all(ex["angleT_pts"][sM[i]:sM[i + 1]] == Ty)
        # (2) the MPC side's MINCO over that message == the oracle's (== the reference header's), bit for bit
        for Dim, Pn, pts, dur, got in ((2, N, pp, Tx, ex["c_mpc_xy"][ocx[i]:ocx[i + 1]]), (1, M, ap.reshape(M + 1, 1), Ty, ex["c_mpc_yaw"][ocy[i]:ocy[i + 1]])):
            head = np.zeros((3, Dim)); tail = np.zeros((3, Dim))
            head[0] = pts[0]; tail[0] = pts[Pn]
            inPs = np.ascontiguousarray(pts[1:Pn]).ravel()        # Dim x (P - 1) column-major == (P - 1) x Dim row-major
            want = np.zeros(6 * Pn * Dim)
            orc.orc_minco_generate(Dim, Pn, P(inPs if inPs.size else np.zeros(1)), P(np.full(Pn, dur)), P(np.ascontiguousarray(head).ravel()),
                                   P(np.ascontiguousarray(tail).ravel()), P(want))
            assert np.array_equal(got, want), (i, Dim)
        # (3) planned against tracked: recompute the scan on the host
        mx = ex["c_mpc_xy"][ocx[i]:ocx[i + 1]]; my = ex["c_mpc_yaw"][ocy[i]:ocy[i + 1]]
        tot = min(sum([Tx] * N), sum([Ty] * M))
        t, bp, by = 0.0, 0.0, 0.0
        while t < tot:
            ep = np.hypot(_eval(c2[:6 * N], N, Tx, t) - _eval(mx[:6 * N], N, Tx, t), _eval(c2[6 * N:], N, Tx, t) - _eval(mx[6 * N:], N, Tx, t))
            ey = abs(_eval(cy, M, Ty, t) - _eval(my, M, Ty, t))
            bp, by = max(bp, ep), max(by, ey)
            t += 0.01
        assert abs(ex["dev"][i, 0] - bp) <= 1e-12 and abs(ex["dev"][i, 2] - by) <= 1e-12
        worst = max(worst, bp)
        # the tracked spline interpolates the same waypoints: the difference comes from the boundary derivatives alone and stays small
        assert np.allclose(mx[0:6 * N:6], c2[0:6 * N:6], atol=1e-9)
    assert worst < 0.05     # metres; measured ~ mm: 0.05 m/s of boundary speed (plan_manager.cpp:93-94) against none


def test_export_needs_a_solved_batch_and_skips_oversize_problems(gpu, bumps_map):
    from uneven_planner_b200 import _lib, problems
    from test_gpu_boundary import _with_oversize
    params = _lib.default_params()
    pb = problems.generate(bumps_map, 6, seed=5)
    opt = gpu.BatchALMTrajOpt().init(params).set_environment(bumps_map)
    with pytest.raises(gpu.UalmError):
        opt.mpc_export(pb.N, pb.M)
    pbx = _with_oversize(pb, 2)
    res, cxy, cyaw = opt.optimize(pbx)
    ex = opt.mpc_export(pbx.N, pbx.M)
    opt.close()
    assert res[2].ret_code == _lib.UALM_ELIMIT
    sN = np.concatenate([[0], np.cumsum(pbx.N)])
    assert not ex["pos_pts"][2 * (sN[2] + 2):2 * (sN[3] + 3)].any() and not ex["dev"][2].any()
    assert ex["pos_pts"][2 * (sN[3] + 3)] == cxy[pbx.offsets()[2][3]]
