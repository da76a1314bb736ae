"""The C++ host-side mirror of the reference's ALMTrajOpt interface (include/ualm_traj_opt.hpp)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(built, tmp_path_factory):
    d = tmp_path_factory.mktemp("cpp")
    exe = str(d / "adapter_driver")
    lib = os.path.join(ROOT, "uneven_planner_b200")
    subprocess.run(["g++", "-O2", "-std=c++14", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "adapter_driver.cpp"),
                    "-o", exe, "-L", lib, "-lualm", "-Wl,-rpath," + lib], check=True)
    return exe


def _write_case(path, m, pb, i):
    oxy, oyaw, _, _ = pb.offsets()
    with open(path, "wb") as f:
        f.write(struct.pack("<5i", *m.shape, int(pb.N[i]), int(pb.M[i])))
        f.write(m.cells.tobytes())
        f.write(pb.bnd[i].astype(np.float64).tobytes())
        f.write(struct.pack("<d", float(pb.total_time[i])))
        f.write(pb.inner_xy[oxy[i]:oxy[i + 1]].astype(np.float64).tobytes())
        f.write(pb.inner_yaw[oyaw[i]:oyaw[i + 1]].astype(np.float64).tobytes())


def test_result_containers_and_mpc_message_export(built, tmp_path):
    """make_traj / Piece::getValue / locatePieceIdx / toSE2TrajMsg of include/ualm_traj_opt.hpp against hand-computed values
    (se2traj.hpp:106-118, 343-367, 682-695; plan_manager.cpp:151-185) -- host only."""
    exe = str(tmp_path / "msg_driver")
    lib = os.path.join(ROOT, "uneven_planner_b200")
    subprocess.run(["g++", "-O2", "-std=c++14", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "msg_driver.cpp"),
                    "-o", exe, "-L", lib, "-lualm", "-Wl,-rpath," + lib], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "OK", r.stdout


def test_adapter_compiles_and_fails_loudly_without_gpu(driver, bumps_map, tmp_path):
    import torch
    from uneven_planner_b200 import problems
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pb = problems.generate(bumps_map, 1, seed=1)
    case = str(tmp_path / "case.bin")
    _write_case(case, bumps_map, pb, 0)
    r = subprocess.run([driver, case], capture_output=True, text=True)
    assert r.returncode == 10 and "EXCEPTION" in r.stdout and "CUDA" in r.stdout


@pytest.mark.gpu
def test_adapter_matches_batch_api(driver, bumps_map, tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from uneven_planner_b200 import _lib, api, problems
    pb = problems.generate(bumps_map, 2, seed=1)
    case = str(tmp_path / "case.bin")
    _write_case(case, bumps_map, pb, 1)
    r = subprocess.run([driver, case], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    head = lines[0].split()
    opt = api.BatchALMTrajOpt().init(_lib.default_params()).set_environment(bumps_map)
    res, cxy, cyaw = opt.optimize(pb)
    _, _, ocx, _ = pb.offsets()
    assert int(head[1]) == res[1].ret_code and int(head[3]) == res[1].n_evals and float(head[5]) == res[1].inner_cost
    N = int(pb.N[1])
    c = cxy[ocx[1]:ocx[2]]
    # getTraj() carries the piece durations of the last evaluation (T1(i), se2traj.hpp:682-695): bit-identical to the oracle's and
    # (tests/test_ref_pin.py) to the reference build's
    import pyoracle as po
    r_o = po.solve_one(po.params_from(_lib.default_params()), po.OracleMap(bumps_map), pb, 1)[0]
    dur = [float(x) for x in lines[2].split()[1:]]
    assert dur[0] == res[1].piece_T_xy and dur[1] == res[1].piece_T_yaw
    tt = 0.0
    for _ in range(int(pb.N[1])):
        tt += dur[0]
    assert tt == res[1].total_T == r_o.total_T
    lines = lines[:2] + lines[3:]
    coeffs = np.array([float(x) for x in lines[2:2 + 12 * N]]).reshape(N, 2, 6)
    for i in range(N):
        for d in range(2):
            assert np.array_equal(coeffs[i, d][::-1], c[6 * i + d * 6 * N: 6 * i + 6 + d * 6 * N])   # highest power first
    start = [float(x) for x in lines[1].split()[1:]]
    assert np.allclose(start, pb.bnd[1][:2], atol=1e-12)
