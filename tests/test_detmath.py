"""include/ualm_detmath.h (deterministic sin/cos/atan2 shared by the CUDA path and the oracle) pinned against libm."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "ualm_detmath.h"
extern "C" void t_sincos(const double*x,int n,double*s,double*c){for(int i=0;i<n;i++)ualm_sincos(x[i],s+i,c+i);}
extern "C" void t_atan2(const double*y,const double*x,int n,double*o){for(int i=0;i<n;i++)o[i]=ualm_atan2(y[i],x[i]);}
'''


@pytest.fixture(scope="module")
def dm(tmp_path_factory):
    d = tmp_path_factory.mktemp("dm")
    (d / "t.cpp").write_text(SRC)
    subprocess.run(["g++", "-O3", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", str(d / "t.so"), str(d / "t.cpp")], check=True)
    return C.CDLL(str(d / "t.so"))


dp = C.POINTER(C.c_double)


def ulp(a, b):
    return np.abs(a - b) / np.spacing(np.abs(b) + 1e-300)


def test_sincos_within_2ulp(dm):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-50, 50, 400000), rng.uniform(-1e5, 1e5, 50000), rng.uniform(-1e-3, 1e-3, 5000), [0.0, np.pi, -np.pi, np.pi / 2]])
    s = np.zeros_like(x); c = np.zeros_like(x)
    dm.t_sincos(x.ctypes.data_as(dp), len(x), s.ctypes.data_as(dp), c.ctypes.data_as(dp))
    assert ulp(s, np.sin(x)).max() <= 2.0
    assert ulp(c, np.cos(x)).max() <= 2.0
    assert np.abs(s * s + c * c - 1).max() < 1e-15


def test_atan2_within_2ulp_and_special(dm):
    rng = np.random.default_rng(1)
    y = rng.standard_normal(400000) * 10 ** rng.uniform(-3, 3, 400000)
    x = rng.standard_normal(400000) * 10 ** rng.uniform(-3, 3, 400000)
    o = np.zeros_like(y)
    dm.t_atan2(y.ctypes.data_as(dp), x.ctypes.data_as(dp), len(y), o.ctypes.data_as(dp))
    assert ulp(o, np.arctan2(y, x)).max() <= 2.0
    ys = np.array([0.0, 0.0, 1.0, -1.0, 0.0]); xs = np.array([1.0, -1.0, 0.0, 0.0, 0.0]); o = np.zeros(5)
    dm.t_atan2(ys.ctypes.data_as(dp), xs.ctypes.data_as(dp), 5, o.ctypes.data_as(dp))
    assert np.allclose(o, [0.0, np.pi, np.pi / 2, -np.pi / 2, 0.0])


def test_so2_wrap_roundtrip(dm):
    """the only use of atan2 on the hot path: diff = atan2(sin d, cos d) (uneven_map.h:284)"""
    rng = np.random.default_rng(2)
    d = rng.uniform(-3.0, 3.0, 200000)
    s = np.zeros_like(d); c = np.zeros_like(d); o = np.zeros_like(d)
    dm.t_sincos(d.ctypes.data_as(dp), len(d), s.ctypes.data_as(dp), c.ctypes.data_as(dp))
    dm.t_atan2(s.ctypes.data_as(dp), c.ctypes.data_as(dp), len(d), o.ctypes.data_as(dp))
    assert np.abs(o - d).max() < 1e-15
