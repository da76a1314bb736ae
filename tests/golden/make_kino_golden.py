"""Generates tests/golden/kino_astar_golden.npz: polylines of the REFERENCE's KinoAstar::plan (front_end/src/kino_astar.cpp compiled unmodified
into oracle/_ref/librefkino.so, see oracle/ref_kino_driver.cpp) for seeded (start, goal) pairs on the synthetic terrain every box can rebuild
(maps.synthetic_terrain("bumps", seed=3), occupancy thresholds min_cnormal 0.8 / max_rho 0.003) with the yaml's kino_astar parameters.  Run in the
build container (needs /root/reference).  The product's ualm_kino_astar_plan must reproduce them bit for bit wherever the tests run."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from uneven_planner_b200 import _lib, front_end, maps

dp = C.POINTER(C.c_double); u8 = C.POINTER(C.c_uint8)
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "librefkino.so"))
ref.ref_kino_plan.argtypes = [dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp, dp, dp, C.c_int, u8, u8]
m = maps.synthetic_terrain("bumps", seed=3)
g = m.geom
cells64 = np.ascontiguousarray(m.cells, dtype=np.float64)
ap = front_end.default_params()
kp = np.array([getattr(ap, n) for n, _ in _lib.AstarParams._fields_])
rng = np.random.default_rng(77)
B = 16
starts = np.column_stack([rng.uniform(-4.3, 4.3, B), rng.uniform(-4.3, 4.3, B), rng.uniform(-np.pi, np.pi, B)])
goals = np.column_stack([rng.uniform(-4.3, 4.3, B), rng.uniform(-4.3, 4.3, B), rng.uniform(-np.pi, np.pi, B)])
paths, lens = [], []
for b in range(B):
    out = np.zeros((1 << 15, 3))
    s = np.ascontiguousarray(starts[b]); e = np.ascontiguousarray(goals[b])
    n = ref.ref_kino_plan(cells64.ctypes.data_as(dp), -2 * g.origin[0], -2 * g.origin[1], g.xy_resolution, g.yaw_resolution, 0.8, 0.003, kp.ctypes.data_as(dp),
                          s.ctypes.data_as(dp), e.ctypes.data_as(dp), out.ctypes.data_as(dp), out.shape[0], None, None)
    assert n >= 0
    paths.append(out[:n].copy()); lens.append(n)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "kino_astar_golden.npz"), starts=starts, goals=goals, lens=np.array(lens),
                    paths=np.concatenate(paths) if sum(lens) else np.zeros((0, 3)), min_cnormal=0.8, max_rho=0.003)
print("wrote", B, "reference front-end paths:", lens)
