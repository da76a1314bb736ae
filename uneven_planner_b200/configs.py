"""The ONE table of per-terrain parameters: deltas of plan_manager/params/run_<terrain>.yaml against run_hill.yaml, plus the
explicit overrides BASELINE.json's configs add.  `maps.py` (occupancy threshold of the problem generator) and every test /
bench workload read it from here; `tests/test_host_tools.py::test_config_table_matches_reference_yaml` re-derives it from the
reference's yaml files where they exist.

yaml facts (diff of /root/reference/src/uneven_planner/plan_manager/params/run_*.yaml, comments stripped):
  run_desert.yaml, run_mountain.yaml, run_all.yaml == run_hill.yaml
  run_vocano.yaml : uneven_map/max_rho 0.08, alm_traj_opt/max_sig 0.08
  run_forest.yaml : uneven_map/max_rho 0.001, alm_traj_opt/{rho_T 500, max_sig 0.001, use_scaling false}
BASELINE configs: 1-2 hill, 3 desert, 4 volcano + {max_kap 0.3, int_K 64} (penalty-dense stress), 5 forest (no ESDF term
exists in the reference).
"""
from . import _lib

YAML = {"hill": "run_hill.yaml", "desert": "run_desert.yaml", "volcano": "run_vocano.yaml", "forest": "run_forest.yaml",
        "mountain": "run_mountain.yaml"}

# params: alm_traj_opt/* deltas; max_rho / min_cnormal: uneven_map/* occupancy thresholds (uneven_map.cpp:169-179)
TERRAINS = {
    "hill": dict(pcd="hill.pcd", params={}, max_rho=0.05, min_cnormal=0.8),
    "desert": dict(pcd="desert.pcd", params={}, max_rho=0.05, min_cnormal=0.8),
    "volcano": dict(pcd="vocano.pcd", params=dict(max_sig=0.08), max_rho=0.08, min_cnormal=0.8),
    "forest": dict(pcd="forest.pcd", params=dict(use_scaling=0, rho_T=500.0, max_sig=0.001), max_rho=0.001, min_cnormal=0.8),
    "mountain": dict(pcd="mountain.pcd", params={}, max_rho=0.05, min_cnormal=0.8),
}

# overrides a BASELINE.json config adds on top of the terrain's yaml (kept explicit, never folded into TERRAINS)
CONFIG_OVERRIDES = {
    "volcano": dict(max_kap=0.3, int_K=64),      # configs[3]: "tight kappa_max=0.3, 64 constraint samples/segment"
}

# BASELINE.json configs -> (terrain, batch per step, scaling)
BASELINE_CONFIGS = {
    1: dict(terrain="hill", batch=1, scaling="weak"),
    2: dict(terrain="hill", batch=1024, scaling="weak"),
    3: dict(terrain="desert", batch=8192, scaling="strong"),
    4: dict(terrain="volcano", batch=1024, scaling="weak"),
    5: dict(terrain="forest", batch=4096, scaling="weak"),
}


def params_for(name, overrides=True):
    """ualm_params_t of a terrain: run_hill.yaml defaults + the terrain's yaml deltas (+ the BASELINE config's overrides)."""
    p = _lib.default_params()
    for k, v in TERRAINS[name]["params"].items():
        setattr(p, k, v)
    if overrides:
        for k, v in CONFIG_OVERRIDES.get(name, {}).items():
            setattr(p, k, v)
    return p


def gen_kwargs(name):
    """Occupancy thresholds of the problem generator (KinoAstar's entry checks use the map's own occupancy grid)."""
    t = TERRAINS[name]
    return dict(max_rho=t["max_rho"], min_cnormal=t["min_cnormal"])
