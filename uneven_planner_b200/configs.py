"""The parameter sets of BASELINE.json's configs, as deltas to run_hill.yaml.

hill     configs 1-2: plan_manager/params/run_hill.yaml
desert   config 3:    run_desert.yaml (max_sig = uneven_map max_rho = 0.08)
volcano  config 4:    run_vocano.yaml (use_scaling=false, rho_T=500, max_sig=0.001) + the config's overrides max_kap=0.3, int_K=64
forest   config 5:    run_forest.yaml == hill parameters (there is no ESDF term in the reference)
"""
from . import _lib

CONFIGS = {
    "hill": dict(params={}, gen=dict(max_rho=0.05)),
    "desert": dict(params=dict(max_sig=0.08), gen=dict(max_rho=0.08)),
    "volcano": dict(params=dict(use_scaling=0, rho_T=500.0, max_sig=0.001, max_kap=0.3, int_K=64), gen=dict(max_rho=0.05)),
    "forest": dict(params={}, gen=dict(max_rho=0.05)),
}


def params_for(name):
    p = _lib.default_params()
    for k, v in CONFIGS[name]["params"].items():
        setattr(p, k, v)
    return p


def gen_kwargs(name):
    return dict(CONFIGS[name]["gen"])
