import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) liboracle.so and libualm.so."""
    import pyoracle
    pyoracle.build()
    from uneven_planner_b200 import build as b
    b.build()
    return True


@pytest.fixture(scope="session")
def bumps_map(built):
    from uneven_planner_b200 import maps
    return maps.synthetic_terrain("bumps", seed=3)


@pytest.fixture(scope="session")
def hill_map(built):
    from uneven_planner_b200 import maps
    m = maps.get_terrain("hill")
    if m is None:
        pytest.skip("maps_built/hill.umap not present (built by __graft_entry__.build() where the reference clouds exist)")
    return m


@pytest.fixture(scope="session")
def terrain(built):
    """terrain(name) -> UnevenMapData of one of the reference's terrains (skips when its .umap did not travel)."""
    from uneven_planner_b200 import maps

    def get(name):
        m = maps.get_terrain(name)
        if m is None:
            pytest.skip(f"maps_built/{name}.umap not present (built by __graft_entry__.build() where the reference clouds exist)")
        return m
    return get


@pytest.fixture(scope="session")
def oparams(built):
    import pyoracle as po
    from uneven_planner_b200 import _lib
    return po.params_from(_lib.default_params())
