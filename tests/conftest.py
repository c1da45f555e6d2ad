import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context for the whole GPU session (fails loudly if the library is missing)."""
    from atm_raytracer_b200 import runtime

    c = runtime.Context(0)
    yield c
    c.close()


def scene(name, scale):
    """(params, Terrain, objects, textures) of a BASELINE workload at a reduced image size."""
    from atm_raytracer_b200 import config, runtime, scenes

    cfg, grid = scenes.make_scene(name, scale=scale)
    terrain = runtime.Terrain.from_arrays(scenes.terrain_arrays(grid))
    params = config.into_params(cfg)
    objects, textures = config.lower_objects(cfg)
    return params, terrain, objects, textures


def ramp_tile(lat0, lon0, n=121, a=3, b=-2, c=500):
    """Planar-ramp tile posts[ilon][ilat] = a*ilon + b*ilat + c (bilinear sampling is exact on it)."""
    ilon, ilat = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    return (a * ilon + b * ilat + c).astype(np.int16)
