import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def cornell():
    from pgi_raytracing_b200 import scenes
    sc = scenes.cornell_like()
    sc.camera = scenes.Camera(64, 48, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)
    return sc


def golden(name):
    return np.load(os.path.join(GOLDEN, name))
