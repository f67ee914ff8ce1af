import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        from lidar_slam_b200 import capi
        return capi.lib().b2_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; when someone runs the whole suite on a CPU box they are skipped
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build(ref=True)
    return O


@pytest.fixture(scope="session")
def scene():
    from lidar_slam_b200 import synth
    return synth.Scene(leg=100.0)


@pytest.fixture(scope="session")
def small_map(scene):
    return scene.make_map(300_000, 2.0)


@pytest.fixture(scope="session")
def scans(scene):
    """a few raw HDL-64 scans with their true poses"""
    out = []
    for k, s in enumerate((41.3, 63.5, 88.0, 127.0)):
        p = scene.path_pose(s)
        out.append((p, scene.scan(100 + k, p)))
    return out


def f32(x):
    return float(np.float32(x))
