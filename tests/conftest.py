import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def cube_mesh():
    m = json.load(open(os.path.join(ROOT, "tests", "golden", "cube_mesh.json")))
    return np.array(m["vertices"], np.float32), np.array(m["faces"], np.int32)


@pytest.fixture(scope="session")
def cube_pair(oracle, cube_mesh):
    """The fixture of reference test/test_gicp_alignment.cpp:32-47: cube.ply sampled with 5000 points (libc rand(),
    never seeded) = source; target = source yawed by 0.175 rad through Utils::rotateCloud."""
    V, F = cube_mesh
    src = oracle.sample_mesh(V, F, 5000)
    T = oracle.rotation_rpy(0.0, 0.0, 0.175)
    tgt = oracle.transform(T, src)
    return src, tgt, T


@pytest.fixture(scope="session")
def engine():
    from leica_point_cloud_processing_b200 import Engine
    e = Engine(0)
    yield e
    e.close()
