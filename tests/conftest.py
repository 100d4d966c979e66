import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def mazes():
    z = np.load(os.path.join(REPO, "ditreeonlineplanner_b200", "data", "mazes.npz"))
    return {k: z[k].astype(np.float32) for k in z.files}


@pytest.fixture(scope="session")
def car_meta():
    z = np.load(os.path.join(REPO, "ditreeonlineplanner_b200", "data", "metadata_carmaze.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def ant_meta():
    z = np.load(os.path.join(REPO, "ditreeonlineplanner_b200", "data", "metadata_antmaze.npz"))
    return {k: z[k] for k in z.files}
