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


def _have_b200():
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need an sm_100 device: on any other box they are skipped (not failed), so a plain `pytest`
    shows the CPU-runnable oracle / golden / host-logic tests green."""
    if _have_b200():
        return
    skip = pytest.mark.skip(reason="needs a B200 (sm_100) CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
