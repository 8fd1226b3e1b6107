"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path.

`python -m pytest tests -m "not gpu"` runs on any CPU box (oracle vs golden vectors, host logic, C-ABI
export checks, gloo world_size-2 sharding).  `-m gpu` needs a B200 and the built CUDA library and calls
the kernels through the C-ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CKPT = os.path.join(GOLDEN, "pointnet2-inview-0.55884-0001.pth")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200) and the built libpn12_b200.so")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        return cache[name]

    return load


@pytest.fixture(scope="session")
def ckpt_path():
    return CKPT


@pytest.fixture(scope="session")
def ckpt_state():
    """The reference's shipped PointNet2SemSeg checkpoint as {name: ndarray} without the module. prefix."""
    import torch

    from oracle import oracle as orc

    return orc.numpy_state_dict(torch.load(CKPT, map_location="cpu"))
