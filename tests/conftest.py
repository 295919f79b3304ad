"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path.

`-m "not gpu"` runs on the CPU-only build container (oracle vs golden vectors, host logic,
C-ABI symbol checks, gloo world_size-2 tests); `-m gpu` runs on a B200 and drives the CUDA
path through the C ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def scipy_vectors(golden_dir):
    import numpy as np
    return np.load(os.path.join(golden_dir, "scipy_vectors.npz"))
