import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "contexture-nerf_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_run_nerf_helpers.npz"))


@pytest.fixture(scope="session")
def libpath():
    """Path of libctxnerf.so, building it (nvcc cross-compiles without a GPU) if absent."""
    from ctxnerf import build
    return build.build()


@pytest.fixture(scope="session")
def cuda(libpath):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)
