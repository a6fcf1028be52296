import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    """The product package; builds libpsagpu.so first if the tree has not been built yet (nvcc cross-compiles
    sm_100a without a GPU)."""
    import subprocess
    import psa_b200_loader
    lib = os.path.join(psa_b200_loader.PKG_DIR, "libpsagpu.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-s", "-C", ROOT, "all"])
    return psa_b200_loader.load_package()


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False)
    return load
