import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "3d-brain-tumor-segmentation_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def b3d():
    lib = os.path.join(ROOT, PKG, "csrc", "libb3d.so")
    if not os.path.isfile(lib):
        import __graft_entry__ as g
        g.build()
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def dev():
    import torch
    if not torch.cuda.is_available():      # `-m gpu` on a CPU-only runner: skip, do not error
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")
