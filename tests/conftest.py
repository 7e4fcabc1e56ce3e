import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def pytest_collection_modifyitems(config, items):
    """GPU-marked tests skip (not fail) on a machine without a CUDA device."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library, (re)built if stale -- nvcc cross-compiles without a GPU."""
    from bigcn_b200.csrc.build import build
    return build()


@pytest.fixture(scope="session")
def dev(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
