"""pytest wiring: the `gpu` marker, repo root on sys.path, and the in-tree library.

`-m "not gpu"` runs here (no GPU, no libcuda): oracle vs golden vectors, host logic, the
C-ABI surface, allocator bookkeeping against the mock driver.  `-m gpu` runs on the B200
box and calls the CUDA path through the C ABI.
"""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # build once if the shared library did not travel with the tree
    from vattention_b200.build import LIB_PATH, build_library
    if not LIB_PATH.exists():
        build_library()


@pytest.fixture(scope="session")
def repo_root() -> Path:
    return ROOT


def has_cuda() -> bool:
    import torch
    return torch.cuda.is_available()


def pytest_collection_modifyitems(config, items):
    if has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
