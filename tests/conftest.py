import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Make sure libctc_b200.so exists and is current before any test touches the engine (a no-op when it is; nvcc
    cross-compiles for sm_100a without a GPU).  The engine itself never builds or falls back: it fails loudly."""
    from aes_lac_2018_b200 import build
    try:
        build.build()
    except RuntimeError:
        if not os.path.exists(build.LIB):
            raise
    yield
