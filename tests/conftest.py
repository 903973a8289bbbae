import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA B200 device (run on the GPU box via gpurun)")
    config.addinivalue_line("markers", "reference: needs the read-only reference checkout (build container only)")


@pytest.fixture(scope="session")
def built_lib():
    import __graft_entry__ as g
    g.build()
    from windgym_b200 import _lib
    return _lib.load()


def pytest_collection_modifyitems(config, items):
    """CPU-only machines: skip (not fail) everything marked ``gpu`` when no CUDA device is visible."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA B200 device (run through gpurun)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
