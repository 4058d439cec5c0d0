import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_available() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_lib

    return oracle_lib.load()


@pytest.fixture(scope="session")
def ref_shim():
    """The reference's own CUDA kernels (oracle/_ref/libref_shim.so); GPU tests only."""
    from tests import oracle_lib

    lib = oracle_lib.load_ref_shim()
    if lib is None:
        pytest.skip("oracle/_ref/libref_shim.so not built")
    return lib
