import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the `gpu` tests are skipped, not failed (the driver's CPU run deselects them with -m "not gpu";
    this covers a plain `pytest tests/` on a machine without a GPU)."""
    try:
        import torch

        has_cuda = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run on the B200 box: pytest -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import mg_oracle

    mg_oracle.build()
    return mg_oracle


@pytest.fixture(scope="session")
def cuda_lib():
    from marlgrid_b200 import _lib

    return _lib.load()
