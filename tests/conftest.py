import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    # a GPU test that ends up on a machine without a GPU must fail loudly, never skip
    # silently into a CPU path; but plain `pytest tests/` on the CPU box deselects them.
    if config.getoption("-m"):
        return
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        skip = pytest.mark.skip(reason="no CUDA device (GPU tests run with -m gpu on a B200)")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)
