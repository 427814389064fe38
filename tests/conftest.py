import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def pytest_collection_modifyitems(config, items):
    """Two-rank GPU tests (tests/test_gpu_multi.py) need two devices: on a one-GPU box they are deselected, not skipped
    (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` runs them; result kept under profiles/)."""
    multi = [it for it in items if "test_gpu_multi" in it.nodeid]
    if not multi:
        return
    import torch
    if torch.cuda.is_available() and torch.cuda.device_count() >= 2:
        return
    config.hook.pytest_deselected(items=multi)
    items[:] = [it for it in items if "test_gpu_multi" not in it.nodeid]
