import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device: a plain `pytest` run skips the gpu-marked tests (they cannot run here); a run that ASKS for
    them (`-m gpu`) fails loudly instead of passing silently."""
    import torch
    if torch.cuda.is_available():
        return
    markexpr = config.getoption("-m") or ""
    asked = "gpu" in markexpr and "not gpu" not in markexpr
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if asked and gpu_items:
        raise pytest.UsageError("`-m gpu` needs a CUDA device: none is visible (the library has no CPU path)")
    for it in gpu_items:
        it.add_marker(pytest.mark.skip(reason="no CUDA device (run with -m gpu on the B200 box)"))
