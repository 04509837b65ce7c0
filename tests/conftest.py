"""pytest configuration: the ``gpu`` marker and shared helpers.

``-m "not gpu"`` runs on a CPU-only box (oracle vs golden vectors, host logic,
C-ABI symbol table, gloo world_size-2 sharding); ``-m gpu`` runs the parity
tests proper through the C ABI on a B200.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


# The callers' tests build on the hot path's: run the hot path first, then the tile
# reader, the optimizer and the scripts (which use both) last, so that a failure
# is reported where it originates (pytest -x stops early).
_LATE = {"test_tiles.py": 1, "test_optim.py": 2, "test_scripts.py": 3}


def pytest_collection_modifyitems(config, items):
    import torch
    items.sort(key=lambda it: _LATE.get(os.path.basename(str(it.fspath)), 0))   # stable
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
