import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib_built():
    """Build libvitb200.so in-tree if nvcc is here and it is stale/missing."""
    from vit_flax_b200 import _lib
    if os.path.exists("/usr/local/cuda/bin/nvcc"):
        from vit_flax_b200.build import build
        build()
    if not _lib.LIB_PATH.exists():
        pytest.fail("libvitb200.so is not built (python -m vit_flax_b200.build)")
    return _lib.load()
