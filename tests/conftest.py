import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _cuda_device_present() -> bool:
    if os.environ.get("KB_FORCE_GPU_TESTS"):
        return True
    return os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need a B200 and the in-tree CUDA library: on a box without a CUDA device they are skipped (with the reason),
    not errored.  On a GPU box nothing is skipped - a missing library fails loudly there (kalibr_b200.capi.load_library)."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box (the B200 hot path has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle_api

    oracle_api.build()
    return oracle_api
