import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "face-recognition-platform_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import frp_b200
        return frp_b200.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: only auto-skip when the
    # user did not ask for gpu tests explicitly.
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_l2():
    return dict(np.load(os.path.join(GOLDEN, "l2_c1.npz")))


@pytest.fixture(scope="session")
def golden_cos():
    return dict(np.load(os.path.join(GOLDEN, "cos_small.npz")))


@pytest.fixture(scope="session")
def golden_align():
    return dict(np.load(os.path.join(GOLDEN, "align_small.npz")))
