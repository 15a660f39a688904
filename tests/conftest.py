import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests call the CUDA library (no CPU fallback): skip them where there is no device"""
    try:
        import torch

        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        # no GPU test may stall a lease: a chain that stops accepting (the reference has such cases) or a
        # lost peer must fail the test, not hang it (pytest-timeout, when installed)
        if config.pluginmanager.hasplugin("timeout"):
            for item in items:
                if "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
                    item.add_marker(pytest.mark.timeout(420))
        return
    skip = pytest.mark.skip(reason="needs CUDA (B200); gravinv3dhmc_b200 has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    class G:
        def __getitem__(self, name):
            return np.load(os.path.join(GOLDEN, name + ".npz"))

    return G()
