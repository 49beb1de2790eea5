import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "stif-continuous-video-representation_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "timeout: per-test limit (pytest-timeout; ignored if the plugin is absent)")


def pytest_collection_modifyitems(config, items):
    """Every GPU test gets a 150 s limit (pytest-timeout, thread method: the process is torn down, which also tears down a
    hung CUDA context).  The persistent kernels synchronise with named barriers that have no timeout of their own; a
    protocol bug must fail one test fast, not stall the whole run."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(150, method="thread"))


@pytest.fixture(scope="session")
def built_lib():
    """Path of libstif_b200.so; builds it in-tree if it is missing (nvcc cross-compiles without a GPU)."""
    so = os.path.join(PKG, "lib", "libstif_b200.so")
    if not os.path.isfile(so):
        subprocess.check_call(["bash", os.path.join(PKG, "csrc", "build.sh")])
    return so


@pytest.fixture(scope="session")
def stif(built_lib):
    import stif_b200
    return stif_b200
