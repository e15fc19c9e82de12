import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A kernel that waits on a barrier forever must fail the test, not hang the GPU box: every GPU test gets a
    time limit (pytest-timeout, when it is installed)."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(420))


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_py
    return oracle_py.load()


@pytest.fixture(scope="session")
def rc_ctx():
    """One libRNAcode_cuda context on cuda:0 for the whole GPU session."""
    from rnacode_b200 import capi
    ctx = capi.Context(0)
    yield ctx
    ctx.close()
