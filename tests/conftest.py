import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """A device kernel that never finishes (a spin-wait that is never satisfied) would hold the GPU box
    until the outer limit: every GPU test gets a hard limit.  method="thread": the process is blocked
    inside a C call, where the signal-based timeout cannot fire."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(900, method="thread"))


@pytest.fixture(scope="session")
def params():
    import helpers
    return helpers.load_params()


@pytest.fixture(scope="session")
def scoring(params):
    import helpers
    return helpers.load_scoring(params)
