import faulthandler
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# A GPU test that stops making progress must not take the rest of the suite with it (round 1: one hanging kernel hid
# every test collected after it).  Per-test limit through pytest-timeout when it is installed (thread method: a hang
# inside a CUDA call never returns to Python, so a signal handler would not run), plus a traceback dump shortly before.
GPU_TEST_TIMEOUT_S = 420


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu through gpurun)")
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout)")


def pytest_collection_modifyitems(config, items):
    has_timeout = config.pluginmanager.hasplugin("timeout")
    for item in items:
        if "gpu" in item.keywords and has_timeout and not any(m.name == "timeout" for m in item.iter_markers()):
            item.add_marker(pytest.mark.timeout(GPU_TEST_TIMEOUT_S, method="thread"))


@pytest.fixture(autouse=True)
def _dump_traceback_when_stuck(request):
    if "gpu" in request.keywords:
        faulthandler.dump_traceback_later(GPU_TEST_TIMEOUT_S - 20, exit=False)
        yield
        faulthandler.cancel_dump_traceback_later()
    else:
        yield


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle

    oracle.build()
    return oracle
