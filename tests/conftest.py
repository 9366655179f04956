import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_water_rhs.npz"))


# A fuse for test modules that set STOP_AFTER_TIMEOUT = True (code not yet run on hardware): after the first test of such a module
# that ends in a pytest-timeout, the rest of the module is skipped instead of waiting for the same hang again and again.
_timed_out_modules = set()


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_makereport(item, call):
    outcome = yield
    rep = outcome.get_result()
    if getattr(item.module, "STOP_AFTER_TIMEOUT", False) and call.excinfo is not None and "Timeout" in str(call.excinfo.value)[:200]:
        _timed_out_modules.add(item.module.__name__)
    return rep


def pytest_runtest_setup(item):
    if item.module.__name__ in _timed_out_modules:
        pytest.skip("an earlier test of this module timed out")
