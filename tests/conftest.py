import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _debug_bounds_guard_check(request):
    """SEQDIFF_DEBUG_BOUNDS=1 (debug build of the library, the stand-in for compute-sanitizer): after every GPU test all workspace
    guard bands must be intact.  A no-op with the product build and for CPU tests."""
    yield
    if os.environ.get("SEQDIFF_DEBUG_BOUNDS") != "1" or request.node.get_closest_marker("gpu") is None:
        return
    import ctypes
    import seqdiff_b200 as sd
    nb, nk = ctypes.c_int(0), ctypes.c_int(0)
    rc = sd.lib().seqdiff_debug_check_guards(ctypes.byref(nb), ctypes.byref(nk), None)
    assert rc == 0 and nk.value == 0, f"{nk.value} of {nb.value} guard bands overwritten: {sd.lib().seqdiff_last_error()}"
