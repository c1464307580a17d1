import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def dsat_lib():
    from diffusionsat_b200 import build, _lib
    build.build()
    return _lib.load_library()


@pytest.fixture()
def ctx(dsat_lib):
    from diffusionsat_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def pytest_report_header(config):
    try:
        from diffusionsat_b200 import _lib
        info = _lib.load_library().dsat_build_info()
        return "libdsat build: tcgen05=%d assert=%d" % (info & 1, (info >> 1) & 1)
    except Exception as exc:        # library not built yet
        return "libdsat build: not loaded (%s)" % type(exc).__name__
