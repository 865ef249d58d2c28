import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle (test infrastructure): build on demand."""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def gpu_capi():
    """The product library.  No skip: on a box without the CUDA build or a
    device this raises -- GPU tests must never pass on a fallback."""
    from fauxgl_b200 import build as fbuild
    from fauxgl_b200 import context
    if not os.path.exists(context.LIB_PATH):
        fbuild.build_library()
    lib = context.capi()
    assert lib.fgl_device_count() > 0, "no CUDA device: fauxgl_b200 has no CPU fallback"
    return lib
