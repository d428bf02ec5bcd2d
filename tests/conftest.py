import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def wam():
    """The product package (directory name has a hyphen, hence importlib)."""
    build = importlib.import_module("webaudio-modem_b200.build")
    build.build()
    return importlib.import_module("webaudio-modem_b200")


@pytest.fixture(scope="session")
def gpu_wam(wam):
    import ctypes

    n = ctypes.c_int(0)
    rc = wam.lib().wam_device_count(ctypes.byref(n))
    if rc != 0 or n.value == 0:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the GPU box (no CPU fallback exists)")
    return wam
