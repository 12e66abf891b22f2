import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kref():
    from oracle import kref as m
    if not m.available():
        pytest.skip("oracle/_ref/libkaori_ref.so not built")
    return m


@pytest.fixture(scope="session")
def port():
    from oracle import port as m
    if not m.available():
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    return m
