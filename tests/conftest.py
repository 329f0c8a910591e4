import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle_kind():
    """The strongest oracle available: the compiled reference if its prebuilt library travelled, else the port."""
    from oracle import oracle as O
    if O.available("ref"):
        return "ref"
    if not O.available("port"):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    return "port"
