import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the oracle (checker) and the host VM (trace generator) are plain C++; build them on demand
    for d, lib in (("oracle", "liboracle.so"), ("distaff_b200/hostvm", "libdistaff_vm.so")):
        if not os.path.exists(os.path.join(ROOT, d, lib)):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, d)], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="session")
def fib13():
    from distaff_b200 import hostvm
    return hostvm.fibonacci(13)
