import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture
def libopt():
    """Setter for the library's kernel-selection options (include/b200pose.h "options"); every option is restored
    when the test ends.  usage: libopt("conv_mode", 1)"""
    from rnnpose_b200 import ops
    saved = {n: ops.get_option(n) for n in ops.option_names()}

    def set_(name, value):
        ops.set_option(name, int(value))
    yield set_
    for n, v in saved.items():
        ops.set_option(n, v)
