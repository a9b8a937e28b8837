import os
import sys

# The in-process communicator (ranks = threads sharing one device and one CUDA context, tests/localranks.py) runs kernels that
# wait for kernels of another stream.  With CUDA's default LAZY module loading the first launch of a kernel can block until
# the kernels already running have finished -- a deadlock when one of those waits for the kernel being loaded (documented
# CUDA caveat).  EAGER loading must be chosen before the CUDA context exists, i.e. before any test module touches the library.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
