import os
import sys

import pytest

# several contexts with three streams each live in one process in the multi-context tests: give every stream its own
# hardware queue so that a polling kernel can never sit in front of the kernel it is waiting for
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import build as obuild
    obuild.build_oracle()
    from oracle.binding import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def product_lib():
    """The CUDA library must already be built (it travels to the GPU box prebuilt); build it here if nvcc exists."""
    from openclrenderer_b200 import _build, rr
    if not os.path.exists(rr.library_path()):
        _build.build_product()
    return rr.load_library()
