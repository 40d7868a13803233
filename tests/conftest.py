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


def _have_gpu():
    """nvidia-smi lists a device (no CUDA context is created in the pytest process for this check)."""
    import shutil
    import subprocess
    smi = shutil.which("nvidia-smi")
    if not smi:
        return False
    try:
        return subprocess.run([smi, "-L"], capture_output=True, text=True, timeout=30).stdout.count("GPU ") > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """a plain `pytest tests/` on a GPU-less host skips the gpu-marked tests instead of failing in rr_create;
    with `-m gpu` they run regardless, so a GPU box without a usable device fails loudly."""
    if "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or ""):
        return
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this host (gpu-marked tests run on the B200 box: pytest -m gpu)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


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
