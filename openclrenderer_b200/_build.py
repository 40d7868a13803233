"""Build recipe of the product (in-tree, offline): csrc/*.cu -> librr_b200.so for sm_100a.
(The checkers have their own recipe in oracle/build.py.)"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "openclrenderer_b200", "csrc")
LIB_PRODUCT = os.path.join(ROOT, "openclrenderer_b200", "librr_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r


def build_product(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a … -> openclrenderer_b200/librr_b200.so"""
    srcs = [os.path.join(CSRC, f) for f in ("rr_api.cu", "rr_kernels.cuh", "rr_math.cuh")] + [os.path.join(ROOT, "include", "rr.h")]
    if not force and _newer(LIB_PRODUCT, srcs):
        return LIB_PRODUCT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PRODUCT, os.path.join(CSRC, "rr_api.cu")]
    r = _run(cmd)
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB_PRODUCT


if __name__ == "__main__":
    print(build_product(force="--force" in sys.argv, verbose="-v" in sys.argv))
