"""Build recipes (in-tree, offline): the sm_100a product library, the CPU oracle, and — when the reference tree is
present — oracle/_ref (cl2.cl's own source ranges compiled as C++ through oracle/cl_shim.h).

Only the first is product. The other two are checkers; building a checker is not using it.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "openclrenderer_b200", "csrc")
LIB_PRODUCT = os.path.join(ROOT, "openclrenderer_b200", "librr_b200.so")
LIB_ORACLE = os.path.join(ROOT, "oracle", "liboracle.so")
LIB_REF = os.path.join(ROOT, "oracle", "_ref", "libcl2ref.so")
REFERENCE = "/root/reference"

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


def build_oracle(force=False):
    """g++ -O2 -ffp-contract=off -fopenmp -> oracle/liboracle.so (test infrastructure)"""
    srcs = [os.path.join(ROOT, "oracle", "oracle.cpp"), os.path.join(ROOT, "include", "rr.h")]
    if not force and _newer(LIB_ORACLE, srcs):
        return LIB_ORACLE
    _run(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-o", LIB_ORACLE, srcs[0]])
    return LIB_ORACLE


def build_ref(force=False):
    """oracle/_ref/libcl2ref.so from /root/reference/cl2.cl (only where the reference tree exists)."""
    if not os.path.isdir(REFERENCE):
        return LIB_REF if os.path.exists(LIB_REF) else None
    script = os.path.join(ROOT, "oracle", "build_ref.py")
    if not os.path.exists(script):
        return None
    srcs = [script, os.path.join(ROOT, "oracle", "cl_shim.h"), os.path.join(ROOT, "oracle", "ref_driver.cpp")]
    if not force and _newer(LIB_REF, [s for s in srcs if os.path.exists(s)]):
        return LIB_REF
    _run([sys.executable, script])
    return LIB_REF


def build_all(force=False):
    return build_product(force), build_oracle(force), build_ref(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
