"""Build recipes of the checkers (TEST INFRASTRUCTURE): the CPU oracle, and — where the reference tree is present —
oracle/_ref/cl2.cl.gz, the input of oracle/ref_opencl.py. Building a checker is not using it."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_ORACLE = os.path.join(ROOT, "oracle", "liboracle.so")
LIB_REF = os.path.join(ROOT, "oracle", "_ref", "cl2.cl.gz")
REFERENCE = "/root/reference"


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


def build_oracle(force=False):
    """g++ -O2 -ffp-contract=off -fopenmp -> oracle/liboracle.so (test infrastructure)"""
    srcs = [os.path.join(ROOT, "oracle", "oracle.cpp"), os.path.join(ROOT, "include", "rr.h")]
    if not force and _newer(LIB_ORACLE, srcs):
        return LIB_ORACLE
    _run(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-o", LIB_ORACLE, srcs[0]])
    return LIB_ORACLE


def build_ref(force=False):
    """oracle/_ref/cl2.cl.gz from /root/reference/cl2.cl (only where the reference tree exists): the input of
    oracle/ref_opencl.py, which runs the reference's own kernels through the NVIDIA OpenCL ICD on the GPU box."""
    script = os.path.join(ROOT, "oracle", "build_ref.py")
    if not os.path.isdir(REFERENCE):
        return LIB_REF if os.path.exists(LIB_REF) else None
    if not force and _newer(LIB_REF, [script, os.path.join(REFERENCE, "cl2.cl")]):
        return LIB_REF
    _run([sys.executable, script])
    return LIB_REF


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv), build_ref())
