"""Recipe for oracle/_ref/ (git-ignored, travels to the GPU box): a gzip of the reference's cl2.cl, taken from where
it lies under /root/reference. oracle/ref_opencl.py feeds it, unmodified, to the NVIDIA OpenCL compiler on the GPU box.
Nothing from the reference is committed to this repository."""
import gzip
import os
import shutil
import sys

SRC = "/root/reference/cl2.cl"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
OUT = os.path.join(OUT_DIR, "cl2.cl.gz")


def main():
    if not os.path.exists(SRC):
        print("reference tree absent; keeping", OUT if os.path.exists(OUT) else "nothing")
        return 0
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(SRC, "rb") as f, gzip.GzipFile(OUT, "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
