"""Builds libssdbox.so (the C-ABI CUDA library) in-tree for sm_100a.

    python object-detection-pytorch_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The library links the static CUDA runtime and has no
dependency on torch / python: it is the drop-in boundary described in include/ssdbox.h.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "ssdbox", "lib")
OUT = os.path.join(OUT_DIR, "libssdbox.so")
STAMP = OUT + ".stamp"

SOURCES = ["abi.cu", "boxops.cu", "match.cu", "loss.cu", "detect.cu", "evalpost.cu", "heads.cu", "voceval.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-cudart", "static",
]


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "ssdbox.h")]
    for f in files:
        h.update(os.path.relpath(f, ROOT).encode())     # relative: the stamp stays valid when the tree is copied (GPU box)
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=True):
    os.makedirs(OUT_DIR, exist_ok=True)
    dig = _digest()
    if not force and os.path.isfile(OUT) and os.path.isfile(STAMP) and open(STAMP).read().strip() == dig:
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC] + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(STAMP, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
