"""Builds pypic_b200/libpic_b200.so (hand-written sm_100a CUDA + the C ABI of
include/pic_b200.h) in-tree with nvcc.  No torch, no JIT cache: the .so travels
with the repository snapshot to the GPU box.

    python -m pypic_b200.build [--force]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpic_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")
SOURCES = ["grid_kernels.cu", "dd_kernels.cu", "periodic_kernels.cu", "gc_kernels.cu", "init_kernels.cu", "abi_host.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                 # parity: never contract a*b+c (see csrc/common.cuh)
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isfile(c) or c == "nvcc"):
            return c
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(ROOT, "include", "pic_b200.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update((" ".join(NVCC_FLAGS) + os.environ.get("PIC_NVCC_EXTRA", "")).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    dig = _digest()
    if not force and os.path.isfile(LIB) and os.path.isfile(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    extra = os.environ.get("PIC_NVCC_EXTRA", "").split()
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libpic_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
