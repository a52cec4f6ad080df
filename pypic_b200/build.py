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
OBJ = os.path.join(HERE, "_obj")
SOURCES = ["grid_kernels.cu", "dd_kernels.cu", "slab_kernels.cu", "periodic_kernels.cu", "gc_kernels.cu", "init_kernels.cu", "abi_host.cu", "mt_host.cpp"]
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


def _digest(paths, extra=""):
    h = hashlib.sha256()
    for f in paths:
        with open(f, "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update((" ".join(NVCC_FLAGS) + os.environ.get("PIC_NVCC_EXTRA", "") + extra).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """One object per source (recompiled only when it, a header of csrc/ or the ABI header changed;
    the objects compile in parallel), then one link."""
    from concurrent.futures import ThreadPoolExecutor
    headers = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "pic_b200.h"))
    os.makedirs(OBJ, exist_ok=True)
    extra = os.environ.get("PIC_NVCC_EXTRA", "").split()
    compile_flags = [f for f in NVCC_FLAGS if f not in ("-shared",)]
    jobs, objs = [], []
    digs = [_digest([os.path.join(CSRC, src)] + headers) for src in SOURCES]
    link_dig = hashlib.sha256("".join(digs).encode()).hexdigest()
    if not force and os.path.isfile(LIB) and os.path.isfile(STAMP) and open(STAMP).read().strip() == link_dig:
        return LIB          # the library matches the sources (the objects need not exist, e.g. on the GPU box)
    for src, dig in zip(SOURCES, digs):
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src + ".o")
        objs.append(obj)
        stamp = obj + ".stamp"
        fresh = os.path.isfile(obj) and os.path.isfile(stamp) and open(stamp).read().strip() == dig
        if force or not fresh:
            jobs.append((path, obj, stamp, dig))

    def compile_one(job):
        path, obj, stamp, dig = job
        cmd = [_nvcc()] + compile_flags + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode == 0:
            with open(stamp, "w") as fh:
                fh.write(dig)
        return path, res

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for path, res in ex.map(compile_one, jobs):
                if res.returncode != 0:
                    sys.stderr.write(res.stdout + res.stderr)
                    raise RuntimeError("nvcc failed compiling %s" % path)
                if verbose:
                    sys.stderr.write(res.stderr)
    if True:
        cmd = [_nvcc()] + NVCC_FLAGS + objs + ["-o", LIB]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed linking libpic_b200.so")
        with open(STAMP, "w") as fh:
            fh.write(link_dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
