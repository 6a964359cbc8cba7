"""Builds libpvw_b200.so (hand-written CUDA kernels + the C ABI of include/pvw_b200.h) in-tree for sm_100a.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpvw_b200.so")
SOURCES = ["capi.cu", "ntt.cu", "mac.cu", "decode.cu", "wire.cu", "imma.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def _digest() -> str:
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for root in (CSRC, os.path.join(HERE, "..", "include"), os.path.join(HERE, "..", "examples")):
        for fn in sorted(os.listdir(root)):
            if fn.endswith((".cu", ".cuh", ".hpp", ".h", ".cpp")):
                with open(os.path.join(root, fn), "rb") as f:
                    h.update(fn.encode())
                    h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    os.makedirs(OBJ, exist_ok=True)

    def cc(src: str) -> str:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with cf.ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(cc, SOURCES))
    subprocess.check_call([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    # the C++ host-mirror example (include/pvw_b200.hpp) -- plain g++ against the C ABI
    example = os.path.join(HERE, "..", "examples", "pvw.cpp")
    if os.path.exists(example):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(HERE, "..", "include"), example, "-o",
                               os.path.join(OBJ, "pvw_example"), "-L", HERE, "-lpvw_b200", "-Wl,-rpath," + HERE, "-pthread"])
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
