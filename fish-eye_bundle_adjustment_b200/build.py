"""In-tree build of the CUDA library (sm_100a only) and of the MEX-gateway compile check.

``nvcc`` cross-compiles without a GPU; the shared object lands next to this file
(``libfeba.so``) so that it travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfeba.so")
SOURCES = ("feba_api.cu", "feba_kernels.cu", "feba_assemble.cu", "feba_chol.cu", "feba_dist.cu", "feba_green.cu",
           "feba_pack.cpp")            # the last one is host-only C++ (problem build, include/feba_pack.h)
HEADERS = ("feba_dev.h", "feba_kernels.h", "feba_model.cuh", "feba_sparse.h", "feba_order.h", "feba_chunks.h", os.path.join("..", "..", "include", "feba.h"),
           os.path.join("..", "..", "include", "feba_pack.h"))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall,-pthread,-Wno-unknown-pragmas", "--use_fast_math=false"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return exe


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libfeba.so for sm_100a.  Returns the path."""
    if not force and not needs_build():
        return LIB
    objs, procs = [], []
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    for src in SOURCES:                      # translation units compile concurrently
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        cmd = [_nvcc(), *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs, "-lcudart", "-ldl", "-lpthread"],
                   check=True)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
