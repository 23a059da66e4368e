"""Builds libocp_b200.so (sm_100a) in-tree with nvcc.  ``python -m ocp_b200.build`` or ``build()``."""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["capi.cu", "buoy_kernels.cu", "fe_kernels.cu", "sparse_solver.cu", "multifrontal.cu", "host_lu.cpp",
           "multifrontal.cpp"]
HEADERS = ["element_math.cuh", "kernels.cuh", "sparse_solver.cuh", "host_lu.hpp", "multifrontal.hpp",
           "multifrontal.cuh", "../../include/ocp_b200.h"]
LIB = os.path.join(_HERE, "libocp_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(_HERE, "csrc", f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = [
        os.path.join(cuda, "bin", "nvcc"), "-O3", "-std=c++17",
        "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
        "-Xcompiler", "-fPIC", "-shared",
        *[os.path.join(_HERE, "csrc", f) for f in SOURCES],
        "-o", LIB, "-lcusolver", "-lcusparse",
        "-Xlinker", f"-rpath={cuda}/lib64",
    ]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
