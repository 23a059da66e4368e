"""Builds libocp_b200.so (sm_100a) in-tree with nvcc.  ``python -m ocp_b200.build [--force] [-v]`` or ``build()``.

Every translation unit is compiled to an object file under ``csrc/_obj/`` and re-compiled only when the SHA-256 of
its source, of every header and of the compiler flags changes (recorded next to the object); the library is
re-linked when any object changed.  The hash of everything that went into the library is stored in
``libocp_b200.so.sha256`` so that a stale binary can never be mistaken for the current sources (mtimes are not
trusted: checkouts and the snapshot shipped to the GPU box do not preserve them).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_OBJ = os.path.join(_CSRC, "_obj")
SOURCES = ["capi.cu", "buoy_kernels.cu", "fe_kernels.cu", "sparse_solver.cu", "multifrontal.cu", "comm.cu",
           "host_lu.cpp", "multifrontal.cpp"]
HEADERS = ["element_math.cuh", "kernels.cuh", "sparse_solver.cuh", "host_lu.hpp", "multifrontal.hpp",
           "multifrontal.cuh", "comm.cuh", "../../include/ocp_b200.h"]
LIB = os.path.join(_HERE, "libocp_b200.so")
STAMP = LIB + ".sha256"
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]
LINK = ["-lcusolver", "-lcusparse", "-ldl"]


def _sha(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _header_paths():
    return [os.path.join(_CSRC, f) for f in HEADERS]


def source_hash() -> str:
    """Hash of everything the library is built from (sources, headers, flags)."""
    return _sha([os.path.join(_CSRC, f) for f in SOURCES] + _header_paths(), " ".join(FLAGS + LINK))


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False, ptxas_log: str | None = None) -> str:
    """Compile what changed and link.  ``ptxas_log``: also write nvcc's ``-Xptxas -v`` output (registers, spills,
    shared memory of every kernel) to that file - forces a full re-compile."""
    if ptxas_log:
        force = True
    if not force and not needs_build():
        return LIB
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    nvcc = os.path.join(cuda, "bin", "nvcc")
    os.makedirs(_OBJ, exist_ok=True)
    flags = FLAGS + (["-Xptxas=-v"] if ptxas_log else [])
    hdr = _header_paths()

    def compile_one(src):
        path = os.path.join(_CSRC, src)
        obj = os.path.join(_OBJ, src + ".o")
        want = _sha([path] + hdr, " ".join(flags))
        stamp = obj + ".sha256"
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == want:
            return obj, ""
        cmd = [nvcc, *flags, "-c", path, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise subprocess.CalledProcessError(r.returncode, cmd)
        with open(stamp, "w") as fh:
            fh.write(want)
        return obj, f"==== {src}\n{r.stdout}{r.stderr}"

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    if ptxas_log:
        os.makedirs(os.path.dirname(os.path.abspath(ptxas_log)), exist_ok=True)
        with open(ptxas_log, "w") as fh:
            fh.write("# nvcc " + " ".join(flags) + "\n" + "".join(t for _, t in results))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", LIB, *LINK,
           "-Xlinker", f"-rpath={cuda}/lib64"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(STAMP, "w") as fh:
        fh.write(source_hash())
    return LIB


if __name__ == "__main__":
    log = None
    if "--ptxas" in sys.argv:
        log = sys.argv[sys.argv.index("--ptxas") + 1]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, ptxas_log=log))
