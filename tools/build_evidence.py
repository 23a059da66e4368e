#!/usr/bin/env python
"""Build evidence for profiles/: register / spill / shared-memory table of every kernel (nvcc -Xptxas -v, compiled into
a scratch directory - the in-tree library is not touched) and the SASS instruction mix of the hot kernels
(cuobjdump -sass of the in-tree libocp_b200.so).

    python tools/build_evidence.py profiles/ptxas_r2.txt profiles/sass_r2_mix.txt
"""
import os
import re
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocp_b200  # noqa: E402,F401
from ocp_b200 import build as B  # noqa: E402

HOT = ("buoy_", "dense_apply", "assemble_gather", "mf_forward64", "mf_backward64", "mf_factor", "mf_leaf", "mf_big_factor",
       "mf_dinv64")
MNEMONICS = ("DFMA", "DMUL", "DADD", "LDG", "STG", "LDS", "STS", "ATOMG", "RED", "ATOMS", "UBLKCP", "SYNCS", "UCGABAR_ARV",
             "UCGABAR_WAIT", "BAR", "SHFL", "MUFU", "LDL", "STL", "ACQBULK", "PREEXIT")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    res = []
    for n in out[: len(names)]:
        n = n.replace("ocp::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("ocp::", "")
        n = re.sub(r"\(.*", "", n).replace("void ", "")
        res.append(n)
    return res


def ptxas_table(path):
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    tmp = tempfile.mkdtemp()
    srcs = [s for s in B.SOURCES if s.endswith(".cu")]

    def one(src):
        r = subprocess.run([nvcc, *B.FLAGS, "-Xptxas=-v", "-c", os.path.join(B._CSRC, src), "-o", os.path.join(tmp, src + ".o")],
                           capture_output=True, text=True)
        return r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=len(srcs)) as pool:
        text = "".join(pool.map(one, srcs))
    rows = []
    cur = None
    for line in text.split("\n"):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = {"name": m.group(1), "stack": 0, "st": 0, "ld": 0, "regs": 0, "smem": 0}
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            cur["stack"], cur["st"], cur["ld"] = map(int, m.groups())
        m = re.search(r"Used (\d+) registers", line)
        if m:
            cur["regs"] = int(m.group(1))
            m2 = re.search(r"(\d+) bytes smem", line)
            cur["smem"] = int(m2.group(1)) if m2 else 0
    names = demangle([r["name"] for r in rows])
    with open(path, "w") as fh:
        fh.write("# nvcc " + " ".join(B.FLAGS) + " -Xptxas=-v  (python tools/build_evidence.py), round 2, final state\n")
        fh.write("# kernel | registers | stack B | spill stores B | spill loads B | static smem B\n\n")
        for n, r in sorted(zip(names, rows), key=lambda x: x[0]):
            fh.write(f"{n:<72s} {r['regs']:4d} {r['stack']:5d} {r['st']:5d} {r['ld']:5d} {r['smem']:6d}\n")
    return len(rows)


def sass_mix(path):
    cuobjdump = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "cuobjdump")
    text = subprocess.run([cuobjdump, "-sass", B.LIB], capture_output=True, text=True).stdout
    funcs = {}
    cur = None
    for line in text.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = {"total": 0}
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        funcs[cur]["total"] += 1
        base = op.split(".")[0]
        key = None
        if base == "UCGABAR_ARV" or base == "UCGABAR_WAIT":
            key = base
        elif base in MNEMONICS:
            key = base
        elif base == "ATOM" or base == "ATOMG":
            key = "ATOMG"
        if key:
            funcs[cur][key] = funcs[cur].get(key, 0) + 1
    names = demangle(list(funcs))
    with open(path, "w") as fh:
        fh.write("# SASS instruction mix of the hot kernels (cuobjdump -sass libocp_b200.so, sm_100a), round 2, final state\n")
        fh.write("# UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, DFMA/DMUL/DADD = fp64 pipe, RED/ATOMG = global atomics,\n"
                 "# UCGABAR = cluster barrier, ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents (programmatic dependent launch)\n\n")
        for n, f in sorted(zip(names, funcs.values()), key=lambda x: x[0]):
            if not any(h in n for h in HOT):
                continue
            mix = " ".join(f"{k}={f[k]}" for k in MNEMONICS if f.get(k))
            fh.write(f"{n:<70s} total {f['total']:6d} | {mix}\n")
    return len(funcs)


if __name__ == "__main__":
    print("kernels:", ptxas_table(sys.argv[1]), "functions:", sass_mix(sys.argv[2]))
