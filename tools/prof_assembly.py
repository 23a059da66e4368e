#!/usr/bin/env python
"""Profiling driver: Taylor-Hood cell assembly (Newton matrix + residual, adjoint operator) on a refined square mesh
(cfg5 meshes); run plain for CUDA-event times, or under ncu for DRAM bytes / FP64 pipe.

    MESH_N=256 REPS=5 python tools/prof_assembly.py

Algorithmic bytes per cell: 48 (geometry) + 60 (15 dof ids) + 120 (coefficients)
+ the cell's share of the CSR values (nnz * 8 / nc, written once) + 15 * 8 residual ~ 1.3 KB.  The default gather
assembly adds 360 B per cell of pair tables, the atomic scatter kernels (OCP_ASSEMBLY=atomic) a 900 B slot table.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocp_b200  # noqa: E402,F401
from ocp_b200.capi import Context  # noqa: E402
from ocp_b200.fespace import TaylorHood  # noqa: E402
from ocp_b200.mesh import square_mesh  # noqa: E402

N = int(os.environ.get("MESH_N", "256"))
REPS = int(os.environ.get("REPS", "5"))
V = TaylorHood(square_mesh(N))
ctx = Context(V, 1.0, 0.005, 200, (1.0, 1.0))
dev = torch.device("cuda")
rng = np.random.default_rng(0)
nc, nnz, n = V.mesh.num_cells, int(V.csr_col.size), V.ndofs
w = torch.from_numpy(rng.standard_normal(n) * 0.1).to(dev)
f = torch.zeros(V.num_nodes, 2, dtype=torch.float64, device=dev)
vals = torch.empty(nnz, dtype=torch.float64, device=dev)
res = torch.empty(n, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
alg = nc * (48 + 60 + 120 + 120) + nnz * 8.0     # algorithmic bytes per forward assembly (values written once)
print("gather assembly" if ctx.option("gather_assembly") else "atomic scatter assembly", flush=True)
out = []
for rep in range(REPS):
    flush.zero_()
    ev[0].record()
    ctx.assemble_forward(w, f, vals, res, False)
    ev[1].record()
    flush.zero_()
    ev[2].record()
    ctx.assemble_adjoint(w, vals, False)
    ev[3].record()
    torch.cuda.synchronize()
    tf, ta = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])
    out.append((tf, ta))
    print(f"N={N} cells {nc} nnz {nnz}: forward {tf:.3f} ms ({nc / tf / 1e3:.1f} Mcell/s, {alg / tf / 1e6:.0f} GB/s alg.)"
          f" | adjoint {ta:.3f} ms", flush=True)
tf, ta = min(o[0] for o in out), min(o[1] for o in out)
print(json.dumps({"mesh": N, "cells": nc, "nnz": nnz, "forward_ms": tf, "adjoint_ms": ta,
                  "algorithmic_bytes": alg, "forward_gbs": alg / tf / 1e6,
                  "mode": "gather" if ctx.option("gather_assembly") else "atomic",
                  "note": "times include the facet kernel (and, in atomic mode, the two memsets) of ocp_assemble_forward"}))
