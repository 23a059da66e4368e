#!/usr/bin/env python
"""GPU check / timing of the multifrontal LU on refined meshes (large fronts take the group kernels).

    MESH_N=128 python tools/check_bigfront.py            # Newton solve, residuals, line items
    MESH_N=32 OCP_MF_BIG=96 python tools/check_bigfront.py   # force the large-front path on the reference mesh

Prints the Newton history of a zero-initialised forward solve with the Pipeline_limits control, the true residual
norm of the final state re-assembled without boundary rows, and the factor / solve times per call.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocp_b200  # noqa: E402,F401
from ocp_b200.capi import Context  # noqa: E402
from ocp_b200.fespace import TaylorHood  # noqa: E402
from ocp_b200.mesh import square_mesh  # noqa: E402
from ocp_b200.pipeline import initial_control  # noqa: E402

N = int(os.environ.get("MESH_N", "128"))
t0 = time.time()
V = TaylorHood(square_mesh(N))
t_space = time.time() - t0
dev = torch.device("cuda")
t0 = time.time()
ctx = Context(V, 1.0, 0.005, 200, (1.0, 1.0))
t_ctx = time.time() - t0
f = torch.from_numpy(initial_control(V, "PL")).to(dev)
w = torch.zeros(V.ndofs, dtype=torch.float64, device=dev)
out = {"mesh": N, "ndofs": V.ndofs, "nnz": int(V.csr_col.size), "space_s": t_space, "create_s": t_ctx}
ctx.set_profiling(True)
for rep in range(int(os.environ.get("REPS", "2"))):
    ctx.reset_solver_stats()
    torch.cuda.synchronize()
    t0 = time.time()
    its, hist = ctx.forward_solve(f, w, True)
    torch.cuda.synchronize()
    st = ctx.solver_stats()
    out.update({"newton_its": its, "residuals": [float(h) for h in hist[: its + 1]], "forward_solve_s": time.time() - t0,
                "factor_ms_each": st["factor_ms"] / max(st["n_factor"], 1), "solve_ms_each": st["solve_ms"] / max(st["n_solve"], 1),
                "assemble_ms": st["assemble_ms"], "n_factor": st["n_factor"], "n_solve": st["n_solve"]})
    print(json.dumps(out), flush=True)
ctx.set_profiling(False)
# independent check of the solution: residual of F(w) = 0 assembled afresh (boundary rows applied)
res = torch.empty(V.ndofs, dtype=torch.float64, device=dev)
ctx.assemble_forward(w, f, None, res, True)
print("final residual norm", float(res.norm()), "|w|", float(w.norm()), flush=True)
# adjoint-type solves: transposed factors against the assembled adjoint matrix
g = torch.zeros(V.mesh.num_vertices, 4, dtype=torch.float64, device=dev)
ctx.project_grad(w, g)
print("grad projection |g|", float(g.norm()), flush=True)
b = torch.zeros(2 * V.num_nodes + 2, dtype=torch.float64, device=dev)
b[: 2 * V.num_nodes] = torch.from_numpy(np.random.default_rng(0).standard_normal(2 * V.num_nodes)).to(dev)
z = torch.empty(V.ndofs, dtype=torch.float64, device=dev)
t0 = time.time()
ctx.adjoint_solve(w, b, z)
torch.cuda.synchronize()
print("adjoint solve ok |z|", float(z.norm()), "s", time.time() - t0, flush=True)
if os.environ.get("SAVE"):
    np.save(os.environ["SAVE"], torch.cat([w, z]).cpu().numpy())
if os.environ.get("COMPARE") and os.path.exists(os.environ["COMPARE"]):
    ref = np.load(os.environ["COMPARE"])
    cur = torch.cat([w, z]).cpu().numpy()
    n = V.ndofs
    print("vs reference file: w rel diff", float(np.linalg.norm(cur[:n] - ref[:n]) / np.linalg.norm(ref[:n])),
          "z rel diff", float(np.linalg.norm(cur[n:] - ref[n:]) / np.linalg.norm(ref[n:])), flush=True)
