#!/usr/bin/env python
"""Profiling driver: buoy forward / backward sweeps at sweep size (run under ncu / gpurun)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocp_b200  # noqa: E402
from ocp_b200.fespace import TaylorHood  # noqa: E402
from ocp_b200.mesh import square_mesh  # noqa: E402
from ocp_b200.pipeline import OCP, Parameters  # noqa: E402

K = int(os.environ.get("K", str(1 << 20)))
N = int(os.environ.get("MESH_N", "32"))
V = TaylorHood(square_mesh(N))
rng = np.random.default_rng(0)
x0 = np.stack([rng.uniform(0.1, 1.9, K), rng.uniform(0.1, 1.9, K)], 1)
ocp = OCP(V, Parameters(), x0, np.zeros((K, 200, 2)))
if N == 32:
    w = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "fields.npz"))["velocity_100"]).cuda()
else:
    from ocp_b200.pipeline import initial_control
    w = ocp.forward_solve(torch.from_numpy(initial_control(V, "PL")).cuda()).d_w
ocp.ctx.project_grad(w, ocp.d_g)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for rep in range(int(os.environ.get("REPS", "3"))):
    ev[0].record()
    ocp._primal(w, ocp.d_x, ocp.d_u, ocp.d_mask)
    ev[1].record()
    if rep == 0:
        ocp.d_ud.copy_(1.1 * ocp.d_u)
    ocp.d_acc.zero_()
    ev[2].record()
    ocp.ctx.buoy_adjoint_scatter(ocp.d_vel, ocp.d_g, K, ocp.d_x, ocp.d_u, ocp.d_ud, ocp.d_mask, ocp.d_parked, None, ocp.d_acc)
    ev[3].record()
    torch.cuda.synchronize()
    tf, tb = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])
    st = K * 200
    print(f"K={K} N={N}: forward {tf:.3f} ms {32 * st / tf / 1e6:.0f} GB/s | backward {tb:.3f} ms {48 * st / tb / 1e6:.0f} GB/s | masked {int(ocp.d_mask.sum())}")
