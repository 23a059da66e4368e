#!/usr/bin/env python
"""Profiling driver: two forward Navier-Stokes solves on the 32x32 square (run under ncu / gpurun)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocp_b200  # noqa: E402
from ocp_b200.fespace import TaylorHood  # noqa: E402
from ocp_b200.mesh import square_mesh  # noqa: E402
from ocp_b200.pipeline import OCP, Parameters, initial_control  # noqa: E402

V = TaylorHood(square_mesh(int(os.environ.get("MESH_N", "32"))))
ocp = OCP(V, Parameters(), np.array([[0.5, 0.5]]), np.zeros((1, 200, 2)))
f = torch.from_numpy(initial_control(V, "PL")).cuda()
ocp.ctx.set_profiling(True)
for _ in range(int(os.environ.get("REPS", "2"))):
    ocp.ctx.reset_solver_stats()
    ocp.forward_solve(f)
    torch.cuda.synchronize()
    print(ocp.last_newton_its, ocp.ctx.solver_stats())
