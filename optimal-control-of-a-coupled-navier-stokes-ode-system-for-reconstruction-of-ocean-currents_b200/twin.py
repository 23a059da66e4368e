"""Twin experiment that produces the measurement data ``u_d`` / ``x_0``.

Reference: ``plotting/ud_construction_pipeline.py`` - a Navier-Stokes solve on the square with Dirichlet data on
the whole boundary (``:95-106``: no-slip on y=0,2; ``inflow`` on x=0,2, nodal P2 interpolation, applied last so the
corners take the inflow value; p = 0 on the facets of x=0), then K buoys advected through the field
(``:116-151``) and saved as ``u_d_array.npy`` / ``x_0_array.npy`` (``:264-268``).  The same forward solver, buoy
kernel and C ABI as the OCP are used; only the Dirichlet set differs (``ocp_set_dirichlet``).
"""
from __future__ import annotations

import copy
from typing import Callable, Tuple

import numpy as np
import torch

from . import capi
from .fespace import TaylorHood
from .pipeline import OCP, Parameters, State


def twin_dirichlet(V: TaylorHood, inflow: Callable) -> Tuple[np.ndarray, np.ndarray]:
    """Constrained dofs and values of the twin problem on the square [0,2]^2 (later BCs override earlier ones)."""
    m = V.mesh
    xy = V.node_coords
    near = lambda a, b: np.abs(a - b) < 3e-16 * 1e2            # dolfin's near(): DOLFIN_EPS-scaled tolerance
    vals = {}
    noslip = near(xy[:, 1], 0.0) | near(xy[:, 1], 2.0)
    for n in np.nonzero(noslip)[0]:
        vals[int(V.dof_ux[n])] = 0.0
        vals[int(V.dof_uy[n])] = 0.0
    io = near(xy[:, 0], 0.0) | near(xy[:, 0], 2.0)
    ux, uy = inflow(xy[io, 0], xy[io, 1])
    for n, a, b in zip(np.nonzero(io)[0], np.broadcast_to(ux, io.sum()), np.broadcast_to(uy, io.sum())):
        vals[int(V.dof_ux[n])] = float(a)
        vals[int(V.dof_uy[n])] = float(b)
    pin = m.coords[:, 0] < 3e-16
    for v in np.nonzero(pin)[0]:
        vals[int(V.dof_p[v])] = 0.0
    dofs = np.array(sorted(vals), dtype=np.int32)
    return dofs, np.array([vals[int(d)] for d in dofs])


class TwinExperiment:
    def __init__(self, V: TaylorHood, viscosity: float = 1.0, params: Parameters = Parameters(),
                 inflow: Callable = lambda x, y: (0.1 + 0 * x, 0 * x), device=None):
        if V.mesh.l_shape:
            raise ValueError("the reference's twin experiment is defined on the square only")
        # same space, but no Gamma_1 facets (the boundary markers are never set in the reference script, so its
        # ds(1) terms are empty) and the twin Dirichlet set
        Vt = copy.copy(V)
        Vt.g1_cell, Vt.g1_local = V.g1_cell[:0], V.g1_local[:0]
        Vt.g1_nodes, Vt.g1_len, Vt.g1_normal = V.g1_nodes[:0], V.g1_len[:0], V.g1_normal[:0]
        self.dofs, self.vals = twin_dirichlet(V, inflow)
        Vt.dirichlet_dofs = self.dofs
        p = Parameters(viscosity=viscosity, t0=params.t0, T=params.T, dt=params.dt, alpha=params.alpha)
        self.V, self.params, self.device = Vt, p, device
        self._ocp = None

    def solve(self, x0: np.ndarray):
        """Returns (w, x, u): the twin state (numpy, W numbering) and the trajectories / velocities (K,nt,2) that the
        reference stores as ``x_0_array.npy`` / ``u_d_array.npy``."""
        x0 = np.ascontiguousarray(x0, np.float64).reshape(-1, 2)
        K = x0.shape[0]
        ocp = OCP(self.V, self.params, x0, np.zeros((K, self.params.nt, 2)), device=self.device)
        self._ocp = ocp
        ocp.ctx.set_dirichlet(self.dofs, self.vals)
        f = torch.zeros((self.V.num_nodes, 2), device=ocp.device, dtype=torch.float64)
        st = ocp.forward_solve(f)
        self.newton_its, self.res_hist = ocp.last_newton_its, ocp.last_res_hist
        mask = np.zeros(K)
        x, u = ocp.solve_primal_ode(st, mask)
        w = st.vector()
        self.norms = ocp.field_norms(st)
        ocp.close()
        return w, x, u
