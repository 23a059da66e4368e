"""Shared fixtures for the parity tests: problem set-up, golden data, and an oracle-side restatement of one
gradient-descent iteration (OCP_dolfin.py:309-450) built only from oracle/ functions."""
import json
import os
from functools import lru_cache

import numpy as np

import ocp_b200  # noqa: F401  (import alias of the hyphenated package directory)
from ocp_b200.fespace import TaylorHood
from ocp_b200.mesh import lshape_mesh, square_mesh
from oracle.buoy_oracle import BuoyOracle
from oracle.fe_oracle import FEOracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
H, NT = 0.005, 200
CENTER = np.array([1.0, 1.0])


@lru_cache(maxsize=None)
def square32():
    return TaylorHood(square_mesh(32))


@lru_cache(maxsize=None)
def lshape(m=12, jitter=0.2):
    return TaylorHood(lshape_mesh(m, jitter=jitter))


@lru_cache(maxsize=None)
def fields():
    f = np.load(os.path.join(GOLD, "fields.npz"))
    return {k: f[k] for k in f.files}


@lru_cache(maxsize=None)
def scalars():
    with open(os.path.join(GOLD, "scalars.json")) as fh:
        return json.load(fh)


def field_for(K):
    return fields()[{2: "velocity_2", 4: "velocity_2", 6: "velocity_2", 10: "velocity_10",
                     100: "velocity_100", 400: "velocity_100", 10000: "velocity_100"}[K]]


def traj(K):
    t = np.load(os.path.join(GOLD, f"traj_{K}_buoys.npz"))
    return t["x_0_array"], t["u_d_array"]


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def q_nodal(V):
    """reference_runs/u_bar_chapter_6.3.3/q_backup/q.h5 (collapsed P2^2 numbering) as a nodal field."""
    F = fields()
    qv, qcd = F["q_vector"], F["q_cell_dofs"].reshape(-1, 12)
    q = np.zeros((V.num_nodes, 2))
    q[V.cell_nodes, 0] = qv[qcd[:, :6]]
    q[V.cell_nodes, 1] = qv[qcd[:, 6:]]
    return q


def recover_control(V, O, w):
    """the control whose state is w, from the Gamma_1 rows of the residual."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    nn = V.num_nodes
    R = O.forward_residual(w, np.zeros((nn, 2)))
    g1 = np.unique(V.g1_nodes)
    loc = np.array([[4, -1, 2], [-1, 4, 2], [2, 2, 16]]) / 30.0
    r = np.repeat(V.g1_nodes, 3, axis=1).ravel()
    c = np.tile(V.g1_nodes, (1, 3)).ravel()
    M = sp.coo_matrix(((V.g1_len[:, None, None] * loc[None]).ravel(), (r, c)), shape=(nn, nn)).tocsr()
    Mg = M[g1][:, g1].tocsc()
    f = np.zeros((nn, 2))
    f[g1, 0] = spla.spsolve(Mg, R[V.dof_ux[g1]])
    f[g1, 1] = spla.spsolve(Mg, R[V.dof_uy[g1]])
    return f, R


class OraclePipeline:
    """One gradient evaluation / GD loop assembled from the oracle pieces only (test infrastructure)."""

    def __init__(self, V, nu, x0, ud, alpha, center=CENTER, h=H, nt=NT):
        self.V, self.O, self.B = V, FEOracle(V, nu), BuoyOracle(V)
        self.x0, self.ud, self.alpha, self.center, self.h, self.nt = x0, ud, alpha, np.asarray(center, float), h, nt

    def cost(self, u, f):
        return self.O.cost(u, self.ud, f, self.h, self.alpha)

    def primal(self, w):
        return self.B.forward(self.V.velocity_nodal(w), self.x0, self.nt, self.h, self.center)

    def gradient_step(self, f):
        V, O, B = self.V, self.O, self.B
        w, its, hist = O.newton_solve(f, return_history=True)
        g = O.project_gradient(w)
        x, u, cell, mask, parked = self.primal(w)
        mu = B.adjoint(g, x, u, self.ud, mask, self.h)
        bn = B.point_sources(V.velocity_nodal(w), x, self.ud, mu, mask, self.h, self.center)
        z = O.adjoint_solve(w, O.rhs_from_nodal(bn))
        grad = self.alpha * f - V.velocity_nodal(z)
        return dict(w=w, its=its, hist=hist, g=g, x=x, u=u, mask=mask, parked=parked, mu=mu, bnode=bn, z=z,
                    grad=grad, cell=cell)

    def run(self, f0, num_steps, use_line_search, LR=5.0, tau=0.5, c=1e-4, LR_MIN=1e-6):
        f = f0.copy()
        J_array, inner = [], []
        for i in range(num_steps):
            s = self.gradient_step(f)
            n_in = 0
            if use_line_search:
                gradj = -self.O.boundary_inner(s["grad"], s["grad"])
                cond = -c * gradj
                J_old = self.cost(s["u"], f)
                while True:
                    n_in += 1
                    f_ls = f - LR * s["grad"]
                    w_ls = self.O.newton_solve(f_ls)
                    _, u_ls, _, _, _ = self.primal(w_ls)
                    if J_old - self.cost(u_ls, f_ls) >= LR * cond or LR <= LR_MIN:
                        break
                    LR = max(tau * LR, LR_MIN)
            f = f - LR * s["grad"]
            J_array.append(self.cost(s["u"], f))
            inner.append(n_in)
        return dict(J_array=J_array, inner=inner, f=f, LR=LR)
