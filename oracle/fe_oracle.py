"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

NumPy/SciPy restatement of the finite-element half of the reference's hot path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module; the product path never does.

What is restated (file:line of the reference) and which third-party routine did the
arithmetic there (legacy FEniCS 2019.x: dolfin + FFC + UFL, un-vendored and un-pinned,
SURVEY 8(c)):

  forward_residual / forward_jacobian   the UFL forms at OCP_dolfin.py:321-323 (FFC quadrature kernels,
                                        dolfin ``assemble``); boundary term on ds(1) only
  newton_solve                          ``solve(F == 0, w, bcs)`` OCP_dolfin.py:325 (dolfin NewtonSolver defaults:
                                        atol 1e-10, rtol 1e-9, full step, zero initial guess, LU)
  project_gradient                      ``project(grad(w.sub(0)), V_vec)`` OCP_dolfin.py:328-329
  adjoint_matrix / adjoint_solve        ``assemble(aAdj)``, ``bcs[0].apply``, ``solve(A, z, b)`` OCP_dolfin.py:344-371
  cost / boundary_inner / gradj         ``J`` OCP_dolfin.py:258-261, ``assemble(inner(alpha*f - zSol, df)*ds(1))``
                                        OCP_dolfin.py:379, 388

All integrands are polynomials on affine cells, so any exact quadrature reproduces
FFC's result to round-off; this oracle uses a degree-6 (12-point) triangle rule and a
5-point Gauss-Legendre rule on facets, evaluated through the *full* cell basis on the
facet (as FFC does) - deliberately different from the CUDA kernels (7-point rule,
facet trace basis) so that agreement checks exactness of both.

Pinned by tests/test_oracle_golden.py against the reference's FEniCS artefacts
(KAT K2-K5 of SURVEY section 4).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# Dunavant degree-6 rule, 12 points (weights sum to 1; multiply by the cell area)
_A1, _W1 = 0.063089014491502, 0.050844906370207
_A2, _W2 = 0.249286745170910, 0.116786275726379
_A3, _B3, _W3 = 0.053145049844817, 0.310352451033785, 0.082851075618374


def _tri_rule():
    pts, wts = [], []
    for a, w in ((_A1, _W1), (_A2, _W2)):
        b = 1.0 - 2.0 * a
        pts += [(b, a, a), (a, b, a), (a, a, b)]
        wts += [w] * 3
    a, b = _A3, _B3
    c = 1.0 - a - b
    pts += [(a, b, c), (b, a, c), (a, c, b), (c, a, b), (b, c, a), (c, b, a)]
    wts += [_W3] * 6
    return np.array(pts), np.array(wts)


_TRI_L, _TRI_W = _tri_rule()
_GL_X, _GL_W = np.polynomial.legendre.leggauss(5)
_GL_S, _GL_W = 0.5 * (_GL_X + 1.0), 0.5 * _GL_W


def p2_basis(l):
    """l: (..., 3) barycentrics -> (..., 6) values, (..., 6, 3) d/d lambda_i"""
    l0, l1, l2 = l[..., 0], l[..., 1], l[..., 2]
    z = np.zeros_like(l0)
    phi = np.stack([l0 * (2 * l0 - 1), l1 * (2 * l1 - 1), l2 * (2 * l2 - 1), 4 * l1 * l2, 4 * l0 * l2, 4 * l0 * l1], -1)
    d = np.stack([
        np.stack([4 * l0 - 1, z, z], -1),
        np.stack([z, 4 * l1 - 1, z], -1),
        np.stack([z, z, 4 * l2 - 1], -1),
        np.stack([z, 4 * l2, 4 * l1], -1),
        np.stack([4 * l2, z, 4 * l0], -1),
        np.stack([4 * l1, 4 * l0, z], -1),
    ], -2)
    return phi, d


class FEOracle:
    """All FE operators of the path on one space ``V`` (ocp_b200.fespace.TaylorHood)."""

    def __init__(self, V, viscosity: float):
        self.V = V
        self.nu = float(viscosity)
        m = V.mesh
        p = m.coords[m.cells]                                  # (nc,3,2)
        d1, d2 = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
        det = d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0]
        self.area = 0.5 * np.abs(det)
        gl1 = np.stack([d2[:, 1], -d2[:, 0]], 1) / det[:, None]
        gl2 = np.stack([-d1[:, 1], d1[:, 0]], 1) / det[:, None]
        self.gradl = np.stack([-gl1 - gl2, gl1, gl2], 1)        # (nc,3,2)
        phi, dphi = p2_basis(_TRI_L)                            # (nq,6), (nq,6,3)
        self.phi = phi
        self.gphi = np.einsum("qai,cid->cqad", dphi, self.gradl)  # (nc,nq,6,2)
        self.psi = _TRI_L                                       # (nq,3)
        self.wq = _TRI_W[None, :] * self.area[:, None]          # (nc,nq)
        cd = V.cell_dofs.astype(np.int64)
        self._rows = np.repeat(cd, 15, axis=1).reshape(-1)
        self._cols = np.tile(cd, (1, 15)).reshape(-1)
        # facet tables: barycentrics of the facet quadrature points inside the owning cell
        n1 = V.g1_cell.shape[0]
        lam = np.zeros((n1, _GL_S.size, 3))
        cells = m.cells[V.g1_cell]
        for f in range(n1):
            loc = {int(v): i for i, v in enumerate(cells[f])}
            ia, ib = loc[int(V.g1_nodes[f, 0])], loc[int(V.g1_nodes[f, 1])]
            lam[f, :, ia] = 1.0 - _GL_S
            lam[f, :, ib] = _GL_S
        self.f_phi, _ = p2_basis(lam)                           # (n1,nq,6)
        self.f_w = _GL_W[None, :] * V.g1_len[:, None]           # (n1,nq)
        self._p1_mass = None
        self._p1_lu = None

    # ------------------------------------------------------------------ cells
    def _cell_fields(self, w):
        cd = self.V.cell_dofs
        U, Vv, P = w[cd[:, 0:6]], w[cd[:, 6:12]], w[cd[:, 12:15]]
        ux = U @ self.phi.T                                    # (nc,nq)
        uy = Vv @ self.phi.T
        gux = np.einsum("ca,cqad->cqd", U, self.gphi)          # (nc,nq,2)
        guy = np.einsum("ca,cqad->cqd", Vv, self.gphi)
        pr = P @ self.psi.T
        return ux, uy, gux, guy, pr

    def cell_residual(self, w, nu=None):
        nu = self.nu if nu is None else nu
        ux, uy, gux, guy, pr = self._cell_fields(w)
        phi, g, wq = self.phi, self.gphi, self.wq
        cx = ux * gux[..., 0] + uy * gux[..., 1]
        cy = ux * guy[..., 0] + uy * guy[..., 1]
        Rx = np.einsum("cq,cqd,cqad->ca", wq, nu * gux, g) + np.einsum("cq,cq,qa->ca", wq, cx, phi) \
            + np.einsum("cq,cq,cqa->ca", wq, pr, g[..., 0])
        Ry = np.einsum("cq,cqd,cqad->ca", wq, nu * guy, g) + np.einsum("cq,cq,qa->ca", wq, cy, phi) \
            + np.einsum("cq,cq,cqa->ca", wq, pr, g[..., 1])
        div = gux[..., 0] + guy[..., 1]
        Rp = np.einsum("cq,cq,qi->ci", wq, div, self.psi)
        return np.concatenate([Rx, Ry, Rp], axis=1)            # (nc,15)

    def cell_jacobian(self, w, nu=None):
        nu = self.nu if nu is None else nu
        ux, uy, gux, guy, _ = self._cell_fields(w)
        phi, g, wq, psi = self.phi, self.gphi, self.wq, self.psi
        nc = wq.shape[0]
        A = np.zeros((nc, 15, 15))
        stiff = np.einsum("cq,cqad,cqbd->cab", wq, g, g)
        adv = np.einsum("cq,cqb,qa->cab", wq, ux[..., None] * g[..., 0] + uy[..., None] * g[..., 1], phi)
        mass = lambda coef: np.einsum("cq,cq,qa,qb->cab", wq, coef, phi, phi)
        A[:, 0:6, 0:6] = nu * stiff + adv + mass(gux[..., 0])
        A[:, 0:6, 6:12] = mass(gux[..., 1])
        A[:, 6:12, 0:6] = mass(guy[..., 0])
        A[:, 6:12, 6:12] = nu * stiff + adv + mass(guy[..., 1])
        bx = np.einsum("cq,cqa,qj->caj", wq, g[..., 0], psi)
        by = np.einsum("cq,cqa,qj->caj", wq, g[..., 1], psi)
        A[:, 0:6, 12:15], A[:, 6:12, 12:15] = bx, by
        A[:, 12:15, 0:6], A[:, 12:15, 6:12] = bx.transpose(0, 2, 1), by.transpose(0, 2, 1)
        return A

    # ----------------------------------------------------------------- facets
    def _facet_dofs(self):
        cd = self.V.cell_dofs[self.V.g1_cell]
        return cd[:, 0:6], cd[:, 6:12]

    def facet_residual(self, w, f_nodal):
        """-1/2 (u.n)(u.v) ds(1) - f.v ds(1); returns (rows, values)."""
        V = self.V
        dx_, dy_ = self._facet_dofs()
        cn = V.cell_nodes[V.g1_cell]
        ux = np.einsum("fa,fqa->fq", w[dx_], self.f_phi)
        uy = np.einsum("fa,fqa->fq", w[dy_], self.f_phi)
        fx = np.einsum("fa,fqa->fq", f_nodal[cn, 0], self.f_phi)
        fy = np.einsum("fa,fqa->fq", f_nodal[cn, 1], self.f_phi)
        un = ux * V.g1_normal[:, 0:1] + uy * V.g1_normal[:, 1:2]
        Rx = np.einsum("fq,fq,fqa->fa", self.f_w, -0.5 * un * ux - fx, self.f_phi)
        Ry = np.einsum("fq,fq,fqa->fa", self.f_w, -0.5 * un * uy - fy, self.f_phi)
        return np.concatenate([dx_, dy_], 1), np.concatenate([Rx, Ry], 1)

    def facet_jacobian(self, w):
        V = self.V
        dx_, dy_ = self._facet_dofs()
        ux = np.einsum("fa,fqa->fq", w[dx_], self.f_phi)
        uy = np.einsum("fa,fqa->fq", w[dy_], self.f_phi)
        nx, ny = V.g1_normal[:, 0:1], V.g1_normal[:, 1:2]
        un = ux * nx + uy * ny
        mm = lambda coef: np.einsum("fq,fq,fqa,fqb->fab", self.f_w, coef, self.f_phi, self.f_phi)
        n1 = dx_.shape[0]
        B = np.zeros((n1, 12, 12))
        B[:, 0:6, 0:6] = -0.5 * mm(nx * ux + un)
        B[:, 6:12, 0:6] = -0.5 * mm(nx * uy)
        B[:, 0:6, 6:12] = -0.5 * mm(ny * ux)
        B[:, 6:12, 6:12] = -0.5 * mm(ny * uy + un)
        return np.concatenate([dx_, dy_], 1), B

    # ---------------------------------------------------------------- globals
    def forward_residual(self, w, f_nodal):
        n = self.V.ndofs
        R = np.zeros(n)
        np.add.at(R, self.V.cell_dofs.reshape(-1), self.cell_residual(w).reshape(-1))
        rows, vals = self.facet_residual(w, f_nodal)
        np.add.at(R, rows.reshape(-1), vals.reshape(-1))
        return R

    def jacobian_unconstrained(self, w, nu=None):
        n = self.V.ndofs
        A = self.cell_jacobian(w, nu)
        J = sp.coo_matrix((A.reshape(-1), (self._rows, self._cols)), shape=(n, n)).tocsr()
        rows, B = self.facet_jacobian(w)
        r = np.repeat(rows, 12, axis=1).reshape(-1)
        c = np.tile(rows, (1, 12)).reshape(-1)
        J = J + sp.coo_matrix((B.reshape(-1), (r, c)), shape=(n, n)).tocsr()
        return J.tocsr()

    def on_pattern(self, M):
        """Values of sparse M on the space's CSR pattern (structural zeros included)."""
        V = self.V
        n = V.ndofs
        out = np.zeros(V.csr_col.size)
        Mc = M.tocoo()
        key = Mc.row.astype(np.int64) * n + Mc.col
        rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(V.csr_rowptr))
        pkey = rows * n + V.csr_col
        pos = np.searchsorted(pkey, key)
        assert np.array_equal(pkey[pos], key), "matrix entry outside the space's sparsity pattern"
        np.add.at(out, pos, Mc.data)
        return out

    @staticmethod
    def apply_dirichlet_rows(J, dofs):
        """``bc.apply(A)``: zero the rows, unit diagonal, columns untouched."""
        n = J.shape[0]
        mask = np.zeros(n)
        mask[np.asarray(dofs, dtype=np.int64)] = 1.0
        out = (sp.diags(1.0 - mask) @ J.tocsr() + sp.diags(mask)).tocsr()
        out.eliminate_zeros()
        return out

    def forward_jacobian(self, w):
        return self.apply_dirichlet_rows(self.jacobian_unconstrained(w), self.V.dirichlet_dofs)

    def newton_solve(self, f_nodal, w0=None, atol=1e-10, rtol=1e-9, maxit=50, return_history=False,
                     dir_dofs=None, dir_vals=None):
        """dolfin NewtonSolver defaults.  ``dir_dofs/dir_vals`` replace the space's homogeneous
        velocity BC by general Dirichlet data (the twin experiment of plotting/ud_construction_pipeline.py:95-106)."""
        V = self.V
        w = np.zeros(V.ndofs) if w0 is None else w0.copy()
        d = V.dirichlet_dofs if dir_dofs is None else np.asarray(dir_dofs)
        dv = np.zeros(d.size) if dir_vals is None else np.asarray(dir_vals, float)
        hist = []

        def resid(w):
            R = self.forward_residual(w, f_nodal)
            R[d] = w[d] - dv       # bc.apply(b, x): residual of a Dirichlet row is w_d - g_d
            return R

        R = resid(w)
        r0 = np.linalg.norm(R)
        hist.append(r0)
        it = 0
        while not (hist[-1] < atol or (r0 > 0 and hist[-1] / r0 < rtol)):
            if it >= maxit:
                raise RuntimeError("Newton solver did not converge")
            Jm = self.apply_dirichlet_rows(self.jacobian_unconstrained(w), d)
            dx = spla.splu(Jm.tocsc()).solve(R)
            w = w - dx
            it += 1
            R = resid(w)
            hist.append(np.linalg.norm(R))
        if return_history:
            return w, it, hist
        return w

    # -------------------------------------------------------------- projection
    def p1_mass(self):
        if self._p1_mass is None:
            m = self.V.mesh
            loc = (np.ones((3, 3)) + np.eye(3)) / 12.0
            vals = self.area[:, None, None] * loc[None]
            r = np.repeat(m.cells, 3, axis=1).reshape(-1)
            c = np.tile(m.cells, (1, 3)).reshape(-1)
            nv = m.num_vertices
            self._p1_mass = sp.coo_matrix((vals.reshape(-1), (r, c)), shape=(nv, nv)).tocsc()
            self._p1_lu = spla.splu(self._p1_mass)
        return self._p1_mass

    def project_gradient_rhs(self, w):
        _, _, gux, guy, _ = self._cell_fields(w)
        comps = [gux[..., 0], gux[..., 1], guy[..., 0], guy[..., 1]]     # row-major [[g0,g1],[g2,g3]]
        nv = self.V.mesh.num_vertices
        rhs = np.zeros((nv, 4))
        for j, cpt in enumerate(comps):
            loc = np.einsum("cq,cq,qi->ci", self.wq, cpt, self.psi)
            np.add.at(rhs[:, j], self.V.mesh.cells.reshape(-1), loc.reshape(-1))
        return rhs

    def project_gradient(self, w):
        """(nv,4): vertex values of the L2 projection of grad(u) onto continuous P1, [du_x/dx, du_x/dy, du_y/dx, du_y/dy]."""
        self.p1_mass()
        rhs = self.project_gradient_rhs(w)
        return np.stack([self._p1_lu.solve(rhs[:, j]) for j in range(4)], axis=1)

    # ------------------------------------------------------------------ adjoint
    def adjoint_matrix(self, w):
        """assemble(aAdj) then bcs[0].apply(A): transpose of the nu=1 Jacobian, Dirichlet rows -> identity."""
        At = self.jacobian_unconstrained(w, nu=1.0).T.tocsr()
        return self.apply_dirichlet_rows(At, self.V.dirichlet_dofs)

    def rhs_from_nodal(self, bnode):
        b = np.zeros(self.V.ndofs)
        b[self.V.dof_ux] = bnode[:, 0]
        b[self.V.dof_uy] = bnode[:, 1]
        return b

    def adjoint_solve(self, w, b):
        b = b.copy()
        b[self.V.dirichlet_dofs] = 0.0
        return spla.splu(self.adjoint_matrix(w).tocsc()).solve(b)

    # ---------------------------------------------------------- boundary forms
    def boundary_inner(self, a_nodal, b_nodal):
        """int_{Gamma_1} a . b ds for P2 nodal vector fields (nn,2)."""
        cn = self.V.cell_nodes[self.V.g1_cell]
        tot = 0.0
        for c in range(2):
            av = np.einsum("fa,fqa->fq", a_nodal[cn, c], self.f_phi)
            bv = np.einsum("fa,fqa->fq", b_nodal[cn, c], self.f_phi)
            tot += float(np.sum(self.f_w * av * bv))
        return tot

    def cost(self, u_values, u_d, f_nodal, h, alpha):
        """J, OCP_dolfin.py:258-261 (alpha already multiplied by K, OCP_dolfin.py:76)."""
        partA = 0.5 * np.sum(np.sum(h * (np.linalg.norm(u_values - u_d, axis=2) ** 2), axis=1))
        return float(partA + 0.5 * alpha * self.boundary_inner(f_nodal, f_nodal))

    def divergence_norm(self, w):
        """sqrt(assemble(div(u)*div(u)*dx)), OCP_dolfin.py:430."""
        _, _, gux, guy, _ = self._cell_fields(w)
        div = gux[..., 0] + guy[..., 1]
        return float(np.sqrt(np.sum(self.wq * div * div)))

    def l2_h1_norms(self, w):
        """norm(u,'L2'), norm(u,'H1') of the velocity part (Pipeline_limits.py:433-443)."""
        ux, uy, gux, guy, _ = self._cell_fields(w)
        l2 = np.sum(self.wq * (ux * ux + uy * uy))
        h1s = np.sum(self.wq * (np.sum(gux * gux, -1) + np.sum(guy * guy, -1)))
        return float(np.sqrt(l2)), float(np.sqrt(l2 + h1s))
