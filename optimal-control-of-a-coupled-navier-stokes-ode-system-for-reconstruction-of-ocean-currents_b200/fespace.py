"""Taylor-Hood ``[P2]^2 x P1`` space tables (dofmap, CSR pattern, boundary tables, point-location bins).

Reference: ``W = FunctionSpace(mesh, MixedElement([VectorElement('CG', triangle, 2),
FiniteElement('CG', triangle, 1)]))`` and ``V_vec = TensorFunctionSpace(mesh, "Lagrange", 1)``
(OCP_dolfin.py:107-113).

Conventions (validated against the reference's checkpoints):
cell vertices sorted by global index ``v0<v1<v2``; scalar P2 local order
``[v0, v1, v2, e0=mid(v1v2), e1=mid(v0v2), e2=mid(v0v1)]``; cell dofs
``[u_x(6), u_y(6), p(3)]``.

"Nodes" are the P2 nodes: vertex ``v`` is node ``v``, edge ``e`` is node ``nv+e``.
The kernels work on node-indexed tables (velocity as ``double2`` per node); the
``W`` numbering (dolfin's for the 32x32 square, a node-interleaved one otherwise)
only enters through ``dof_ux/dof_uy/dof_p``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from .mesh import Mesh, mark_boundaries, BoundaryMarking

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

LOCATE_TOL = 1.0e-14      # closed-containment tolerance on barycentric coordinates
LOCATE_MARGIN = 1.0e-9    # "strictly inside the previous cell" fast-path margin


@dataclass
class TaylorHood:
    mesh: Mesh
    numbering: str = "auto"       # "auto" | "dolfin32" | "interleaved"
    # filled by __post_init__ ----------------------------------------------
    marking: BoundaryMarking = field(default=None, repr=False)
    node_coords: np.ndarray = field(default=None, repr=False)   # (nn, 2)
    cell_nodes: np.ndarray = field(default=None, repr=False)    # (nc, 6) i4
    dof_ux: np.ndarray = field(default=None, repr=False)        # (nn,) i4
    dof_uy: np.ndarray = field(default=None, repr=False)        # (nn,) i4
    dof_p: np.ndarray = field(default=None, repr=False)         # (nv,) i4
    cell_dofs: np.ndarray = field(default=None, repr=False)     # (nc, 15) i4
    ndofs: int = 0
    csr_rowptr: np.ndarray = field(default=None, repr=False)    # (ndofs+1,) i4
    csr_col: np.ndarray = field(default=None, repr=False)       # (nnz,) i4
    dirichlet_dofs: np.ndarray = field(default=None, repr=False)
    # Gamma_1 facet tables
    g1_cell: np.ndarray = field(default=None, repr=False)       # (n1,) i4
    g1_local: np.ndarray = field(default=None, repr=False)      # (n1,) i4 local edge index in that cell
    g1_nodes: np.ndarray = field(default=None, repr=False)      # (n1, 3) i4 [va, vb, mid] node ids
    g1_len: np.ndarray = field(default=None, repr=False)        # (n1,) f8
    g1_normal: np.ndarray = field(default=None, repr=False)     # (n1, 2) f8 outward unit normal
    # affine maps  lambda_1 = a1 (x-x0) + b1 (y-y0), lambda_2 = a2 (x-x0) + b2 (y-y0)
    cell_geom: np.ndarray = field(default=None, repr=False)     # (nc, 6) f8 [x0, y0, a1, b1, a2, b2]
    cell_nbr: np.ndarray = field(default=None, repr=False)      # (nc, 3) i4 cell across the edge opposite local vertex i, -1 on the boundary
    # point-location bins
    bin_origin: np.ndarray = field(default=None, repr=False)    # (2,)
    bin_inv_h: np.ndarray = field(default=None, repr=False)     # (2,)
    bin_dims: np.ndarray = field(default=None, repr=False)      # (2,) i4 (nbx, nby)
    bin_ptr: np.ndarray = field(default=None, repr=False)       # (nbx*nby+1,) i4
    bin_cells: np.ndarray = field(default=None, repr=False)     # candidates, ascending per bin

    def __post_init__(self):
        m = self.mesh
        nv, ne, nc = m.num_vertices, m.num_edges, m.num_cells
        self.marking = mark_boundaries(m)
        mid = 0.5 * (m.coords[m.edges[:, 0]] + m.coords[m.edges[:, 1]])
        self.node_coords = np.vstack([m.coords, mid])
        self.cell_nodes = np.hstack([m.cells, m.cell_edges + nv]).astype(np.int32)
        self._number_dofs()
        cn = self.cell_nodes
        self.cell_dofs = np.hstack([self.dof_ux[cn], self.dof_uy[cn], self.dof_p[m.cells]]).astype(np.int32)
        self.ndofs = 2 * (nv + ne) + nv
        self._build_csr()
        self._build_boundary_tables()
        self._build_geometry()
        self._build_bins()

    # -- numbering ----------------------------------------------------------
    @property
    def num_nodes(self) -> int:
        return self.node_coords.shape[0]

    def _number_dofs(self):
        m = self.mesh
        nv, nn = m.num_vertices, m.num_vertices + m.num_edges
        mode = self.numbering
        if mode == "auto":
            mode = "dolfin32" if _is_reference_square32(m) else "interleaved"
        if mode == "dolfin32":
            if not _is_reference_square32(m):
                raise ValueError("dolfin numbering is only known for the reference 32x32 square")
            tab = np.load(os.path.join(_DATA, "dolfin_square32_dofmap.npz"))
            if not np.array_equal(tab["topology"], m.cells):
                raise ValueError("mesh topology differs from the stored dolfin mesh")
            ucd = tab["u_cell_dofs"].reshape(-1, 12).astype(np.int64)
            pcd = tab["p_cell_dofs"].reshape(-1, 3).astype(np.int64)
            ux = -np.ones(nn, np.int64)
            uy = -np.ones(nn, np.int64)
            pp = -np.ones(nv, np.int64)
            ux[self.cell_nodes] = ucd[:, :6]
            uy[self.cell_nodes] = ucd[:, 6:]
            pp[m.cells] = pcd
            # every cell must agree on the dof of a shared node (checks the local ordering convention)
            assert np.array_equal(ux[self.cell_nodes], ucd[:, :6])
            assert np.array_equal(uy[self.cell_nodes], ucd[:, 6:])
            assert np.array_equal(pp[m.cells], pcd)
            self.dof_ux, self.dof_uy, self.dof_p = (a.astype(np.int32) for a in (ux, uy, pp))
        elif mode == "interleaved":
            n = np.arange(nn, dtype=np.int32)
            self.dof_ux, self.dof_uy = 2 * n, 2 * n + 1
            self.dof_p = (2 * nn + np.arange(nv)).astype(np.int32)
        else:
            raise ValueError(f"unknown numbering {self.numbering!r}")
        self.numbering = mode

    # -- sparsity -------------------------------------------------------------
    def _build_csr(self):
        """CSR pattern = union over cells of dofs x dofs (equal to dolfin's count on the reference mesh).  Built from the
        unique NODE pairs that share a cell (36 + 18 + 9 per cell instead of 225 dof pairs: 512 x 512 cells would
        otherwise mean sorting 1.2e8 keys), expanded to dof pairs through the (injective) dof maps."""
        cn = self.cell_nodes.astype(np.int64)
        nn, n = self.num_nodes, self.ndofs

        def unique_pairs(a, b):
            """unique (a_i, b_j) over the cells; a (nc, p), b (nc, q) node ids"""
            key = np.sort((a[:, :, None] * nn + b[:, None, :]).reshape(-1))      # (sort + compare: numpy's hash-based
            key = key[np.r_[True, key[1:] != key[:-1]]]                          #  unique is 10x slower on 1e7 keys)
            return key // nn, key % nn

        ux, uy, pp = (d.astype(np.int64) for d in (self.dof_ux, self.dof_uy, self.dof_p))
        a, b = unique_pairs(cn, cn)                      # P2 node - P2 node
        c, v = unique_pairs(cn, cn[:, :3])               # P2 node - vertex
        v1, v2 = unique_pairs(cn[:, :3], cn[:, :3])      # vertex - vertex (the structurally present, zero p-p block)
        rows = np.concatenate([ux[a], ux[a], uy[a], uy[a], ux[c], uy[c], pp[v], pp[v], pp[v1]])
        cols = np.concatenate([ux[b], uy[b], ux[b], uy[b], pp[v], pp[v], ux[c], uy[c], pp[v2]])
        key = np.sort(rows * n + cols)
        self.csr_col = (key % n).astype(np.int32)
        self.csr_rowptr = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(np.bincount(key // n, minlength=n), out=self.csr_rowptr[1:])

    @property
    def cell_slots(self) -> np.ndarray:
        """(nc, 225) i4: CSR position of element entry (i, j) (the library builds its own copy in ocp_create)"""
        cd = self.cell_dofs.astype(np.int64)
        key = (np.repeat(cd, 15, axis=1) * self.ndofs + np.tile(cd, (1, 15))).reshape(-1)
        full = np.repeat(np.arange(self.ndofs, dtype=np.int64), np.diff(self.csr_rowptr)) * self.ndofs + self.csr_col
        return np.searchsorted(full, key).reshape(-1, 225).astype(np.int32)

    # -- boundary -------------------------------------------------------------
    def _build_boundary_tables(self):
        m, nv = self.mesh, self.mesh.num_vertices
        mk = self.marking
        e = mk.dirichlet
        nodes = np.unique(np.r_[m.edges[e, 0], m.edges[e, 1], e + nv])
        self.dirichlet_dofs = np.unique(np.r_[self.dof_ux[nodes], self.dof_uy[nodes]]).astype(np.int32)
        g = mk.gamma1
        cell = m.edge_cells[g, 0]
        loc = np.argmax(m.cell_edges[cell] == g[:, None], axis=1)
        va, vb = m.edges[g, 0], m.edges[g, 1]
        pa, pb = m.coords[va], m.coords[vb]
        t = pb - pa
        length = np.hypot(t[:, 0], t[:, 1])
        nrm = np.stack([t[:, 1], -t[:, 0]], axis=1) / length[:, None]
        opp = m.coords[m.cells[cell, loc]]          # vertex opposite the facet
        flip = np.einsum("ij,ij->i", nrm, opp - pa) > 0
        nrm[flip] *= -1.0
        self.g1_cell = cell.astype(np.int32)
        self.g1_local = loc.astype(np.int32)
        self.g1_nodes = np.stack([va, vb, g + nv], axis=1).astype(np.int32)
        self.g1_len, self.g1_normal = length, nrm

    # -- geometry -------------------------------------------------------------
    def _build_geometry(self):
        p = self.mesh.coords[self.mesh.cells]
        x0, y0 = p[:, 0, 0], p[:, 0, 1]
        j11, j12 = p[:, 1, 0] - x0, p[:, 2, 0] - x0
        j21, j22 = p[:, 1, 1] - y0, p[:, 2, 1] - y0
        det = j11 * j22 - j12 * j21
        self.cell_geom = np.stack([x0, y0, j22 / det, -j12 / det, -j21 / det, j11 / det], axis=1)
        m = self.mesh
        ec = m.edge_cells[m.cell_edges]                               # (nc, 3, 2)
        me = np.arange(m.num_cells, dtype=np.int32)[:, None]
        self.cell_nbr = np.where(ec[:, :, 0] == me, ec[:, :, 1], ec[:, :, 0]).astype(np.int32)

    def _build_bins(self):
        m = self.mesh
        lo, hi = m.coords.min(axis=0), m.coords.max(axis=0)
        p = m.coords[m.cells]
        cmin, cmax = p.min(axis=1), p.max(axis=1)
        hmean = np.sqrt(2.0 * m.cell_areas().mean())
        dims = np.maximum(1, np.floor((hi - lo) / hmean + 0.5)).astype(np.int64)
        inv_h = dims / (hi - lo)
        pad = 1e-9 * (hi - lo).max()
        i0 = np.clip(np.floor((cmin - pad - lo) * inv_h).astype(np.int64), 0, dims - 1)
        i1 = np.clip(np.floor((cmax + pad - lo) * inv_h).astype(np.int64), 0, dims - 1)
        bins, cells = [], []
        span = (i1 - i0 + 1)
        for dx in range(int(span[:, 0].max())):
            for dy in range(int(span[:, 1].max())):
                ok = (dx < span[:, 0]) & (dy < span[:, 1])
                c = np.nonzero(ok)[0]
                bins.append((i0[c, 1] + dy) * dims[0] + (i0[c, 0] + dx))
                cells.append(c)
        bins, cells = np.concatenate(bins), np.concatenate(cells)
        order = np.lexsort((cells, bins))
        bins, cells = bins[order], cells[order]
        nb = int(dims[0] * dims[1])
        self.bin_ptr = np.zeros(nb + 1, dtype=np.int32)
        np.cumsum(np.bincount(bins, minlength=nb), out=self.bin_ptr[1:])
        self.bin_cells = cells.astype(np.int32)
        self.bin_origin, self.bin_inv_h = lo.copy(), inv_h
        self.bin_dims = dims.astype(np.int32)

    # -- helpers ----------------------------------------------------------------
    def velocity_nodal(self, w: np.ndarray) -> np.ndarray:
        """(nn, 2) node-indexed velocity from a W vector."""
        return np.stack([w[self.dof_ux], w[self.dof_uy]], axis=1)

    def interpolate_control(self, fx, fy, degree: int) -> np.ndarray:
        """P2 nodal vector (nn, 2) of a control given as callables of (x, y).

        ``Expression(..., degree=1)`` (OCP_dolfin.py:143-144) is interpolated to P1 per
        cell before quadrature, i.e. the mid-edge value is the mean of the two vertex
        values; ``degree=2`` is nodal P2 interpolation.
        """
        m = self.mesh
        nv = m.num_vertices
        out = np.empty((self.num_nodes, 2))
        x, y = self.node_coords[:, 0], self.node_coords[:, 1]
        out[:, 0], out[:, 1] = fx(x, y), fy(x, y)
        if degree == 1:
            out[nv:] = 0.5 * (out[m.edges[:, 0]] + out[m.edges[:, 1]])
        elif degree != 2:
            raise ValueError("degree must be 1 or 2")
        return out


def _is_reference_square32(m: Mesh) -> bool:
    if m.l_shape or m.num_cells != 2048 or m.num_vertices != 1089:
        return False
    return bool(np.allclose(m.coords[[0, -1]], [[0.0, 0.0], [2.0, 2.0]]))
