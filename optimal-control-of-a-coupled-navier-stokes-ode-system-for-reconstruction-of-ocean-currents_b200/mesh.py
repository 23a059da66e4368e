"""Triangle meshes of the reference's two domains and their topology tables.

Reference: ``RectangleMesh(Point(0,0), Point(2,2), 32, 32)`` (OCP_dolfin.py:99)
and the mshr L-shape ``Rectangle((0,0),(2,1)) + Rectangle((1,1),(2,2))``
(OCP_dolfin.py:82-84).  The square is rebuilt with dolfin's own vertex/cell
numbering ("right" diagonal, SURVEY App. B.2 - checked against the fixtures in
tests/test_mesh.py); mshr/CGAL output is not reproducible, so the L-shape is a
structured (optionally jittered) triangulation of the same domain.

Boundary marking follows OCP_dolfin.py:118-136 (see ``mark_boundaries``).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

DOLFIN_EPS = 3.0e-16  # dolfin/common/constants.h


@dataclass
class Mesh:
    coords: np.ndarray          # (nv, 2) f8
    cells: np.ndarray           # (nc, 3) i4, rows sorted ascending (UFC ordering)
    l_shape: bool = False
    # derived topology -----------------------------------------------------
    edges: np.ndarray = field(default=None, repr=False)        # (ne, 2) i4, v_lo < v_hi
    cell_edges: np.ndarray = field(default=None, repr=False)   # (nc, 3) i4, edge i is opposite local vertex i
    edge_cells: np.ndarray = field(default=None, repr=False)   # (ne, 2) i4, -1 when absent

    def __post_init__(self):
        self.coords = np.ascontiguousarray(self.coords, dtype=np.float64)
        cells = np.sort(np.asarray(self.cells, dtype=np.int64), axis=1)
        self.cells = np.ascontiguousarray(cells.astype(np.int32))
        self._build_edges()

    @property
    def num_vertices(self) -> int:
        return self.coords.shape[0]

    @property
    def num_cells(self) -> int:
        return self.cells.shape[0]

    @property
    def num_edges(self) -> int:
        return self.edges.shape[0]

    def _build_edges(self):
        c = self.cells.astype(np.int64)
        nv = self.num_vertices
        # local edge i is opposite local vertex i: e0=(v1,v2), e1=(v0,v2), e2=(v0,v1)
        lo = np.stack([c[:, 1], c[:, 0], c[:, 0]], axis=1)
        hi = np.stack([c[:, 2], c[:, 2], c[:, 1]], axis=1)
        key = (lo * nv + hi).reshape(-1)
        uniq, inv = np.unique(key, return_inverse=True)
        self.edges = np.stack([uniq // nv, uniq % nv], axis=1).astype(np.int32)
        self.cell_edges = inv.reshape(-1, 3).astype(np.int32)
        ne = uniq.shape[0]
        edge_cells = np.full((ne, 2), -1, dtype=np.int32)
        cell_of = np.repeat(np.arange(self.num_cells, dtype=np.int32), 3)
        order = np.argsort(inv, kind="stable")
        inv_s, cell_s = inv[order], cell_of[order]
        first = np.r_[True, inv_s[1:] != inv_s[:-1]]
        edge_cells[inv_s[first], 0] = cell_s[first]
        edge_cells[inv_s[~first], 1] = cell_s[~first]
        self.edge_cells = edge_cells

    def cell_areas(self) -> np.ndarray:
        p = self.coords[self.cells]
        d1, d2 = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
        return 0.5 * np.abs(d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0])


def rectangle_mesh(x0: float, y0: float, x1: float, y1: float, nx: int, ny: int) -> Mesh:
    """dolfin ``RectangleMesh(..., nx, ny)`` with the default "right" diagonal.

    Vertex (i, j) -> i + (nx+1) j; cell 2(i + nx j) = (v, v+1, v+nx+2),
    cell 2(i + nx j)+1 = (v, v+nx+1, v+nx+2).
    """
    xs = x0 + (x1 - x0) * np.arange(nx + 1) / nx
    ys = y0 + (y1 - y0) * np.arange(ny + 1) / ny
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    coords = np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v = (i + (nx + 1) * j).reshape(-1)
    lower = np.stack([v, v + 1, v + nx + 2], axis=1)
    upper = np.stack([v, v + nx + 1, v + nx + 2], axis=1)
    cells = np.empty((2 * nx * ny, 3), dtype=np.int64)
    cells[0::2], cells[1::2] = lower, upper
    return Mesh(coords, cells, l_shape=False)


def square_mesh(n: int = 32) -> Mesh:
    """The reference's square ``[0,2]^2`` at ``unit_square_resolution = n`` (OCP_dolfin.py:30, 99)."""
    return rectangle_mesh(0.0, 0.0, 2.0, 2.0, n, n)


def lshape_mesh(m: int = 20, jitter: float = 0.0, seed: int = 0) -> Mesh:
    """L-shape ``[0,2]x[0,1] U [1,2]x[1,2]`` (OCP_dolfin.py:82-84) with spacing 1/m.

    ``m=20`` gives 2400 cells, the size mshr produces at ``L_shape_resolution=50``.
    ``jitter`` (fraction of the spacing) perturbs interior vertices so that the
    generic point-location path is exercised on a non-structured mesh.
    """
    h = 1.0 / m
    ids = -np.ones((2 * m + 1, 2 * m + 1), dtype=np.int64)  # [j, i]
    coords = []
    for j in range(2 * m + 1):
        for i in range(2 * m + 1):
            if j <= m or i >= m:
                ids[j, i] = len(coords)
                coords.append((i * h, j * h))
    coords = np.array(coords)
    cells = []
    for j in range(2 * m):
        for i in range(2 * m):
            if j < m or i >= m:
                a, b, c, d = ids[j, i], ids[j, i + 1], ids[j + 1, i], ids[j + 1, i + 1]
                cells.append((a, b, d))
                cells.append((a, c, d))
    cells = np.array(cells)
    mesh = Mesh(coords, cells, l_shape=True)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        on_bnd = np.zeros(mesh.num_vertices, dtype=bool)
        bnd_edges = mesh.edges[mesh.edge_cells[:, 1] < 0]
        on_bnd[bnd_edges.reshape(-1)] = True
        d = rng.uniform(-jitter * h, jitter * h, size=coords.shape)
        d[on_bnd] = 0.0
        mesh = Mesh(coords + d, cells, l_shape=True)
    return mesh


@dataclass
class BoundaryMarking:
    """Boundary facets split as in OCP_dolfin.py:118-136."""
    facets: np.ndarray        # (nb,) edge ids of all boundary facets
    gamma1: np.ndarray        # (n1,) edge ids with marker 1 ("Neumann", the control boundary)
    dirichlet: np.ndarray     # (nd,) edge ids on which DirichletBC(W.sub(0), (0,0), boundary) acts


def mark_boundaries(mesh: Mesh) -> BoundaryMarking:
    """Facet classification.

    ``Neumann.inside`` (OCP_dolfin.py:118-121) and ``boundary`` (OCP_dolfin.py:131-133)
    are evaluated by dolfin at both facet vertices *and* the facet midpoint;
    a facet is marked only if all three pass.  With ``s = x`` (square) or
    ``s = y`` (L-shape):  Gamma_1 = {|x| < eps or |2 - s| < eps},
    Dirichlet = {x > eps and |2 - s| > eps}.  The corner facets of the square's
    bottom/top edges that touch x=0 or x=2 therefore belong to neither set.
    """
    bnd = np.nonzero(mesh.edge_cells[:, 1] < 0)[0].astype(np.int32)
    ev = mesh.edges[bnd]
    pa, pb = mesh.coords[ev[:, 0]], mesh.coords[ev[:, 1]]
    pm = 0.5 * (pa + pb)
    sidx = 1 if mesh.l_shape else 0

    def neumann(p):
        return (np.abs(p[:, 0]) < DOLFIN_EPS) | (np.abs(2.0 - p[:, sidx]) < DOLFIN_EPS)

    def dirich(p):
        return (p[:, 0] > DOLFIN_EPS) & (np.abs(2.0 - p[:, sidx]) > DOLFIN_EPS)

    g1 = neumann(pa) & neumann(pb) & neumann(pm)
    dr = dirich(pa) & dirich(pb) & dirich(pm)
    return BoundaryMarking(facets=bnd, gamma1=bnd[g1], dirichlet=bnd[dr])
