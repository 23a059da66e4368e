"""Mesh / Taylor-Hood tables against the numbers the reference's fixtures pin."""
import os

import numpy as np
import pytest

import helpers as H
from ocp_b200 import h5lite
from ocp_b200.fespace import TaylorHood
from ocp_b200.mesh import lshape_mesh, mark_boundaries, square_mesh


def test_square32_sizes_and_dolfin_numbering():
    V = H.square32()
    m = V.mesh
    assert (m.num_cells, m.num_vertices, m.num_edges) == (2048, 1089, 3136)
    assert V.numbering == "dolfin32"
    assert V.ndofs == 9539 and V.num_nodes == 4225
    assert V.csr_col.size == 276937
    assert int(np.diff(V.csr_rowptr).max()) == 45
    rows = np.repeat(np.arange(V.ndofs), np.diff(V.csr_rowptr))
    assert int(np.abs(rows - V.csr_col).max()) == 305            # bandwidth in dolfin numbering
    assert V.dirichlet_dofs.size == 244 and V.g1_cell.size == 64
    assert (V.dof_ux[0], V.dof_uy[0], V.dof_p[0]) == (4560, 4563, 4566)
    # a permutation of 0..ndofs-1
    alld = np.sort(np.r_[V.dof_ux, V.dof_uy, V.dof_p])
    assert np.array_equal(alld, np.arange(V.ndofs))


def test_square32_mesh_matches_stored_dolfin_mesh():
    V = H.square32()
    tab = np.load(os.path.join(os.path.dirname(h5lite.__file__), "data", "dolfin_square32_dofmap.npz"))
    assert np.array_equal(tab["topology"], V.mesh.cells)
    assert np.array_equal(tab["geometry"], V.mesh.coords)


def test_corner_facet_rule():
    """The four bottom/top facets touching x=0 / x=2 are neither Dirichlet nor Gamma_1 (App. A.2)."""
    m = square_mesh(32)
    mk = mark_boundaries(m)
    assert mk.facets.size == 128 and mk.gamma1.size == 64 and mk.dirichlet.size == 60
    V = H.square32()
    # corner vertices are free
    corners = [0, 32, 33 * 32, 33 * 33 - 1]
    assert not np.isin(V.dof_ux[corners], V.dirichlet_dofs).any()


def test_gamma1_normals_and_lengths():
    V = H.square32()
    x = V.node_coords[V.g1_nodes[:, 2], 0]
    assert np.allclose(V.g1_len, 2.0 / 32)
    assert np.allclose(V.g1_normal[x < 1, 0], -1.0) and np.allclose(V.g1_normal[x > 1, 0], 1.0)
    assert np.allclose(V.g1_normal[:, 1], 0.0)
    assert np.isclose(V.g1_len.sum(), 4.0)


def test_lshape_domain_and_marking():
    m = lshape_mesh(20)
    assert np.isclose(m.cell_areas().sum(), 3.0)
    V = TaylorHood(m)
    assert V.numbering == "interleaved"
    # Gamma_1 = {x = 0, y in [0,1]} U {y = 2, x in [1,2]} (OCP_dolfin.py:120-121 with point = x[1])
    assert np.isclose(V.g1_len.sum(), 2.0)
    mid = V.node_coords[V.g1_nodes[:, 2]]
    assert np.all((np.abs(mid[:, 0]) < 1e-12) | (np.abs(mid[:, 1] - 2.0) < 1e-12))


def test_cell_geometry_is_inverse_affine_map():
    V = H.lshape()
    rng = np.random.default_rng(0)
    lam = rng.dirichlet(np.ones(3), size=V.mesh.num_cells)
    p = np.einsum("ci,cid->cd", lam, V.mesh.coords[V.mesh.cells])
    g = V.cell_geom
    l1 = g[:, 2] * (p[:, 0] - g[:, 0]) + g[:, 3] * (p[:, 1] - g[:, 1])
    l2 = g[:, 4] * (p[:, 0] - g[:, 0]) + g[:, 5] * (p[:, 1] - g[:, 1])
    assert np.allclose(l1, lam[:, 1], atol=1e-13) and np.allclose(l2, lam[:, 2], atol=1e-13)


def test_bins_cover_every_cell_vertex_and_centroid():
    for V in (H.square32(), H.lshape()):
        m = V.mesh
        pts = np.vstack([m.coords[m.cells].mean(axis=1), m.coords[m.cells[:, 0]]])
        owner = np.r_[np.arange(m.num_cells), np.arange(m.num_cells)]
        ij = np.floor((pts - V.bin_origin) * V.bin_inv_h).astype(int)
        ij = np.clip(ij, 0, V.bin_dims - 1)
        b = ij[:, 1] * V.bin_dims[0] + ij[:, 0]
        for k in range(pts.shape[0]):
            cand = V.bin_cells[V.bin_ptr[b[k]]:V.bin_ptr[b[k] + 1]]
            assert owner[k] in cand
            assert np.all(np.diff(cand) > 0)


@pytest.mark.skipif(not os.path.isdir("/root/reference/reference_runs"), reason="reference mount absent")
def test_h5lite_reads_reference_checkpoints():
    ds = h5lite.read_checkpoint("/root/reference/reference_runs/u_bar_chapter_6.3.3/paraview/checkpoint/u.h5", "u")
    assert ds["topology"].shape == (2048, 3) and ds["geometry"].shape == (1089, 2)
    assert ds["cell_dofs"].size == 24576 and ds["vector"].size == 9539
    assert np.array_equal(ds["x_cell_dofs"].ravel(), 12 * np.arange(2049))
    assert np.array_equal(ds["vector"], H.fields()["u_bar"])
    q = h5lite.read_checkpoint("/root/reference/reference_runs/u_bar_chapter_6.3.3/q_backup/q.h5", "f")
    assert q["vector"].size == 8450


@pytest.mark.parametrize("which", ["square8", "square32", "lshape"])
def test_csr_pattern_is_the_union_of_element_patterns(which):
    """The pattern is built from unique NODE pairs (fespace._build_csr); it must equal the brute-force union over the
    cells of dofs x dofs - also in dolfin's numbering of the reference mesh - and cell_slots must address it."""
    V = {"square8": lambda: TaylorHood(square_mesh(8)), "square32": H.square32,
         "lshape": lambda: TaylorHood(lshape_mesh(6, jitter=0.2))}[which]()
    cd = V.cell_dofs.astype(np.int64)
    n = V.ndofs
    key = (np.repeat(cd, 15, axis=1) * n + np.tile(cd, (1, 15))).reshape(-1)
    uniq = np.unique(key)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(V.csr_rowptr))
    assert np.array_equal(rows * n + V.csr_col, uniq)
    assert V.csr_rowptr[0] == 0 and V.csr_rowptr[-1] == uniq.size
    sl = V.cell_slots
    assert sl.shape == (V.mesh.num_cells, 225)
    assert np.array_equal(uniq[sl.reshape(-1)], key)


def test_ensemble_needs_a_gpu():
    """ocp_b200.ensemble (cfg4) has no CPU path either."""
    import torch
    from ocp_b200 import capi
    from ocp_b200.ensemble import Case, Ensemble
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    V = TaylorHood(square_mesh(4))
    c = Case("x", np.array([[0.5, 0.5]]), np.zeros((1, 200, 2)), np.zeros((V.num_nodes, 2)))
    with pytest.raises(capi.OcpError):
        Ensemble(V, [c])
