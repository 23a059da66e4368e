"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares, refuses to run
without a GPU (no CPU fallback), and its element arithmetic / host analysis agree with the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H
from ocp_b200 import capi
from ocp_b200.build import build
from oracle.fe_oracle import FEOracle

HEADER = os.path.join(H.ROOT, "include", "ocp_b200.h")


@pytest.fixture(scope="module")
def lib():
    build()
    return capi.load_library()


def _header_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ocp_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(lib):
    declared = _header_symbols()
    assert sorted(capi.SYMBOLS) == declared
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (ocp_[a-z0-9_]+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    for s in declared:
        getattr(lib, s)


def test_library_is_sm100a_cuda(lib):
    out = subprocess.check_output(["cuobjdump", "-lelf", capi.LIB_PATH], text=True)
    assert "sm_100a" in out


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.ocp_device_available() == -5
    with pytest.raises(capi.OcpError):
        capi.Context(H.square32(), 1.0, 0.005, 200, (1.0, 1.0))
    from ocp_b200.pipeline import OCP, Parameters
    with pytest.raises(capi.OcpError):
        OCP(H.square32(), Parameters(), np.zeros((1, 2)), np.zeros((1, 200, 2)))


def test_cell_element_arithmetic_matches_oracle(lib):
    V = H.lshape()
    nu = 0.37
    O = FEOracle(V, nu)
    w = np.random.default_rng(1).standard_normal(V.ndofs)
    A, R = O.cell_jacobian(w), O.cell_residual(w)
    for c in range(0, V.mesh.num_cells, 7):
        Ac, Rc = capi.selftest_cell_matrix(V.cell_geom[c], w[V.cell_dofs[c]], nu)
        assert H.rel(Ac, A[c]) < 1e-13 and H.rel(Rc, R[c]) < 1e-13      # two different exact quadrature rules


def test_facet_element_arithmetic_matches_oracle(lib):
    V = H.lshape()
    O = FEOracle(V, 1.0)
    rng = np.random.default_rng(2)
    w, f = rng.standard_normal(V.ndofs), rng.standard_normal((V.num_nodes, 2))
    rows, B = O.facet_jacobian(w)
    _, Rv = O.facet_residual(w, f)
    for k in range(V.g1_cell.size):
        nodes = V.g1_nodes[k]
        uv = np.r_[w[V.dof_ux[nodes]], w[V.dof_uy[nodes]]]
        Af, Rf = capi.selftest_facet_matrix(V.g1_len[k], *V.g1_normal[k], uv, np.r_[f[nodes, 0], f[nodes, 1]])
        gd = np.r_[V.dof_ux[nodes], V.dof_uy[nodes]]
        idx = [int(np.nonzero(rows[k] == d)[0][0]) for d in gd]
        assert H.rel(Af, B[k][np.ix_(idx, idx)]) < 1e-13 and H.rel(Rf, Rv[k][idx]) < 1e-13
        off = np.ones(12, bool)
        off[idx] = False
        assert np.abs(B[k][off]).max() < 1e-14                     # only the facet's own nodes carry trace


@pytest.mark.parametrize("which", ["square32", "lshape"])
def test_host_lu_analysis_solves_the_newton_matrix(lib, which):
    V = H.square32() if which == "square32" else H.lshape()
    O = FEOracle(V, 1.0)
    n = V.ndofs
    w = 0.1 * np.random.default_rng(0).standard_normal(n)
    vals = O.on_pattern(O.forward_jacobian(w))
    A = sp.csr_matrix((vals, V.csr_col, V.csr_rowptr), shape=(n, n))
    xy = np.zeros((n, 2))
    xy[V.dof_ux], xy[V.dof_uy] = V.node_coords, V.node_coords
    xy[V.dof_p] = V.node_coords[:V.mesh.num_vertices]
    b = np.random.default_rng(1).standard_normal(n)
    x, nnz_lu, p, q = capi.host_lu_probe(V.csr_rowptr, V.csr_col, vals, xy, b)
    assert np.abs(A @ x - b).max() / np.abs(b).max() < 1e-10
    assert sorted(p.tolist()) == list(range(n)) and sorted(q.tolist()) == list(range(n))
    assert nnz_lu < 12 * vals.size                                  # nested dissection keeps the fill modest


@pytest.mark.parametrize("which", ["square32", "lshape"])
def test_multifrontal_analysis_solves_forward_and_adjoint_matrices(lib, which):
    """Symbolic analysis (ND tree, pressure lifting, fronts, extend-add maps) + host restatement of the numeric
    phase; the jittered L-shape needs the pressure lifting to keep every restricted pivot block non-singular."""
    V = H.square32() if which == "square32" else H.lshape(20, 0.2)
    O = FEOracle(V, 1.0)
    n = V.ndofs
    w = 0.1 * np.random.default_rng(0).standard_normal(n)
    xy = np.zeros((n, 2))
    xy[V.dof_ux], xy[V.dof_uy] = V.node_coords, V.node_coords
    xy[V.dof_p] = V.node_coords[:V.mesh.num_vertices]
    kind = np.zeros(n, np.uint8)
    kind[V.dof_p] = 1
    b = np.random.default_rng(1).standard_normal(n)
    for M in (O.forward_jacobian(w), O.adjoint_matrix(w), O.forward_jacobian(0 * w)):
        vals = O.on_pattern(M)
        A = sp.csr_matrix((vals, V.csr_col, V.csr_rowptr), shape=(n, n))
        for window in (1, 1 << 30):     # static pivoting (what the GPU runs) vs restricted partial pivoting
            x, st = capi.host_mf_probe(V.csr_rowptr, V.csr_col, vals, xy, kind, b, pivot_window=window)
            assert np.abs(A @ x - b).max() / np.abs(b).max() < 1e-11
            assert st["min_pivot"] > 1e-6 and st["levels"] <= 16


def test_scripts_parse():
    """bench.py, the graft entry and every tool compile (they only run on the GPU box)."""
    import ast
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")] + glob.glob(os.path.join(root, "tools", "*.py"))
    assert len(files) >= 6
    for f in files:
        ast.parse(open(f).read(), filename=f)


def _header_struct_fields():
    """(name, kind) of every member of ocp_problem_desc in include/ocp_b200.h, in declaration order."""
    txt = open(HEADER).read()
    body = txt[txt.index("typedef struct {"):txt.index("} ocp_problem_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for stmt in body.split(";"):
        stmt = " ".join(stmt.replace("typedef struct {", "").split())
        if not stmt:
            continue
        m = re.match(r"(const )?(int32_t|double) (.*)", stmt)
        assert m, stmt
        for name in m.group(3).split(","):
            name = name.strip()
            kind = "ptr" if name.startswith("*") else m.group(2)
            out.append((name.lstrip("*"), kind))
    return out


def test_problem_desc_bindings_match_the_header():
    """The ctypes struct shipped in capi.py AND the one printed in INTEGRATION.md (what a maintainer would paste into
    the reference repo) mirror ocp_problem_desc field for field - a missing member silently mis-binds every pointer
    after it."""
    import ctypes as C
    kinds = {C.c_int32: "int32_t", C.c_double: "double", C.c_void_p: "ptr"}
    hdr = _header_struct_fields()
    assert [(n, kinds[t]) for n, t in capi.ProblemDesc._fields_] == hdr
    md = open(os.path.join(H.ROOT, "INTEGRATION.md")).read()
    code = md[md.index("class ProblemDesc(C.Structure):"):md.index("def p(a): return")]
    ns = {"C": C}
    exec(code, ns)
    assert [(n, kinds[t]) for n, t in ns["ProblemDesc"]._fields_] == hdr
    assert C.sizeof(ns["ProblemDesc"]) == C.sizeof(capi.ProblemDesc)


@pytest.mark.parametrize("mesh", ["square32", "lshape", "square20"])
def test_gather_assembly_tables_reproduce_the_oracle_matrix(lib, mesh):
    """The atomic-free assembly is organised by CSR row (CTAs of consecutive rows, one (row, cell) pair per thread,
    rounds over the cells of a row).  Its tables are emulated on the host with the kernels' element arithmetic:
    cell part of the Newton matrix and residual, and the transposition permutation used for the adjoint operator."""
    from ocp_b200.fespace import TaylorHood
    from ocp_b200.mesh import square_mesh
    V = {"square32": H.square32, "lshape": H.lshape, "square20": lambda: TaylorHood(square_mesh(20))}[mesh]()
    O = FEOracle(V, 0.7)
    w = 0.3 * np.random.default_rng(2).standard_normal(V.ndofs)
    vals, res, vt, st = capi.host_gather_probe(V, w, 0.7)
    n = V.ndofs
    Jc = sp.coo_matrix((O.cell_jacobian(w).reshape(-1), (O._rows, O._cols)), shape=(n, n)).tocsr()
    assert H.rel(vals, O.on_pattern(Jc)) < 1e-12
    assert H.rel(vt, O.on_pattern(Jc.T.tocsr())) < 1e-12
    Rc = np.zeros(n)
    np.add.at(Rc, V.cell_dofs.reshape(-1), O.cell_residual(w).reshape(-1))
    assert H.rel(res, Rc) < 1e-12
    assert 1 <= st["rounds"] <= 15 and st["colors"] >= 2 and st["smem_bytes"] <= 40 * 1024
    assert st["ctas"] >= 15 * V.mesh.num_cells // 256
