"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle and the committed FEniCS
fixtures.  Bars (BASELINE.json north_star): cell indices and CSR pattern bit-exact, assembled matrices 1e-12 rel,
trajectories 1e-10 (we get bit-exact), cost and gradient 1e-8 rel."""
import os

import numpy as np
import pytest
import torch

import helpers as H
from ocp_b200 import capi
from ocp_b200.pipeline import OCP, Knobs, Parameters, State, initial_control
from oracle.buoy_oracle import BuoyOracle
from oracle.fe_oracle import FEOracle

pytestmark = pytest.mark.gpu
MAT_TOL = 1e-12
TRAJ_TOL = 1e-10
COST_TOL = 1e-8


def dev():
    return torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


@pytest.fixture(scope="module")
def sq():
    V = H.square32()
    xr, ud = H.traj(100)
    ocp = OCP(V, Parameters(), xr[:, 0, :].copy(), ud, device=dev())
    yield V, ocp
    ocp.close()


def test_extension_is_loaded_and_on_gpu():
    lib = capi.load_library()
    assert lib.ocp_device_available() == 0
    with open("/proc/self/maps") as fh:
        assert "libocp_b200.so" in fh.read()


@pytest.mark.parametrize("K", [2, 4, 6, 10, 100, 400])
def test_buoy_forward_bit_exact_vs_oracle_and_fenics(K):
    V = H.square32()
    xr, ur = H.traj(K)
    w = H.field_for(K)
    ocp = OCP(V, Parameters(), xr[:, 0, :].copy(), ur, device=dev())
    cell = torch.empty((200, K), device=dev(), dtype=torch.int32)
    ocp._primal(T(w), ocp.d_x, ocp.d_u, ocp.d_mask, d_cell=cell)
    x, u = ocp._to_reference_layout(ocp.d_x), ocp._to_reference_layout(ocp.d_u)
    xo, uo, co, mo, po = BuoyOracle(V, brute=True).forward(V.velocity_nodal(w), xr[:, 0, :], 200, H.H, H.CENTER)
    assert np.array_equal(ocp._cells_to_host(cell), co)         # point-location cell indices: bit-exact
    assert np.array_equal(x, xo) and np.array_equal(u, uo)      # trajectories: bit-exact vs the oracle
    assert np.abs(x - xr).max() < 4e-16 and np.abs(u - ur).max() < 4e-16   # and round-off level vs FEniCS
    assert float(ocp.d_mask.sum()) == 0
    # host-buffer entry point = the reference's solve_primal_ode(wSol, buoy_mask)
    mask = np.zeros(K)
    x2, u2 = ocp.ctx.solve_primal_ode_host(w, xr[:, 0, :], mask)
    assert np.array_equal(x2, xo) and np.array_equal(u2, uo) and mask.sum() == 0
    ocp.close()


def test_buoy_forward_masks_and_parks_like_the_reference():
    V = H.lshape()
    P = Parameters()
    O = FEOracle(V, 1.0)
    f = initial_control(V, "PL")
    w = O.newton_solve(f)
    rng = np.random.default_rng(3)
    K = 700
    x0 = np.stack([rng.uniform(-0.05, 2.05, K), rng.uniform(-0.05, 2.05, K)], 1)
    # two buoys that leave only with the last step, to hit the "parked" branch
    ocp = OCP(V, P, x0, np.zeros((K, 200, 2)), device=dev())
    mask = np.zeros(K)
    x, u = ocp.solve_primal_ode(State(T(w)), mask)
    xo, uo, co, mo, po = BuoyOracle(V, brute=True).forward(V.velocity_nodal(w), x0, 200, H.H, ocp.center_of_domain)
    assert mo.sum() > 50                                         # many start outside the L
    assert np.array_equal(mask, mo) and np.array_equal(x, xo) and np.array_equal(u, uo)
    assert np.array_equal(ocp._buoy_vector_to_host(ocp.d_parked), po)
    ocp.close()


def test_parked_last_sample_branch():
    V = H.square32()
    vel = np.zeros((V.num_nodes, 2))
    vel[:, 0] = 1.0
    w = np.zeros(V.ndofs)
    w[V.dof_ux] = 1.0
    x0 = np.array([[1.9, 1.0], [2.5, 1.0], [0.5, 0.5], [2.0 - 198 * H.H - 1e-9, 0.7]])
    ud = np.random.default_rng(0).standard_normal((4, 200, 2))
    ocp = OCP(V, Parameters(), x0, ud, device=dev())
    mask = np.zeros(4)
    x, u = ocp.solve_primal_ode(State(T(w)), mask)
    B = BuoyOracle(V, brute=True)
    xo, uo, co, mo, po = B.forward(vel, x0, 200, H.H, H.CENTER)
    assert np.array_equal(x, xo) and np.array_equal(u, uo) and mask.tolist() == [1, 1, 0, 0]
    assert ocp.d_parked.cpu().numpy().tolist() == [0, 0, 0, 1]
    # backward sweep with a parked buoy: the scatter re-evaluates u(centre)
    g = np.random.default_rng(1).standard_normal((V.mesh.num_vertices, 4))
    nn = V.num_nodes
    acc = torch.zeros(2 * nn + 2, device=dev(), dtype=torch.float64)
    mu = torch.empty_like(ocp.d_x)
    ocp.ctx.buoy_adjoint_scatter(T(vel), T(g), 4, ocp.d_x, ocp.d_u, ocp.d_ud, ocp.d_mask, ocp.d_parked, mu, acc)
    muo = B.adjoint(g, xo, uo, ud, mo, H.H)
    bo = B.point_sources(vel, xo, ud, muo, mo, H.H, H.CENTER)
    assert H.rel(ocp._to_reference_layout(mu), muo) < 1e-12
    assert H.rel(acc[:2 * nn].cpu().numpy().reshape(-1, 2), bo) < 1e-12
    assert abs(float(acc[2 * nn]) - B.misfit(uo, ud, H.H)) / B.misfit(uo, ud, H.H) < 1e-13
    assert float(acc[2 * nn + 1]) == 2.0
    ocp.close()


def test_assembled_matrices_and_residual(sq):
    V, ocp = sq
    O = FEOracle(V, 1.0)
    rng = np.random.default_rng(0)
    w = 0.3 * rng.standard_normal(V.ndofs)
    f = initial_control(V, "OCP")
    vals = torch.zeros(V.csr_col.size, device=dev(), dtype=torch.float64)
    res = torch.zeros(V.ndofs, device=dev(), dtype=torch.float64)
    ocp.ctx.assemble_forward(T(w), T(f), vals, res, False)
    assert H.rel(vals.cpu().numpy(), O.on_pattern(O.jacobian_unconstrained(w))) < MAT_TOL
    assert H.rel(res.cpu().numpy(), O.forward_residual(w, f)) < MAT_TOL
    ocp.ctx.assemble_forward(T(w), T(f), vals, res, True)
    assert H.rel(vals.cpu().numpy(), O.on_pattern(O.forward_jacobian(w))) < MAT_TOL
    r = O.forward_residual(w, f)
    r[V.dirichlet_dofs] = w[V.dirichlet_dofs]
    assert H.rel(res.cpu().numpy(), r) < MAT_TOL
    ocp.ctx.assemble_adjoint(T(w), vals, True)
    assert H.rel(vals.cpu().numpy(), O.on_pattern(O.adjoint_matrix(w))) < MAT_TOL


def test_adjoint_matrix_ignores_viscosity():
    """aAdj has no viscosity factor (OCP_dolfin.py:344): it is the transpose of the nu=1 Jacobian even when nu != 1."""
    V = H.lshape()
    P = Parameters(viscosity=0.01)
    ocp = OCP(V, P, np.array([[1.5, 0.5]]), np.zeros((1, 200, 2)), device=dev())
    O = FEOracle(V, 0.01)
    w = 0.2 * np.random.default_rng(5).standard_normal(V.ndofs)
    vals = torch.zeros(V.csr_col.size, device=dev(), dtype=torch.float64)
    ocp.ctx.assemble_adjoint(T(w), vals, True)
    assert H.rel(vals.cpu().numpy(), O.on_pattern(O.adjoint_matrix(w))) < MAT_TOL
    ocp.ctx.assemble_forward(T(w), T(np.zeros((V.num_nodes, 2))), vals, None, True)
    assert H.rel(vals.cpu().numpy(), O.on_pattern(O.forward_jacobian(w))) < MAT_TOL
    ocp.close()


def test_newton_matches_oracle_iterate_by_iterate(sq):
    V, ocp = sq
    O = FEOracle(V, 1.0)
    for name in ("PL", "OCP"):
        f = initial_control(V, name)
        st = ocp.forward_solve(T(f))
        wo, its, hist = O.newton_solve(f, return_history=True)
        assert ocp.last_newton_its == its == 3
        assert np.allclose(ocp.last_res_hist[:-1], hist[:-1], rtol=1e-6)
        assert H.rel(st.vector(), wo) < 1e-11


def test_newton_reproduces_fenics_u_bar(sq):
    """KAT K3 on the GPU: Newton from zero with the recovered control gives the stored FEniCS state."""
    V, ocp = sq
    O = FEOracle(V, 1.0)
    ubar = H.fields()["u_bar"]
    f, _ = H.recover_control(V, O, ubar)
    st = ocp.forward_solve(T(f))
    assert ocp.last_newton_its == 4
    assert np.abs(st.vector() - ubar).max() < 1e-11


def test_projection_adjoint_chain_cost_against_fenics_artefacts():
    """KAT K4 + K5 on the GPU."""
    V = H.square32()
    O, B = FEOracle(V, 1.0), BuoyOracle(V)
    xr, ud = H.traj(6)
    ocp = OCP(V, Parameters(), xr[:, 0, :].copy(), ud, device=dev())
    ubar = H.fields()["u_bar"]
    st = State(T(ubar))
    g = ocp.project_grad(st)
    assert H.rel(g.cpu().numpy(), O.project_gradient(ubar)) < 1e-12
    mask = np.zeros(6)
    x, u = ocp.solve_primal_ode(st, mask)
    q = H.q_nodal(V)
    Jref = H.scalars()["u_bar_chapter_6.3.3"]["J_array"][0]
    assert abs(ocp.J(u, q) - Jref) / Jref < 1e-12                # K4: FEniCS' own J_array value
    mu = ocp.solve_adjoint_ode(st, g, x, mask, u)
    muo = B.adjoint(O.project_gradient(ubar), x, u, ud, mask, H.H)
    assert H.rel(mu, muo) < 1e-12
    z = ocp.adjoint_solve(st, x, u, mask, g)
    f, _ = H.recover_control(V, O, ubar)
    zn = V.velocity_nodal(z.vector())
    g1 = np.unique(V.g1_nodes)
    z_implied = (q - (1 - 4.0 * 6e-6) * f) / 4.0
    assert np.abs(zn[g1] - z_implied[g1]).max() < 1e-11           # K5: implied by FEniCS' stored control update
    gradj = ocp.gradient(f, z, np.full((V.num_nodes, 2), 0.1))
    go = O.boundary_inner(6e-6 * f - zn, np.full((V.num_nodes, 2), 0.1))
    assert abs(gradj - go) <= COST_TOL * abs(go)
    d, l2, h1 = ocp.field_norms(st)
    assert abs(d - O.divergence_norm(ubar)) < 1e-13
    ocp.close()


def test_gradient_host_entry_point_matches_oracle():
    V = H.square32()
    xr, ud = H.traj(10)
    ocp = OCP(V, Parameters(), xr[:, 0, :].copy(), ud, device=dev())
    f = initial_control(V, "PL")
    ocp.ctx.set_observations_host(xr[:, 0, :], ud)
    n0 = capi.launch_count()
    w, z, mask, sc = ocp.ctx.gradient_host(f)
    assert capi.launch_count() - n0 > 20
    P = H.OraclePipeline(V, 1.0, xr[:, 0, :].copy(), ud, 1e-5)
    s = P.gradient_step(f)
    assert sc["newton_its"] == s["its"] and sc["n_masked"] == 0
    assert H.rel(w, s["w"]) < 1e-11 and H.rel(z, s["z"]) < 1e-9
    J = sc["misfit"] + 0.5 * 1e-5 * sc["f_norm2"]
    assert abs(J - 0.025045819440590228) < COST_TOL * 0.025
    ocp.close()


def test_gd_loop_with_line_search_matches_oracle_pipeline():
    """OCP_dolfin defaults on the 6-buoy case: cost history, Armijo trial counts and the control."""
    V = H.square32()
    xr, ud = H.traj(6)
    ocp = OCP(V, Parameters(), xr[:, 0, :].copy(), ud, device=dev())
    f0 = initial_control(V, "OCP")
    r = ocp.run(f0, Knobs(num_steps=3, use_line_search=True))
    ref = H.OraclePipeline(V, 1.0, xr[:, 0, :].copy(), ud, 6e-6).run(f0, 3, True)
    assert r.inner_iterations == ref["inner"] == [1, 1, 1] and r.newton_its == [3, 3, 3]
    assert np.allclose(r.J_array, ref["J_array"], rtol=COST_TOL, atol=0)
    assert np.allclose(r.J_array, [0.54412, 0.43381, 0.36199], atol=2e-5)
    g1 = np.unique(V.g1_nodes)
    assert H.rel(r.f[g1], ref["f"][g1]) < 1e-8
    ocp.close()


def test_grad_check_reproduces_K6_table():
    """Pipeline_limits defaults + 10_buoys + grad_check: J0, gradj and the FD error table."""
    V = H.square32()
    xr, ud = H.traj(10)
    ocp = OCP(V, Parameters(), xr[:, 0, :].copy(), ud, device=dev())
    r = ocp.run(initial_control(V, "PL"), Knobs(num_steps=2, use_line_search=False, grad_check=True, exit_rule="ten"))
    t = r.grad_tables
    assert abs(t["J0"] - 0.025045819440590228) < COST_TOL * 0.025
    assert abs(t["gradj"] - (-0.02603602981198921)) < COST_TOL * 0.026
    one = [row[2] for row in t["one_sided"]]
    cen = [row[2] for row in t["centered"]]
    assert np.allclose(one[:4], [1.19e-3, 1.19e-4, 1.19e-5, 1.17e-6], rtol=0.02)
    assert 5e-6 < cen[0] < 6e-6 and all(1.5e-8 < c < 3e-8 for c in cen[2:7])    # intrinsic 2.1e-8 plateau
    assert np.allclose(r.J_array, [0.02505, 0.04713], atol=1e-5)
    ocp.close()


def test_cfg1_masked_buoys_in_first_iteration():
    """cfg1: OCP_dolfin q0 with the 10-buoy set masks 4 of 10 buoys in iteration 0."""
    V = H.square32()
    xr, ud = H.traj(10)
    ocp = OCP(V, Parameters(), xr[:, 0, :].copy(), ud, device=dev())
    f0 = initial_control(V, "OCP")
    r = ocp.run(f0, Knobs(num_steps=1, use_line_search=True))
    s = H.OraclePipeline(V, 1.0, xr[:, 0, :].copy(), ud, 1e-5).gradient_step(f0)
    assert r.n_masked[0] == int(s["mask"].sum()) == 4
    ocp.close()


@pytest.mark.parametrize("K", [100_000, 1_000_003])
def test_large_sweep_properties(K):
    """Size-independent properties at sweep sizes the oracle cannot reach: exact Euler recursion, misfit
    identities, partition of unity of the deposited sources, linearity of the scatter in gamma."""
    V = H.square32()
    rng = np.random.default_rng(0)
    x0 = np.stack([rng.uniform(0.05, 1.95, K), rng.uniform(0.05, 1.95, K)], 1)
    ocp = OCP(V, Parameters(), x0, np.zeros((K, 200, 2)), device=dev())
    w = T(H.field_for(100))
    ocp._primal(w, ocp.d_x, ocp.d_u, ocp.d_mask)
    x, u = ocp.d_x, ocp.d_u
    ok = ocp.d_mask == 0
    assert int(ok.sum()) > 0.9 * K
    # x_{k+1} == x_k + h u_k exactly (two roundings), for every buoy that stayed inside
    inside = ok & (ocp.d_parked == 0)
    assert bool((x[1:, inside] == x[:-1, inside] + H.H * u[:-1, inside]).all())
    # parked buoys: only the last sample was moved to the centre (OCP_dolfin.py:226-229)
    pk = ocp.d_parked != 0
    if int(pk.sum()) > 0:
        assert bool((x[1:-1, pk] == x[:-2, pk] + H.H * u[:-2, pk]).all())
        assert bool((x[-1, pk] == torch.tensor(H.CENTER, device=x.device)).all()) and bool((u[-1, pk] == 0).all())
    # a sample of buoys against the oracle, bit-exact
    idx = rng.choice(K, 64, replace=False)
    xo, uo, *_ = BuoyOracle(V).forward(V.velocity_nodal(H.field_for(100)), x0[idx], 200, H.H, H.CENTER)
    didx = ocp.inv_perm[torch.from_numpy(idx).to(x.device)] if ocp.perm is not None else idx     # device order
    assert np.array_equal(x[:, didx].cpu().numpy().transpose(1, 0, 2), xo)
    assert np.array_equal(u[:, didx].cpu().numpy().transpose(1, 0, 2), uo)
    nn = V.num_nodes
    g = torch.zeros((V.mesh.num_vertices, 4), device=dev(), dtype=torch.float64)
    # u_d = u  =>  misfit 0 and, with G = 0 (mu = 0), gamma = 0: nothing is deposited
    ocp.d_ud.copy_(u)
    acc = torch.zeros(2 * nn + 2, device=dev(), dtype=torch.float64)
    ocp.ctx.buoy_adjoint_scatter(ocp.d_vel, g, K, x, u, ocp.d_ud, ocp.d_mask, ocp.d_parked, None, acc)
    assert float(acc[2 * nn]) == 0.0
    n_parked = int(ocp.d_parked.sum())
    if n_parked == 0:
        assert float(acc[:2 * nn].abs().max()) == 0.0
    else:
        # a parked last sample stores u = 0 but the scatter re-evaluates u(centre):
        # gamma_199 = h (u_d - u(centre)) = -h u(centre) per parked buoy, deposited with partition of unity
        _, uc, *_ = BuoyOracle(V).forward(V.velocity_nodal(H.field_for(100)), H.CENTER[None, :], 200, H.H, H.CENTER)
        tot = acc[:2 * nn].view(-1, 2).sum(0).cpu().numpy()
        assert np.allclose(tot, -H.H * n_parked * uc[0, 0], rtol=1e-9, atol=1e-15)
    assert float(acc[2 * nn + 1]) == float(ocp.d_mask.sum())
    # u_d = u + c  =>  gamma = h c for every sample: sum of b = K_ok * nt * h * c (partition of unity), misfit closed form
    cvec = torch.tensor([0.3, -0.7], device=dev(), dtype=torch.float64)
    ocp.d_ud.copy_(u + cvec)
    acc.zero_()
    ocp.ctx.buoy_adjoint_scatter(ocp.d_vel, g, K, x, u, ocp.d_ud, ocp.d_mask, ocp.d_parked, None, acc)
    nok = int(ok.sum())
    b = acc[:2 * nn].view(-1, 2).sum(0).cpu().numpy()
    expect = nok * 200 * H.H * np.array([0.3, -0.7])
    if n_parked == 0:
        assert np.allclose(b, expect, rtol=1e-10)
    assert abs(float(acc[2 * nn]) - 0.5 * K * 200 * H.H * (0.09 + 0.49)) / (K * 0.58) < 1e-9
    # linearity in the offset: b(2c) - b(0) = 2 (b(c) - b(0))   (b(0) is the parked-buoy term, zero without them)
    acc0 = torch.zeros_like(acc)
    ocp.d_ud.copy_(u)
    ocp.ctx.buoy_adjoint_scatter(ocp.d_vel, g, K, x, u, ocp.d_ud, ocp.d_mask, ocp.d_parked, None, acc0)
    acc2 = torch.zeros_like(acc)
    ocp.d_ud.copy_(u + 2 * cvec)
    ocp.ctx.buoy_adjoint_scatter(ocp.d_vel, g, K, x, u, ocp.d_ud, ocp.d_mask, ocp.d_parked, None, acc2)
    lhs, rhs = acc2[:2 * nn] - acc0[:2 * nn], 2 * (acc[:2 * nn] - acc0[:2 * nn])
    assert float((lhs - rhs).abs().max()) <= 1e-9 * float(acc[:2 * nn].abs().max())
    ocp.close()


def _lshape_problem(K_side=10, seed=0):
    """cfg2: L-shape, synthetic 100-buoy twin experiment (the reference's L-shape run only has K=3 analytic buoys,
    OCP_dolfin.py:169-196): start points on a 10x10 grid in the lower-left block, u_d from a twin forward solve."""
    V = H.lshape(16, 0.15)
    gx, gy = np.meshgrid(np.linspace(0.15, 0.85, K_side), np.linspace(0.15, 0.85, K_side))
    x0 = np.stack([gx.ravel(), gy.ravel()], 1)
    O = FEOracle(V, 1.0)
    f_true = V.interpolate_control(lambda x, y: 0.3 * np.sin(np.pi * y) + 0 * x, lambda x, y: -0.2 * np.cos(np.pi * x), 2)
    w_true = O.newton_solve(f_true)
    _, ud, _, mask, _ = BuoyOracle(V).forward(V.velocity_nodal(w_true), x0, 200, H.H, np.array([1.0, 0.5]))
    assert mask.sum() == 0
    return V, x0, ud


def test_cfg2_lshape_100_buoys_line_search_matches_oracle():
    V, x0, ud = _lshape_problem()
    ocp = OCP(V, Parameters(), x0, ud, device=dev())
    assert np.array_equal(ocp.center_of_domain, [1.0, 0.5])
    f0 = initial_control(V, "PL")
    r = ocp.run(f0, Knobs(num_steps=3, use_line_search=True))
    ref = H.OraclePipeline(V, 1.0, x0, ud, 1e-6 * 100, center=[1.0, 0.5]).run(f0, 3, True)
    assert r.inner_iterations == ref["inner"]
    assert np.allclose(r.J_array, ref["J_array"], rtol=COST_TOL, atol=0)
    assert r.J_array[-1] < r.J_array[0]                                   # Armijo steps decrease the cost
    g1 = np.unique(V.g1_nodes)
    assert H.rel(r.f[g1], ref["f"][g1]) < 1e-7
    ocp.close()


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_cfg4_initial_control_cases(case):
    """initial_control_test.py cases 0-3 (6_buoys, no line search, LR = 5): first iterations vs the oracle pipeline."""
    V = H.square32()
    xr, ud = H.traj(6)
    ocp = OCP(V, Parameters(), xr[:, 0, :].copy(), ud, device=dev())
    f0 = initial_control(V, "ICT", case)
    r = ocp.run(f0, Knobs(num_steps=2, use_line_search=False, exit_rule="ten"))
    ref = H.OraclePipeline(V, 1.0, xr[:, 0, :].copy(), ud, 6e-6).run(f0, 2, False)
    assert np.allclose(r.J_array, ref["J_array"], rtol=COST_TOL, atol=1e-14)
    ocp.close()


def test_device_buoy_order_does_not_change_results():
    """Spatial sorting of the buoys on the device is invisible at the host boundary."""
    V = H.square32()
    rng = np.random.default_rng(7)
    K = 777
    x0 = np.stack([rng.uniform(0.05, 1.95, K), rng.uniform(0.05, 1.95, K)], 1)
    ud = 0.1 * rng.standard_normal((K, 200, 2))
    w = T(H.field_for(100))
    outs = []
    for sort in (True, False):
        ocp = OCP(V, Parameters(), x0, ud, device=dev(), sort_buoys=sort)
        assert (ocp.perm is not None) == sort
        mask = np.zeros(K)
        x, u = ocp.solve_primal_ode(State(w), mask)
        g = ocp.project_grad(State(w))
        mu = ocp.solve_adjoint_ode(State(w), g, x, mask, u)
        z = ocp.adjoint_solve(State(w), x, u, mask, g)
        outs.append((x, u, mask, mu, z.vector(), ocp.J(u, np.zeros((V.num_nodes, 2)))))
        ocp.close()
    a, b = outs
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert H.rel(a[3], b[3]) < 1e-12                     # mu: per buoy, but grad(u) projection is atomically assembled per context
    assert H.rel(a[4], b[4]) < 1e-11 and abs(a[5] - b[5]) <= 1e-12 * abs(b[5])   # sums: order changes


@pytest.mark.parametrize("K,nu", [(10, 0.01), (6, 1.0)])
def test_twin_experiment_regenerates_reference_data(K, nu):
    """the u_d generator on the GPU path reproduces the reference's stored field and its
    x_0_array / u_d_array files."""
    from ocp_b200.twin import TwinExperiment
    V = H.square32()
    xr, ur = H.traj(K)
    if K == 6:
        inflow = lambda x, y: (-np.cos(np.pi * x) * np.sin(np.pi * y), np.sin(np.pi * x) * np.cos(np.pi * y))
    else:
        inflow = lambda x, y: (0.1 + 0 * x, 0 * x)
    tw = TwinExperiment(V, viscosity=nu, inflow=inflow, device=dev())
    w, x, u = tw.solve(xr[:, 0, :])
    assert tw.newton_its == 3
    assert np.abs(w - H.field_for(K)).max() < 5e-12
    assert np.abs(x - xr).max() < 1e-11 and np.abs(u - ur).max() < 1e-11


def test_error_norm_table_matches_oracle(sq):
    """Pipeline_limits.py:433-443: L2 / H1 norm of u - u_bar."""
    V, ocp = sq
    O = FEOracle(V, 1.0)
    w = H.field_for(100)
    ubar = H.fields()["u_bar"]
    l2, h1 = ocp.error_norms(State(T(w)), ubar)
    l2o, h1o = O.l2_h1_norms(w - ubar)
    assert abs(l2 - l2o) < 1e-12 * l2o and abs(h1 - h1o) < 1e-12 * h1o


def _sharded_worker(rank, world, port, q):
    import os
    import sys
    sys.path.insert(0, H.ROOT)
    sys.path.insert(0, os.path.join(H.ROOT, "tests"))
    import torch.distributed as dist
    from ocp_b200.sharding import shard_buoys
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # gloo: both ranks share the one test GPU
    try:
        V = H.square32()
        xr, ud = H.traj(100)
        x0, ud_loc = shard_buoys(xr[:, 0, :].copy(), ud, rank, world)
        ocp = OCP(V, Parameters(), x0, ud_loc, device=torch.device("cuda:0"), group=dist.group.WORLD)
        assert ocp.K_global == 100 and abs(ocp.alpha - 1e-4) < 1e-18
        r = ocp.run(initial_control(V, "PL"), Knobs(num_steps=2, use_line_search=True))
        q.put((rank, r.J_array, r.inner_iterations, r.f, float(ocp.d_acc[2 * V.num_nodes])))
        ocp.close()
    finally:
        dist.destroy_process_group()


def test_buoys_sharded_over_two_ranks_match_single_rank():
    """buoys partitioned over ranks, one all-reduce of [b | misfit | n_masked] per gradient evaluation
    (plus the misfit scalar in the line search); every rank ends with the single-rank cost history and control."""
    import torch.multiprocessing as mp
    V = H.square32()
    xr, ud = H.traj(100)
    ocp = OCP(V, Parameters(), xr[:, 0, :].copy(), ud, device=dev())
    ref = ocp.run(initial_control(V, "PL"), Knobs(num_steps=2, use_line_search=True))
    ocp.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, J, inner, f, misfit = q.get(timeout=600)
        got[rank] = (J, inner, f, misfit)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    g1 = np.unique(V.g1_nodes)
    for rank in (0, 1):
        J, inner, f, misfit = got[rank]
        assert inner == ref.inner_iterations
        assert np.allclose(J, ref.J_array, rtol=1e-11, atol=0)
        assert H.rel(f[g1], ref.f[g1]) < 1e-9
    assert got[0][3] == got[1][3]                                # the all-reduced accumulator is identical on both ranks


def test_run_writes_the_reference_output_files_and_resumes(tmp_path):
    """Output files of a run (names / formats of OCP_dolfin.py:476-511, 578-588) and resume from the control
    checkpoint (OCP_dolfin.py:151-160)."""
    from ocp_b200 import checkpoint
    V = H.square32()
    xr, ud = H.traj(6)
    ocp = OCP(V, Parameters(), xr[:, 0, :].copy(), ud, device=dev())
    out = str(tmp_path / "run")
    f0 = initial_control(V, "OCP")
    r3 = ocp.run(f0, Knobs(num_steps=3, use_line_search=True))
    r2 = ocp.run(f0, Knobs(num_steps=2, use_line_search=True), out_dir=out)
    for name in ("timings.txt", "u_divergence.txt", "variables.txt", "J_array.npy", "J.svg", "q_backup/q.xdmf", "q_backup/q.h5",
                 "checkpoints/q.h5", "paraview/checkpoint/u.h5", "paraview/checkpoint/p.xdmf"):
        assert os.path.exists(os.path.join(out, name)), name
    assert np.allclose(np.load(os.path.join(out, "J_array.npy")), r2.J_array)
    assert "buoy count: 6" in open(os.path.join(out, "variables.txt")).read()
    # resume: third iteration from the checkpointed control equals the uninterrupted run
    q = checkpoint.read_control(os.path.join(out, "checkpoints", "q.h5"), V)
    assert np.array_equal(q, r2.f)
    from ocp_b200 import h5lite
    assert h5lite.checkpoint_counters(os.path.join(out, "checkpoints", "q.h5"), "f") == [0, 1]   # one group per iteration
    r1 = ocp.run(q, Knobs(num_steps=1, use_line_search=True), LR=r2.LR)
    assert abs(r1.J_array[0] - r3.J_array[2]) <= 1e-10 * abs(r3.J_array[2])
    w = checkpoint.read_state(os.path.join(out, "paraview/checkpoint/u.h5"), V, "u")
    assert w.shape == (V.ndofs,)
    ocp.close()


@pytest.mark.parametrize("n,big", [(48, 160), (32, 96)])
def test_refined_mesh_large_front_solver_matches_oracle(n, big, monkeypatch):
    """Refined square meshes (cfg5) take the large-front group kernels of the multifrontal LU.  OCP_MF_BIG lowers the
    front order at which a front is 'large' so that the path is exercised at a size the oracle finishes in seconds:
    one full gradient evaluation (Newton, projection with four right-hand sides, transposed adjoint solve) against
    the oracle pipeline on the same mesh."""
    from ocp_b200.fespace import TaylorHood
    from ocp_b200.mesh import square_mesh
    monkeypatch.setenv("OCP_MF_BIG", str(big))
    V = H.square32() if n == 32 else TaylorHood(square_mesh(n))
    rng = np.random.default_rng(3)
    K = 64
    x0 = np.stack([rng.uniform(0.2, 1.0, K), rng.uniform(0.3, 1.7, K)], 1)
    f = initial_control(V, "PL")
    P = H.OraclePipeline(V, 1.0, x0, np.zeros((K, 200, 2)), 1e-6 * K)
    _, ud, _, _, _ = P.primal(P.O.newton_solve(1.3 * f))          # twin observations from a different control
    P.ud = ud
    s = P.gradient_step(f)
    ocp = OCP(V, Parameters(), x0, ud, device=dev())
    ocp.ctx.set_observations_host(x0, ud)
    w, z, mask, sc = ocp.ctx.gradient_host(f)
    assert sc["newton_its"] == s["its"] and sc["n_masked"] == int(s["mask"].sum())
    assert H.rel(w, s["w"]) < 1e-11
    assert H.rel(z, s["z"]) < 1e-9
    g = ocp.project_grad(State(T(s["w"])))
    assert H.rel(g.cpu().numpy(), s["g"]) < 1e-11
    J = sc["misfit"] + 0.5 * 1e-6 * K * sc["f_norm2"]
    Jo = P.cost(s["u"], f)
    assert abs(J - Jo) <= COST_TOL * abs(Jo)
    st = ocp.ctx.solver_stats()
    assert st["n_factor"] >= 2
    ocp.close()


def test_solver_variants_of_round_two_agree_with_the_oracle_and_with_each_other(monkeypatch):
    """The factor / solve variants introduced in round 2, each against the oracle on the reference mesh and against
    the default build: programmatic dependent launches between the solve levels, the side-stream assembly of the
    adjoint operator, the inverse kernels on a side stream (same arithmetic, different launch mechanics), extend-add
    child by child vs fp64 atomics, the shared-memory-resident leaf kernel on 0 / 1 / 3 bottom levels (a different
    summation order of the Schur complement).  Everything agrees to rounding (measured on B200: w 1e-16 ... 7e-16,
    z 8e-15 ... 1.6e-14 between the variants); the triangular solves still add the children's contributions to the
    right-hand side with atomics, so not even two runs of one build are bitwise equal."""
    V = H.square32()
    rng = np.random.default_rng(11)
    K = 48
    x0 = np.stack([rng.uniform(0.2, 1.0, K), rng.uniform(0.3, 1.7, K)], 1)
    f = initial_control(V, "PL")
    P = H.OraclePipeline(V, 1.0, x0, np.zeros((K, 200, 2)), 1e-6 * K)
    _, ud, _, _, _ = P.primal(P.O.newton_solve(1.3 * f))
    P.ud = ud
    s = P.gradient_step(f)

    def run(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ocp = OCP(V, Parameters(), x0, ud, device=dev())
        ocp.ctx.set_observations_host(x0, ud)
        out = [ocp.ctx.gradient_host(f) for _ in range(3)]       # plain path, graph capture, graph replay
        ocp.close()
        for k in env:
            monkeypatch.delenv(k)
        for w, z, _, sc in out:
            assert sc["newton_its"] == s["its"]
            assert H.rel(w, s["w"]) < 1e-11 and H.rel(z, s["z"]) < 1e-9
        return out

    base = run({})
    for w, z, _, _ in base[1:]:                                  # plain path, captured graph, replayed graph
        assert H.rel(w, base[0][0]) < 1e-12 and H.rel(z, base[0][1]) < 1e-10
    for env in ({"OCP_MF_PDL": "0"}, {"OCP_STEP_OVERLAP": "0"}, {"OCP_MF_DINV_PAIR": "1"}, {"OCP_MF_OVERLAP": "1"},
                {"OCP_MF_OVERLAP_FROM": "0"}, {"OCP_MF_OVERLAP_FROM": "1"}, {"OCP_MF_OVERLAP_FROM": "3"},
                {"OCP_MF_EA_ATOMIC": "1"}, {"OCP_MF_LEAF_LEVELS": "0"}, {"OCP_MF_LEAF_LEVELS": "3"}):
        out = run(env)
        assert H.rel(out[0][0], base[0][0]) < 1e-12 and H.rel(out[0][1], base[0][1]) < 1e-10, env


def test_ensemble_of_independent_cases_runs_concurrently_with_identical_results():
    """cfg4: independent cases on one GPU, one context / stream / host thread each (ocp_b200.ensemble).  The concurrent
    run must reproduce the one-after-the-other run and the oracle's cost of every case."""
    from ocp_b200.ensemble import Case, Ensemble
    V = H.square32()
    xr6, ud6 = H.traj(6)
    xr10, ud10 = H.traj(10)
    cases = [Case(f"ICT {c}", xr6[:, 0, :].copy(), ud6, initial_control(V, "ICT", c)) for c in range(4)]
    cases.append(Case("PL 10", xr10[:, 0, :].copy(), ud10, initial_control(V, "PL")))
    E = Ensemble(V, cases, device=dev())
    seq = E.gradients(concurrent=False)
    for _ in range(3):
        con = E.gradients(concurrent=True)
        for a, b in zip(seq, con):
            assert a["newton_its"] == b["newton_its"] and a["n_masked"] == b["n_masked"]
            assert abs(a["J"] - b["J"]) <= 1e-12 * abs(a["J"])
            assert H.rel(b["w"], a["w"]) < 1e-13 and H.rel(b["z"], a["z"]) < 1e-10
    s = H.OraclePipeline(V, 1.0, xr10[:, 0, :].copy(), ud10, 1e-5).gradient_step(cases[4].f0)
    assert abs(con[4]["J"] - 0.025045819440590228) < COST_TOL * 0.025
    g1 = np.unique(V.g1_nodes)
    assert H.rel(con[4]["grad"][g1], s["grad"][g1]) < 1e-8
    h2 = E.descend(2, concurrent=True)
    h1 = E.descend(2, concurrent=False)
    assert np.allclose(np.array(h1), np.array(h2), rtol=1e-11, atol=0)
    E.close()
