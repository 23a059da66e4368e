"""Parity at the BENCHMARKED sizes and on the paths the small cases do not reach (round-2 gate):

  * cfg3, K = 10 000 (the headline bench workload): w, b, z, gradient and J of one full gradient evaluation;
  * K = 20 000 >= 8 cells' worth per cell: the per-SM private-copy scatter, element-wise against the oracle;
  * 128 x 128 mesh with the DEFAULT large-front threshold (fronts up to 1559, many-CTA groups), against the
    committed oracle fixture tests/golden/mesh128_gradient.npz (tools/make_golden_mesh128.py);
  * the reference's own L-shape experiment (K = 3, analytic u_d, OCP_dolfin.py:163-196);
  * the reference-shaped host entry point solve_adjoint_ode.
"""
import os

import numpy as np
import pytest
import torch

import helpers as H
from ocp_b200.pipeline import OCP, Knobs, Parameters, State, initial_control, lshape_reference_observations
from oracle.buoy_oracle import BuoyOracle
from oracle.fe_oracle import FEOracle

pytestmark = pytest.mark.gpu
COST_TOL = 1e-8


def dev():
    return torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def cfg3_problem():
    """BASELINE.json config 3: 100 x 100 start grid, u_d advected through the stored 10000_buoys field."""
    V = H.square32()
    gx, gy = np.meshgrid(np.linspace(0.1, 0.4, 100), np.linspace(0.25, 1.75, 100))
    x0 = np.stack([gx.ravel(), gy.ravel()], 1)
    _, ud, _, mask, _ = BuoyOracle(V).forward(V.velocity_nodal(H.field_for(10000)), x0, 200, H.H, H.CENTER)
    assert mask.sum() == 0
    return V, x0, ud


def test_cfg3_10000_buoys_full_gradient_vs_oracle():
    V, x0, ud = cfg3_problem()
    K = x0.shape[0]
    f = initial_control(V, "PL")
    P = H.OraclePipeline(V, 1.0, x0, ud, 1e-6 * K)
    s = P.gradient_step(f)
    ocp = OCP(V, Parameters(), x0, ud, device=dev())
    ocp.set_control(f)
    ocp.gradient_step(ocp.d_f)
    nn = V.num_nodes
    w = ocp.d_w.cpu().numpy()
    assert ocp.last_newton_its == s["its"] == 3
    assert H.rel(w, s["w"]) < 1e-11
    x, u = ocp._to_reference_layout(ocp.d_x), ocp._to_reference_layout(ocp.d_u)
    assert np.abs(x - s["x"]).max() < 1e-10 and np.abs(u - s["u"]).max() < 1e-10
    acc = ocp.d_acc.cpu().numpy()
    b = acc[:2 * nn].reshape(-1, 2)
    assert H.rel(b, s["bnode"]) < 1e-11                       # end to end (two different LU solvers upstream)
    # the same chain evaluated by the oracle on the GPU's own state: isolates the buoy kernels -> 1e-12
    B, O = P.B, P.O
    g_o = O.project_gradient(w)
    xo, uo, co, mo, po = B.forward(V.velocity_nodal(w), x0, 200, H.H, H.CENTER)
    assert np.array_equal(x, xo) and np.array_equal(u, uo)    # bit-exact on an identical field
    muo = B.adjoint(g_o, xo, uo, ud, mo, H.H)
    bo = B.point_sources(V.velocity_nodal(w), xo, ud, muo, mo, H.H, H.CENTER)
    assert H.rel(ocp.d_g.cpu().numpy(), g_o) < 1e-12
    assert H.rel(b, bo) < 1e-12
    assert abs(acc[2 * nn] - B.misfit(uo, ud, H.H)) <= 1e-12 * B.misfit(uo, ud, H.H) and acc[2 * nn + 1] == 0
    z = ocp.d_z.cpu().numpy()
    assert H.rel(z, s["z"]) < 1e-9
    g1 = np.unique(V.g1_nodes)
    assert H.rel(ocp.d_grad.cpu().numpy()[g1], s["grad"][g1]) < COST_TOL
    J, Jo = ocp._cost_from_acc(ocp.d_f), P.cost(s["u"], f)
    assert abs(J - Jo) <= COST_TOL * abs(Jo)
    # and through the C-ABI host-buffer call the bench times
    ocp.ctx.set_observations_host(x0, ud)
    w2, z2, mask2, sc = ocp.ctx.gradient_host(f)
    assert sc["newton_its"] == 3 and sc["n_masked"] == 0
    assert H.rel(w2, s["w"]) < 1e-11 and H.rel(z2, s["z"]) < 1e-9
    assert abs(sc["misfit"] + 0.5 * 1e-6 * K * sc["f_norm2"] - Jo) <= COST_TOL * abs(Jo)
    ocp.close()


@pytest.mark.parametrize("K", [20_000, 40_001])
def test_private_copy_scatter_elementwise_vs_oracle(K):
    """K >= 8 * #cells switches the point-source deposit to per-SM private copies of b (buoy_private_copies);
    b, mu and the misfit element-wise against the oracle, trajectories bit-exact."""
    V = H.square32()
    assert K >= 8 * V.mesh.num_cells
    rng = np.random.default_rng(11)
    x0 = np.stack([rng.uniform(0.05, 1.95, K), rng.uniform(0.05, 1.95, K)], 1)
    w = H.field_for(100)
    vel = V.velocity_nodal(w)
    B, O = BuoyOracle(V), FEOracle(V, 1.0)
    xo, uo, co, mo, po = B.forward(vel, x0, 200, H.H, H.CENTER)
    ud = 1.1 * uo + 0.01 * rng.standard_normal(uo.shape)
    g = O.project_gradient(w)
    muo = B.adjoint(g, xo, uo, ud, mo, H.H)
    bo = B.point_sources(vel, xo, ud, muo, mo, H.H, H.CENTER)
    ocp = OCP(V, Parameters(), x0, ud, device=dev())
    ocp._primal(T(w), ocp.d_x, ocp.d_u, ocp.d_mask)
    x, u = ocp._to_reference_layout(ocp.d_x), ocp._to_reference_layout(ocp.d_u)
    assert np.array_equal(x, xo) and np.array_equal(u, uo)
    assert np.array_equal(ocp._buoy_vector_to_host(ocp.d_mask), mo)
    nn = V.num_nodes
    acc = torch.zeros(2 * nn + 2, device=dev(), dtype=torch.float64)
    mu = torch.empty_like(ocp.d_x)
    ocp.ctx.buoy_adjoint_scatter(ocp.d_vel, T(g), K, ocp.d_x, ocp.d_u, ocp.d_ud, ocp.d_mask, ocp.d_parked, mu, acc)
    a = acc.cpu().numpy()
    assert H.rel(ocp._to_reference_layout(mu), muo) < 1e-12
    assert H.rel(a[:2 * nn].reshape(-1, 2), bo) < 1e-12
    assert abs(a[2 * nn] - B.misfit(uo, ud, H.H)) <= 1e-12 * B.misfit(uo, ud, H.H)
    assert a[2 * nn + 1] == mo.sum()
    ocp.close()


def test_mesh128_default_large_front_solver_vs_oracle_fixture():
    """cfg5 mesh (148 739 dofs): Newton, projection, buoy sweeps and the transposed adjoint solve with the large
    fronts on the many-CTA group kernels at their DEFAULT threshold, against the oracle's stored results."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk128", os.path.join(H.ROOT, "tools", "make_golden_mesh128.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    G = np.load(os.path.join(H.GOLD, "mesh128_gradient.npz"))
    V, x0, f = mk.problem()
    K, st = x0.shape[0], int(G["stride"])
    assert "OCP_MF_BIG" not in os.environ
    ocp = OCP(V, Parameters(), x0, G["ud"], device=dev())
    ocp.set_control(f)
    ocp.gradient_step(ocp.d_f)
    assert ocp.last_newton_its == int(G["its"])
    assert np.allclose(ocp.last_res_hist[:-1], G["hist"][:-1], rtol=1e-6)
    w, z, g = ocp.d_w.cpu().numpy(), ocp.d_z.cpu().numpy(), ocp.d_g.cpu().numpy()
    assert H.rel(w[::st], G["w_sample"]) < 1e-10 and abs(np.linalg.norm(w) - G["w_norm"]) < 1e-11 * G["w_norm"]
    assert H.rel(g[::st], G["g_sample"]) < 1e-9 and abs(np.linalg.norm(g) - G["g_norm"]) < 1e-10 * G["g_norm"]
    assert H.rel(z[::st], G["z_sample"]) < 1e-8 and abs(np.linalg.norm(z) - G["z_norm"]) < 1e-9 * G["z_norm"]
    g1 = np.unique(V.g1_nodes)
    assert H.rel(V.velocity_nodal(z)[g1], G["z_g1"]) < 1e-8
    assert H.rel(ocp.d_grad.cpu().numpy()[g1], G["grad_g1"]) < COST_TOL
    nn = V.num_nodes
    b = ocp.d_acc.cpu().numpy()[:2 * nn].reshape(-1, 2)
    assert np.array_equal(np.flatnonzero(np.abs(b).sum(1) > 0), G["b_nodes"])
    assert H.rel(b[G["b_nodes"]], G["b_vals"]) < 1e-9
    assert np.abs(ocp._to_reference_layout(ocp.d_x)[:, -1, :] - G["x_last"]).max() < 1e-10
    J = ocp._cost_from_acc(ocp.d_f)
    assert abs(J - float(G["J"])) <= COST_TOL * abs(float(G["J"]))
    ocp.close()


def test_reference_lshape_experiment_K3_analytic_ud():
    """OCP_dolfin.py with L_shape = True: K = 3, analytic u_d(t) on linspace(t0, T, 200) (quirk A.7(1)), sinusoidal
    q0, Armijo line search, alpha = 3e-6, exit rule K/2.  The mshr mesh is not reproducible; same domain, jittered
    structured triangulation, CUDA path against the oracle pipeline on that mesh."""
    V = H.lshape(16, 0.15)
    P = Parameters()
    x0, ud = lshape_reference_observations(P)
    assert ud.shape == (3, 200, 2) and abs(ud[0, 1, 0] - 0.5 * (np.cos(np.pi * (1 / 199 - 0.5)))) < 1e-15
    ocp = OCP(V, P, x0, ud, device=dev())
    assert np.array_equal(ocp.center_of_domain, [1.0, 0.5]) and abs(ocp.alpha - 3e-6) < 1e-20
    f0 = initial_control(V, "OCP")
    r = ocp.run(f0, Knobs(num_steps=3, use_line_search=True))
    ref = H.OraclePipeline(V, 1.0, x0, ud, 3e-6, center=[1.0, 0.5]).run(f0, 3, True)
    assert r.inner_iterations == ref["inner"]
    assert np.allclose(r.J_array, ref["J_array"], rtol=COST_TOL, atol=0)
    g1 = np.unique(V.g1_nodes)
    assert H.rel(r.f[g1], ref["f"][g1]) < 1e-7
    assert abs(r.LR - ref["LR"]) == 0
    # first gradient evaluation in detail
    s = H.OraclePipeline(V, 1.0, x0, ud, 3e-6, center=[1.0, 0.5]).gradient_step(f0)
    ocp.set_control(f0)
    ocp.gradient_step(ocp.d_f)
    assert ocp.last_newton_its == s["its"]
    assert np.array_equal(ocp._buoy_vector_to_host(ocp.d_mask), s["mask"])
    assert np.abs(ocp._to_reference_layout(ocp.d_x) - s["x"]).max() < 1e-10
    assert H.rel(ocp.d_z.cpu().numpy(), s["z"]) < 1e-9
    ocp.close()


def test_solve_adjoint_ode_host_entry_point_matches_oracle():
    """ocp_solve_adjoint_ode_host = the reference's solve_adjoint_ode(wSol, grad_u, x, buoy_mask, u_values_array)
    with host (K,nt,2) arrays (OCP_dolfin.py:234-252), masked buoys skipped."""
    V = H.lshape()
    O, B = FEOracle(V, 1.0), BuoyOracle(V)
    w = O.newton_solve(initial_control(V, "PL"))
    g = O.project_gradient(w)
    rng = np.random.default_rng(5)
    K = 300
    x0 = np.stack([rng.uniform(-0.05, 2.05, K), rng.uniform(-0.05, 2.05, K)], 1)
    ud = 0.05 * rng.standard_normal((K, 200, 2))
    center = np.array([1.0, 0.5])
    xo, uo, co, mo, po = B.forward(V.velocity_nodal(w), x0, 200, H.H, center)
    assert 0 < mo.sum() < K
    muo = B.adjoint(g, xo, uo, ud, mo, H.H)
    ocp = OCP(V, Parameters(), x0, ud, device=dev())
    mu = ocp.ctx.solve_adjoint_ode_host(g, xo, uo, ud, mo)
    assert mu.shape == (K, 200, 2)
    assert H.rel(mu, muo) < 1e-12
    assert np.all(mu[mo != 0] == 0)
    ocp.close()


def _nccl_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, H.ROOT)
    sys.path.insert(0, os.path.join(H.ROOT, "tests"))
    import torch.distributed as dist
    from ocp_b200.sharding import shard_buoys
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        V = H.square32()
        xr, ud = H.traj(400)
        x0, ud_loc = shard_buoys(xr[:, 0, :].copy(), ud, rank, world)
        ocp = OCP(V, Parameters(), x0, ud_loc, device=torch.device("cuda", rank), group=dist.group.WORLD)
        assert ocp.ctx.comm_size() == world                      # the library owns an NCCL communicator
        f = initial_control(V, "PL")
        ocp.set_control(f)
        ocp.gradient_step(ocp.d_f)
        acc = ocp.d_acc.cpu().numpy()
        z = ocp.d_z.cpu().numpy()
        # the C-ABI host call with the collective inside
        ocp.ctx.set_observations_host(x0, ud_loc)
        w2, z2, mask2, sc = ocp.ctx.gradient_host(f)
        q.put((rank, acc, z, z2, sc["misfit"], sc["n_masked"]))
        ocp.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL refuses two ranks on one device)")
def test_nccl_allreduce_inside_the_library_two_gpus():
    """SURVEY 8(e) on hardware: 400 buoys sharded over two GPUs, the accumulator summed by ncclAllReduce INSIDE
    libocp_b200 (ocp_comm_init / ocp_allreduce), against the single-GPU run and the oracle."""
    import torch.multiprocessing as mp
    V = H.square32()
    xr, ud = H.traj(400)
    x0 = xr[:, 0, :].copy()
    f = initial_control(V, "PL")
    ocp = OCP(V, Parameters(), x0, ud, device=dev())
    ocp.set_control(f)
    ocp.gradient_step(ocp.d_f)
    acc1, z1 = ocp.d_acc.cpu().numpy(), ocp.d_z.cpu().numpy()
    ocp.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        r = q.get(timeout=600)
        got[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    nn2 = 2 * V.num_nodes
    for rank in (0, 1):
        acc, z, z2, misfit, nmask = got[rank]
        assert H.rel(acc[:nn2], acc1[:nn2]) < 1e-12
        assert abs(acc[nn2] - acc1[nn2]) <= 1e-12 * abs(acc1[nn2]) and acc[nn2 + 1] == acc1[nn2 + 1]
        assert H.rel(z, z1) < 1e-10 and H.rel(z2, z1) < 1e-10
        assert abs(misfit - acc1[nn2]) <= 1e-12 * abs(acc1[nn2]) and nmask == int(acc1[nn2 + 1])
    assert np.array_equal(got[0][0], got[1][0])                 # identical on both ranks after the all-reduce
    s = H.OraclePipeline(V, 1.0, x0, ud, 1e-6 * 400).gradient_step(f)
    assert H.rel(got[0][0][:nn2].reshape(-1, 2), s["bnode"]) < 1e-11 and H.rel(got[0][1], s["z"]) < 1e-9


def _sweep_both_ways(V, x0, ud, w, g, center, deterministic=False):
    """forward + backward sweep; returns host copies of everything the sweeps produce"""
    K = x0.shape[0]
    ocp = OCP(V, Parameters(), x0, ud, device=dev())
    if deterministic:
        ocp.ctx.set_deterministic(True)
    cell = torch.empty((200, K), device=dev(), dtype=torch.int32)
    ocp._primal(T(w), ocp.d_x, ocp.d_u, ocp.d_mask, d_cell=cell)
    nn = V.num_nodes
    acc = torch.zeros(2 * nn + 2, device=dev(), dtype=torch.float64)
    mu = torch.empty_like(ocp.d_x)
    ocp.ctx.buoy_adjoint_scatter(ocp.d_vel, T(g), K, ocp.d_x, ocp.d_u, ocp.d_ud, ocp.d_mask, ocp.d_parked, mu, acc)
    out = dict(x=ocp._to_reference_layout(ocp.d_x), u=ocp._to_reference_layout(ocp.d_u),
               cell=ocp._cells_to_host(cell), mask=ocp._buoy_vector_to_host(ocp.d_mask),
               parked=ocp._buoy_vector_to_host(ocp.d_parked), mu=ocp._to_reference_layout(mu), acc=acc.cpu().numpy(),
               staged=ocp.ctx.option("buoy_staged"))
    ocp.close()
    return out


@pytest.mark.parametrize("mesh", ["square32", "lshape"])
def test_staged_shared_memory_tables_equal_the_global_table_path(mesh, monkeypatch):
    """The TMA-staged kernels (mesh tables in shared memory, persistent grid) and the global-table kernels (per-cell
    records through L1/L2) run the same arithmetic: cell indices, trajectories, masks and mu bit-identical, the
    atomically summed b to round-off.  Buoys that leave the domain and the warp-cooperative bin search included."""
    V = H.square32() if mesh == "square32" else H.lshape()
    center = H.CENTER if mesh == "square32" else np.array([1.0, 0.5])
    O = FEOracle(V, 1.0)
    w = O.newton_solve(initial_control(V, "PL")) if mesh == "lshape" else H.field_for(100)
    g = O.project_gradient(w)
    rng = np.random.default_rng(21)
    K = 5000
    x0 = np.stack([rng.uniform(-0.03, 2.03, K), rng.uniform(-0.03, 2.03, K)], 1)
    x0[:64] = V.mesh.coords[rng.choice(V.mesh.num_vertices, 64)]          # exactly on vertices: the tie rule
    ud = 0.05 * rng.standard_normal((K, 200, 2))
    monkeypatch.setenv("OCP_BUOY_STAGED", "1")
    a = _sweep_both_ways(V, x0, ud, w, g, center)
    monkeypatch.delenv("OCP_BUOY_STAGED")
    b = _sweep_both_ways(V, x0, ud, w, g, center)
    assert a["staged"] == 1 and b["staged"] == 0
    for k in ("x", "u", "cell", "mask", "parked"):
        assert np.array_equal(a[k], b[k]), k
    assert H.rel(a["mu"], b["mu"]) < 1e-13        # (the small-launch global path composes mu chunk-wise: same to round-off)
    nn2 = 2 * V.num_nodes
    assert H.rel(a["acc"][:nn2], b["acc"][:nn2]) < 1e-13 and a["acc"][nn2 + 1] == b["acc"][nn2 + 1]
    assert abs(a["acc"][nn2] - b["acc"][nn2]) <= 1e-13 * abs(b["acc"][nn2])
    # and both against the oracle
    B = BuoyOracle(V, brute=True)
    xo, uo, co, mo, po = B.forward(V.velocity_nodal(w), x0, 200, H.H, center)
    assert np.array_equal(a["x"], xo) and np.array_equal(a["u"], uo) and np.array_equal(a["cell"], co)
    assert np.array_equal(a["mask"], mo) and np.array_equal(a["parked"], po)
    muo = B.adjoint(g, xo, uo, ud, mo, H.H)
    bo = B.point_sources(V.velocity_nodal(w), xo, ud, muo, mo, H.H, center)
    assert H.rel(a["mu"], muo) < 1e-12 and H.rel(a["acc"][:nn2].reshape(-1, 2), bo) < 1e-12


@pytest.mark.parametrize("K,staged", [(3000, False), (50_000, False), (3000, True), (50_000, True)])
def test_deterministic_deposit_is_bit_reproducible(K, staged, monkeypatch):
    """SURVEY section 5 (determinism): with ocp_set_deterministic the point sources are accumulated as exact
    fixed-point digit sums with integer atomics - two runs give bit-identical b (also through the per-SM private
    copies at K >= 8 cells' worth), and b agrees with the fp64-atomic variant and the oracle to 1e-12."""
    if staged:
        monkeypatch.setenv("OCP_BUOY_STAGED", "1")
    V = H.square32()
    O = FEOracle(V, 1.0)
    w = H.field_for(100)
    g = O.project_gradient(w)
    rng = np.random.default_rng(4)
    x0 = np.stack([rng.uniform(0.05, 1.95, K), rng.uniform(0.05, 1.95, K)], 1)
    ud = 0.05 * rng.standard_normal((K, 200, 2))
    r1 = _sweep_both_ways(V, x0, ud, w, g, H.CENTER, deterministic=True)
    r2 = _sweep_both_ways(V, x0, ud, w, g, H.CENTER, deterministic=True)
    at = _sweep_both_ways(V, x0, ud, w, g, H.CENTER, deterministic=False)
    nn2 = 2 * V.num_nodes
    assert np.array_equal(r1["acc"], r2["acc"])                              # bit-identical, every entry
    assert H.rel(r1["acc"][:nn2], at["acc"][:nn2]) < 1e-12
    assert r1["acc"][nn2] == at["acc"][nn2] and r1["acc"][nn2 + 1] == at["acc"][nn2 + 1]
    B = BuoyOracle(V)
    xo, uo, co, mo, po = B.forward(V.velocity_nodal(w), x0, 200, H.H, H.CENTER)
    muo = B.adjoint(g, xo, uo, ud, mo, H.H)
    bo = B.point_sources(V.velocity_nodal(w), xo, ud, muo, mo, H.H, H.CENTER)
    assert H.rel(r1["acc"][:nn2].reshape(-1, 2), bo) < 1e-12


def _assemble_all(V, w, f, nu):
    ocp = OCP(V, Parameters(viscosity=nu), np.array([[1.0, 0.5]]), np.zeros((1, 200, 2)), device=dev())
    nnz = V.csr_col.size
    out = {"gather": ocp.ctx.option("gather_assembly")}
    for name, bc in (("J", False), ("Jbc", True)):
        vals = torch.full((nnz,), 7.0, device=dev(), dtype=torch.float64)      # stale content must be overwritten
        res = torch.full((V.ndofs,), -3.0, device=dev(), dtype=torch.float64)
        ocp.ctx.assemble_forward(T(w), T(f), vals, res, bc)
        out[name], out["R" + name] = vals.cpu().numpy(), res.cpu().numpy()
    vals = torch.full((nnz,), 7.0, device=dev(), dtype=torch.float64)
    ocp.ctx.assemble_adjoint(T(w), vals, True)
    out["A"] = vals.cpu().numpy()
    ocp.close()
    return out


@pytest.mark.parametrize("mesh", ["square32", "lshape", "square48"])
def test_atomic_free_assembly_matches_oracle_and_is_bit_reproducible(mesh, monkeypatch):
    """The default assembly gathers by CSR row (no fp64 atomics, fixed summation order): 1e-12 against the oracle like
    the atomic scatter kernels (OCP_ASSEMBLY=atomic), and bit-identical from run to run."""
    from ocp_b200.fespace import TaylorHood
    from ocp_b200.mesh import square_mesh
    V = {"square32": H.square32, "lshape": H.lshape, "square48": lambda: TaylorHood(square_mesh(48))}[mesh]()
    nu = 0.3
    O = FEOracle(V, nu)
    w = 0.3 * np.random.default_rng(8).standard_normal(V.ndofs)
    f = initial_control(V, "OCP")
    a = _assemble_all(V, w, f, nu)
    b = _assemble_all(V, w, f, nu)
    monkeypatch.setenv("OCP_ASSEMBLY", "atomic")
    c = _assemble_all(V, w, f, nu)
    assert a["gather"] == 1 and c["gather"] == 0
    for k in ("J", "Jbc", "A", "RJ", "RJbc"):
        assert np.array_equal(a[k], b[k]), k                        # bit-reproducible
        assert H.rel(a[k], c[k]) < 1e-12, k                         # equals the atomic variant to round-off
    assert H.rel(a["J"], O.on_pattern(O.jacobian_unconstrained(w))) < 1e-12
    assert H.rel(a["Jbc"], O.on_pattern(O.forward_jacobian(w))) < 1e-12
    assert H.rel(a["A"], O.on_pattern(O.adjoint_matrix(w))) < 1e-12
    assert H.rel(a["RJ"], O.forward_residual(w, f)) < 1e-12


def test_step_graph_replay_equals_the_plain_path_and_falls_back():
    """ocp_gradient_device replays one CUDA graph per gradient evaluation once warm.  The replay must give what the plain
    (launch by launch, host-decided) path gives, for a changing control too; when the expected Newton count does not
    hold (a much larger control needs more iterates) the evaluation is re-run on the plain path with the same result
    as a context that never used graphs."""
    V = H.square32()
    xr, ud = H.traj(100)
    x0 = xr[:, 0, :].copy()
    f = initial_control(V, "PL")
    controls = [f, 1.05 * f, 0.9 * f, 40.0 * f, f]          # 40 f: convection dominated, more Newton iterates
    out = {}
    for graph in (True, False):
        os.environ["OCP_STEP_GRAPH"] = "1" if graph else "0"
        try:
            ocp = OCP(V, Parameters(), x0, ud, device=dev())
        finally:
            del os.environ["OCP_STEP_GRAPH"]
        rows = []
        for fc in controls:
            ocp.set_control(fc)
            ocp.gradient_step(ocp.d_f)
            rows.append((ocp.last_newton_its, list(ocp.last_res_hist), ocp.d_w.cpu().numpy(), ocp.d_z.cpu().numpy(),
                         ocp.d_acc.cpu().numpy(), ocp.d_grad.cpu().numpy()))
        out[graph] = (rows, ocp.ctx.option("step_graph_replays"), ocp.ctx.option("step_plain_runs"))
        ocp.close()
    (rg, n_replay, n_plain_g), (rp, n_replay_p, n_plain_p) = out[True], out[False]
    assert n_replay_p == 0 and n_plain_p == len(controls)
    assert n_replay >= 2 and n_plain_g >= 2                      # first call + the fall-back(s) ran plain
    its = [r[0] for r in rp]
    assert its[3] > its[0] and its[4] == its[0]
    nn2 = 2 * V.num_nodes
    for a, b in zip(rg, rp):
        assert a[0] == b[0]                                       # Newton counts
        assert np.allclose(a[1][:-1], b[1][:-1], rtol=1e-6) and len(a[1]) == len(b[1])
        assert H.rel(a[2], b[2]) < 1e-12 and H.rel(a[3], b[3]) < 1e-9
        assert H.rel(a[4][:nn2], b[4][:nn2]) < 1e-11 and H.rel(a[5], b[5]) < 1e-9
