#!/usr/bin/env python
"""Developer diagnostic: run every device entry point once against the CPU oracle and print the
deviations and timings (does not stop at the first mismatch).  Needs a GPU: run under gpurun."""
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocp_b200  # noqa: E402
from ocp_b200 import capi  # noqa: E402
from ocp_b200.fespace import TaylorHood  # noqa: E402
from ocp_b200.mesh import lshape_mesh, square_mesh  # noqa: E402
from ocp_b200.pipeline import OCP, Knobs, Parameters, initial_control  # noqa: E402
from oracle.buoy_oracle import BuoyOracle  # noqa: E402
from oracle.fe_oracle import FEOracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def section(name):
    print(f"\n=== {name} ===", flush=True)


def timed(fn, n=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    dev = torch.device("cuda:0")
    V = TaylorHood(square_mesh(32))
    F = np.load(os.path.join(GOLD, "fields.npz"))
    P = Parameters()
    O = FEOracle(V, P.viscosity)
    B = BuoyOracle(V)
    tr = np.load(os.path.join(GOLD, "traj_100_buoys.npz"))
    x0, ud = tr["x_0_array"][:, 0, :].copy(), tr["u_d_array"]
    ocp = OCP(V, P, x0, ud, device=dev)
    ctx = ocp.ctx
    ctx.set_profiling(True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    try:
        section("buoy forward vs oracle (100 buoys, twin field)")
        w = F["velocity_100"]
        d_w = t(w)
        cell = torch.empty((ocp.nt, ocp.K), device=dev, dtype=torch.int32)
        ocp._primal(d_w, ocp.d_x, ocp.d_u, ocp.d_mask, d_cell=cell)
        x = ocp._to_reference_layout(ocp.d_x)
        u = ocp._to_reference_layout(ocp.d_u)
        xo, uo, co, mo, po = B.forward(V.velocity_nodal(w), x0, ocp.nt, P.dt, ocp.center_of_domain)
        print("x bit-exact:", np.array_equal(x, xo), "u bit-exact:", np.array_equal(u, uo),
              "cells equal:", np.array_equal(ocp._cells_to_host(cell), co), "max|x-xref|", np.abs(x - tr["x_0_array"]).max())
        print("max |x-xo|", np.abs(x - xo).max(), "max|u-uo|", np.abs(u - uo).max())
    except Exception:
        traceback.print_exc()

    try:
        section("assembly vs oracle")
        rng = np.random.default_rng(0)
        wr = 0.3 * rng.standard_normal(V.ndofs)
        f = initial_control(V, "OCP")
        d_wr, d_f = t(wr), t(f)
        vals = torch.zeros(V.csr_col.size, device=dev, dtype=torch.float64)
        res = torch.zeros(V.ndofs, device=dev, dtype=torch.float64)
        ctx.assemble_forward(d_wr, d_f, vals, res, False)
        Jo = O.on_pattern(O.jacobian_unconstrained(wr))
        Ro = O.forward_residual(wr, f)
        print("J unconstrained rel err", rel(vals.cpu().numpy(), Jo), "residual rel err", rel(res.cpu().numpy(), Ro))
        ctx.assemble_forward(d_wr, d_f, vals, res, True)
        Jb = O.on_pattern(O.forward_jacobian(wr))
        print("J with BC rel err", rel(vals.cpu().numpy(), Jb))
        ctx.assemble_adjoint(d_wr, vals, True)
        Ao = O.on_pattern(O.adjoint_matrix(wr))
        print("adjoint matrix rel err", rel(vals.cpu().numpy(), Ao))
        print("assemble fwd (J+F) ms", timed(lambda: ctx.assemble_forward(d_wr, d_f, vals, res, True)))
    except Exception:
        traceback.print_exc()

    try:
        section("Newton vs oracle / fixture")
        f = initial_control(V, "PL")
        d_f = t(f)
        st = ocp.forward_solve(d_f)
        wo, ito, hist = O.newton_solve(f, return_history=True)
        print("its", ocp.last_newton_its, ito, "hist", ocp.last_res_hist, hist)
        print("w rel err vs oracle", rel(st.vector(), wo))
        print("solver stats", ctx.solver_stats())
        ctx.reset_solver_stats()
        ms = timed(lambda: ocp.forward_solve(d_f), 3)
        print("forward_solve ms", ms, ctx.solver_stats())
    except Exception:
        traceback.print_exc()

    try:
        section("projection / adjoint chain vs oracle (u_bar, 6 buoys: KAT K4/K5)")
        tr6 = np.load(os.path.join(GOLD, "traj_6_buoys.npz"))
        ocp6 = OCP(V, P, tr6["x_0_array"][:, 0, :].copy(), tr6["u_d_array"], device=dev)
        ubar = F["u_bar"]
        d_ub = t(ubar)
        g = ocp6.project_grad(type(st)(d_ub))
        go = O.project_gradient(ubar)
        print("grad proj rel err", rel(g.cpu().numpy(), go))
        mask = np.zeros(6)
        x, u = ocp6.solve_primal_ode(type(st)(d_ub), mask)
        xo, uo, co, mo, po = B.forward(V.velocity_nodal(ubar), ocp6.d_x0.cpu().numpy(), 200, P.dt, [1.0, 1.0])
        print("traj bit-exact", np.array_equal(x, xo), np.array_equal(u, uo))
        mu = ocp6.solve_adjoint_ode(type(st)(d_ub), g, x, mask, u)
        muo = B.adjoint(go, xo, uo, tr6["u_d_array"], mo, P.dt)
        print("mu rel err", rel(mu, muo))
        z = ocp6.adjoint_solve(type(st)(d_ub), x, u, mask, g)
        bo = B.point_sources(V.velocity_nodal(ubar), xo, tr6["u_d_array"], muo, mo, P.dt, [1.0, 1.0])
        nn = V.num_nodes
        print("b rel err", rel(ocp6.d_acc[:2 * nn].cpu().numpy().reshape(-1, 2), bo),
              "misfit", float(ocp6.d_acc[2 * nn]), B.misfit(uo, tr6["u_d_array"], P.dt))
        zo = O.adjoint_solve(ubar, O.rhs_from_nodal(bo))
        print("z rel err", rel(z.vector(), zo))
        print("norms", ocp6.field_norms(type(st)(d_ub)), O.divergence_norm(ubar), O.l2_h1_norms(ubar))
        ocp6.close()
    except Exception:
        traceback.print_exc()

    try:
        section("GD loop, OCP defaults, 6 buoys (SURVEY B.6: 0.54411128163, 0.43379918900, ...)")
        tr6 = np.load(os.path.join(GOLD, "traj_6_buoys.npz"))
        ocp6 = OCP(V, P, tr6["x_0_array"][:, 0, :].copy(), tr6["u_d_array"], device=dev)
        t0 = time.time()
        r = ocp6.run(initial_control(V, "OCP"), Knobs(num_steps=4, use_line_search=True))
        print("J", r.J_array, "LS its", r.inner_iterations, "newton", r.newton_its, "LR", r.LR, "wall", time.time() - t0)
        print("outer", r.outer_time, "inner", r.inner_time)
        ocp6.close()
        section("PL defaults 10 buoys + grad check (SURVEY K6: J0 0.025045819440590228 gradj -0.02603602981198921)")
        tr10 = np.load(os.path.join(GOLD, "traj_10_buoys.npz"))
        ocp10 = OCP(V, P, tr10["x_0_array"][:, 0, :].copy(), tr10["u_d_array"], device=dev)
        r = ocp10.run(initial_control(V, "PL"), Knobs(num_steps=2, use_line_search=False, grad_check=True, exit_rule="ten"))
        print("J", r.J_array, "J0", r.grad_tables["J0"], "gradj", r.grad_tables["gradj"])
        for row in r.grad_tables["one_sided"]:
            print("  one-sided", row)
        for row in r.grad_tables["centered"]:
            print("  centred  ", row)
        ocp10.close()
    except Exception:
        traceback.print_exc()

    try:
        section("buoy kernels at scale (square N=32 field, synthetic buoys)")
        for K in (10_000, 1_000_000):
            rng = np.random.default_rng(0)
            x0 = np.stack([rng.uniform(0.1, 0.4, K), rng.uniform(0.25, 1.75, K)], 1)
            big = OCP(V, P, x0, np.zeros((K, 200, 2)), device=dev)
            d_w = t(F["velocity_100"])
            big.ctx.project_grad(d_w, big.d_g)
            big._primal(d_w, big.d_x, big.d_u, big.d_mask)
            big.d_ud.copy_(1.1 * big.d_u)
            ms_f = timed(lambda: big._primal(d_w, big.d_x, big.d_u, big.d_mask))

            def back():
                big.d_acc.zero_()
                big.ctx.buoy_adjoint_scatter(big.d_vel, big.d_g, K, big.d_x, big.d_u, big.d_ud, big.d_mask,
                                             big.d_parked, None, big.d_acc)
            ms_b = timed(back)
            steps = K * 200
            print(f"K={K}: forward {ms_f:.3f} ms ({steps / ms_f / 1e6:.1f} G buoy-steps/s, {32 * steps / ms_f / 1e6:.1f} GB/s)"
                  f"  backward {ms_b:.3f} ms ({steps / ms_b / 1e6:.1f} G/s, {48 * steps / ms_b / 1e6:.1f} GB/s)"
                  f" masked {int(big.d_mask.sum())}")
            if K == 10_000:
                big.ctx.reset_solver_stats()
                big.ctx.set_profiling(os.environ.get('PROFILE_STEP', '0') == '1')
                big.set_control(initial_control(V, "PL"))
                ms_it = timed(lambda: big.gradient_step(big.d_f), 3)
                print("full gradient step ms", ms_it, big.ctx.solver_stats())
            big.close()
    except Exception:
        traceback.print_exc()

    try:
        section("L-shape (jittered) mesh: locate parity incl. outside points")
        VL = TaylorHood(lshape_mesh(20, jitter=0.2))
        OL = FEOracle(VL, 1.0)
        fL = initial_control(VL, "PL")
        rng = np.random.default_rng(3)
        K = 500
        x0 = np.stack([rng.uniform(-0.05, 2.05, K), rng.uniform(-0.05, 2.05, K)], 1)
        oc = OCP(VL, P, x0, np.zeros((K, 200, 2)), device=dev)
        st = oc.forward_solve(t(fL))
        wL = OL.newton_solve(fL)
        print("L-shape Newton rel err", rel(st.vector(), wL), "its", oc.last_newton_its)
        mask = np.zeros(K)
        x, u = oc.solve_primal_ode(st, mask)
        BL = BuoyOracle(VL, brute=True)
        xo, uo, co, mo, po = BL.forward(VL.velocity_nodal(st.vector()), x0, 200, P.dt, oc.center_of_domain)
        print("bit-exact x,u:", np.array_equal(x, xo), np.array_equal(u, uo), "mask equal", np.array_equal(mask, mo),
              "masked", int(mask.sum()), "parked", int(po.sum()), int(oc.d_parked.sum()))
        oc.close()
    except Exception:
        traceback.print_exc()
    print("\ndone", flush=True)


if __name__ == "__main__":
    main()
