#!/usr/bin/env python
"""Per-configuration timing table: every reference configuration that fits one GPU
is run through ``OCP.run`` - the loop of OCP_dolfin.py:309-450 - and timed the way the reference times itself
(``outer loop time`` = forward + ODEs + adjoint + gradient, ``inner loop time`` = Armijo line search,
OCP_dolfin.py:313, 374-375, 384, 419-423), next to the four per-iteration times the reference publishes
(plotting/histogram_plotting.py:9-10).

    python tools/config_table.py > profiles/config_table_r1.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocp_b200  # noqa: E402,F401
from ocp_b200.fespace import TaylorHood  # noqa: E402
from ocp_b200.mesh import lshape_mesh, square_mesh  # noqa: E402
from ocp_b200.pipeline import OCP, Knobs, Parameters, initial_control  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
PUBLISHED_S_PER_ITERATION = {10: 0.10, 100: 11.98, 400: 77.82, 10000: 1500.0}
STEPS = int(os.environ.get("STEPS", "6"))
dev = torch.device("cuda", 0)


def traj(K):
    t = np.load(os.path.join(GOLD, f"traj_{K}_buoys.npz"))
    return t["x_0_array"][:, 0, :].copy(), t["u_d_array"]


def timed_run(ocp, f0, knobs):
    ocp.run(f0, Knobs(num_steps=1, use_line_search=knobs.use_line_search, exit_rule=knobs.exit_rule))   # warm-up: graphs, allocations
    r = ocp.run(f0, knobs)
    outer = float(np.median(r.outer_time)) * 1e3
    trials = int(sum(r.inner_iterations))
    inner_s = float(sum(r.inner_time))
    return {"iterations_run": len(r.J_array), "outer_ms_per_iteration": outer, "gd_iters_per_sec_outer_only": 1e3 / outer,
            "buoy_steps_per_sec": 3.0 * ocp.K * ocp.nt / (outer * 1e-3),
            "line_search_trials": trials, "line_search_trials_per_sec": (trials / inner_s) if trials else None,
            "newton_its": r.newton_its, "J_first_last": [r.J_array[0], r.J_array[-1]], "exit": r.exit_reason}


rows = []
V = TaylorHood(square_mesh(32))

# cfg1: OCP_dolfin.py defaults on the square, 10 buoys, Armijo line search
x0, ud = traj(10)
ocp = OCP(V, Parameters(), x0, ud, device=dev)
rows.append({"config": "cfg1 OCP_dolfin.py square 32x32, 10 buoys, line search", "K": 10,
             **timed_run(ocp, initial_control(V, "OCP"), Knobs(num_steps=STEPS, use_line_search=True))})
ocp.close()

# Pipeline_limits.py sweep (no line search, LR = 5, constant initial control): the four published points
grid = np.meshgrid(np.linspace(0.1, 0.4, 100), np.linspace(0.25, 1.75, 100))
for K in (10, 100, 400, 10000):
    if K == 10000:
        x0 = np.stack([grid[0].ravel(), grid[1].ravel()], 1)
        field = torch.from_numpy(np.load(os.path.join(GOLD, "fields.npz"))["velocity_100"]).to(dev)
        tmp = OCP(V, Parameters(), x0, np.zeros((K, 200, 2)), device=dev)
        tmp._primal(field, tmp.d_x, tmp.d_u, tmp.d_mask)
        ud = tmp._to_reference_layout(tmp.d_u)
        tmp.close()
    else:
        x0, ud = traj(K)
    ocp = OCP(V, Parameters(), x0, ud, device=dev)
    # the reference run diverges after a few LR = 5 steps (buoys leave the domain): time the first iterations
    row = timed_run(ocp, initial_control(V, "PL"), Knobs(num_steps=min(STEPS, 3), use_line_search=False, exit_rule="ten"))
    pub = PUBLISHED_S_PER_ITERATION[K]
    rows.append({"config": f"Pipeline_limits.py square 32x32, {K} buoys", "K": K, **row,
                 "published_reference_s_per_iteration": pub, "speedup_vs_published": pub * 1e3 / row["outer_ms_per_iteration"]})
    ocp.close()

# initial_control_test.py cases 0-3 (6 buoys)
x0, ud = traj(6)
for case in range(4):
    ocp = OCP(V, Parameters(), x0, ud, device=dev)
    rows.append({"config": f"cfg4 initial_control_test.py case {case}, 6 buoys", "K": 6,
                 **timed_run(ocp, initial_control(V, "ICT", case), Knobs(num_steps=min(STEPS, 3), use_line_search=False, exit_rule="ten"))})
    ocp.close()

# cfg2: L-shape, 100 synthetic buoys (10 x 10 grid in the lower-left block), twin observations from a GPU forward
# solve with a different control, Armijo line search
VL = TaylorHood(lshape_mesh(16, jitter=0.15))
gx, gy = np.meshgrid(np.linspace(0.15, 0.85, 10), np.linspace(0.15, 0.85, 10))
x0 = np.stack([gx.ravel(), gy.ravel()], 1)
f_true = VL.interpolate_control(lambda x, y: 0.3 * np.sin(np.pi * y) + 0 * x, lambda x, y: -0.2 * np.cos(np.pi * x), 2)
twin = OCP(VL, Parameters(), x0, np.zeros((100, 200, 2)), device=dev)
st = twin.forward_solve(torch.from_numpy(f_true).to(dev))
_, ud = twin.solve_primal_ode(st, np.zeros(100))
twin.close()
ocp = OCP(VL, Parameters(), x0, ud, device=dev)
rows.append({"config": f"cfg2 L-shape ({VL.mesh.num_cells} cells, {VL.ndofs} dofs), 100 buoys, line search", "K": 100,
             **timed_run(ocp, initial_control(VL, "PL"), Knobs(num_steps=STEPS, use_line_search=True))})
ocp.close()

print(json.dumps({"gpu": torch.cuda.get_device_name(0), "note": "outer = median wall time of the gradient block per iteration "
                  "(host clock around a synchronised device, as the reference's timings.txt); published = "
                  "plotting/histogram_plotting.py:9-10, hardware unstated", "rows": rows}, indent=1))
