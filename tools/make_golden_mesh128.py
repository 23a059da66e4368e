#!/usr/bin/env python
"""Oracle results of ONE gradient evaluation on the refined 128 x 128 square mesh (cfg5 mesh, 148 739 dofs).

The NumPy/SuperLU oracle needs about five minutes for this on one host core, which is too long to repeat inside
every `pytest -m gpu` run on the GPU box, so its results are computed once here and committed as a small fixture:

    python tools/make_golden_mesh128.py        ->  tests/golden/mesh128_gradient.npz   (~0.3 MB)

The fixture holds every 16th entry of w, z and g (plus their full 2-norms), the full Gamma_1 trace of z and of the
gradient, the full nodal point-source vector's non-zeros, the cost and the Newton history.  Inputs are regenerated
by the test from the same seeds (start points, twin observations from the control 1.3 f).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ocp_b200  # noqa: E402,F401
from ocp_b200.fespace import TaylorHood  # noqa: E402
from ocp_b200.mesh import square_mesh  # noqa: E402
from ocp_b200.pipeline import initial_control  # noqa: E402
import helpers as H  # noqa: E402

N, K, STRIDE = 128, 64, 16


def problem():
    """Shared by this script and tests/test_gpu_parity_sizes.py."""
    V = TaylorHood(square_mesh(N))
    rng = np.random.default_rng(3)
    x0 = np.stack([rng.uniform(0.2, 1.0, K), rng.uniform(0.3, 1.7, K)], 1)
    return V, x0, initial_control(V, "PL")


def main():
    V, x0, f = problem()
    P = H.OraclePipeline(V, 1.0, x0, np.zeros((K, 200, 2)), 1e-6 * K)
    t0 = time.time()
    w_twin = P.O.newton_solve(1.3 * f)
    _, ud, _, mask, _ = P.primal(w_twin)
    assert mask.sum() == 0
    P.ud = ud
    s = P.gradient_step(f)
    print(f"oracle: {time.time() - t0:.0f} s, newton its {s['its']}")
    g1 = np.unique(V.g1_nodes)
    nz = np.flatnonzero(np.abs(s["bnode"]).sum(1) > 0)
    out = os.path.join(ROOT, "tests", "golden", "mesh128_gradient.npz")
    np.savez_compressed(
        out, stride=STRIDE, its=s["its"], hist=np.array(s["hist"]),
        w_sample=s["w"][::STRIDE], z_sample=s["z"][::STRIDE], g_sample=s["g"][::STRIDE],
        w_norm=np.linalg.norm(s["w"]), z_norm=np.linalg.norm(s["z"]), g_norm=np.linalg.norm(s["g"]),
        z_g1=V.velocity_nodal(s["z"])[g1], grad_g1=s["grad"][g1], b_nodes=nz, b_vals=s["bnode"][nz],
        ud=ud.astype(np.float64), J=P.cost(s["u"], f), x_last=s["x"][:, -1, :])
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
