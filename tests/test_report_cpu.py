"""Plot / report layer (OCP_dolfin.py:455-575): file names, the dependency-free cost curve, matplotlib gating."""
import os
from types import SimpleNamespace

import numpy as np

import helpers as H
from ocp_b200 import report


def test_cost_svg_and_matplotlib_gate(tmp_path):
    J = [0.544, 0.434, 0.362, 0.311, 0.272]
    p = report.write_cost_svg(str(tmp_path / "J.svg"), J)
    txt = open(p).read()
    assert txt.startswith("<svg") and "polyline" in txt and "Reduced cost" in txt and txt.count(",") >= len(J)
    V = H.square32()
    g1, rest = report._boundary_segments(V)
    assert len(g1) == 64 and len(rest) == 64                       # Gamma_1 = the edges x = 0 and x = 2 (OCP_dolfin.py:118-136)
    ocp = SimpleNamespace(K=3, V=V)
    out = report.save_plots(ocp, SimpleNamespace(J_array=J), str(tmp_path)) if not report._have_matplotlib() else None
    if out is not None:                                             # the build image has no matplotlib: PNGs are reported as skipped
        assert os.path.exists(out["written"][0]) and "J.png" in out["skipped"] and "mesh.png" in out["skipped"]
        assert "ud_plot_buoy_2.png" in out["skipped"]
