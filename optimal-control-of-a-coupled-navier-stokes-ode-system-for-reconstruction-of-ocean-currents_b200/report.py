"""Plot / report layer of a run - the figures OCP_dolfin.py:455-575 writes next to its text outputs.

File names follow the reference: ``mesh.png`` (OCP_dolfin.py:455-475), ``J.png`` (515-521), ``buoy_movements/frames/
buoy_movement_<k>.png`` (531-552), ``ud_plot_buoy_<k>.png`` (555-567), ``u_field.png`` (570-575).  The reference draws
them with matplotlib + dolfin's ``plot``; matplotlib is an optional dependency here (absent from the build image): with
it the PNGs are produced from host copies of the device arrays, without it only the dependency-free ``J.svg`` is
written and the skipped files are reported.  Nothing in this module touches the hot path.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np


def _have_matplotlib() -> bool:
    try:
        import matplotlib  # noqa: F401
        return True
    except Exception:
        return False


def write_cost_svg(path: str, J: Sequence[float], title: str = "Reduced cost j(q)") -> str:
    """Cost curve of OCP_dolfin.py:515-521 as a stand-alone SVG (no plotting library needed)."""
    J = np.asarray(list(J), float)
    W, Hh, m = 640, 400, 55
    n = max(len(J), 2)
    lo, hi = (float(np.nanmin(J)), float(np.nanmax(J))) if len(J) else (0.0, 1.0)
    if not np.isfinite(lo) or not np.isfinite(hi) or hi == lo:
        lo, hi = lo - 0.5, lo + 0.5
    px = lambda i: m + (W - 2 * m) * i / (n - 1)
    py = lambda v: Hh - m - (Hh - 2 * m) * (v - lo) / (hi - lo)
    pts = " ".join(f"{px(i):.1f},{py(v):.1f}" for i, v in enumerate(J) if np.isfinite(v))
    ticks = "".join(
        f'<text x="{m - 6}" y="{py(v) + 4:.1f}" font-size="11" text-anchor="end">{v:.4g}</text>'
        f'<line x1="{m}" x2="{W - m}" y1="{py(v):.1f}" y2="{py(v):.1f}" stroke="#ddd"/>'
        for v in np.linspace(lo, hi, 5))
    xt = "".join(f'<text x="{px(i):.1f}" y="{Hh - m + 16}" font-size="11" text-anchor="middle">{i}</text>'
                 for i in range(0, n, max(1, n // 10)))
    svg = (f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{Hh}" viewBox="0 0 {W} {Hh}">'
           f'<rect width="{W}" height="{Hh}" fill="white"/>{ticks}{xt}'
           f'<rect x="{m}" y="{m}" width="{W - 2 * m}" height="{Hh - 2 * m}" fill="none" stroke="black"/>'
           f'<polyline points="{pts}" fill="none" stroke="black" stroke-width="1.5"/>'
           f'<text x="{W / 2}" y="24" font-size="15" text-anchor="middle">{title}</text>'
           f'<text x="{W / 2}" y="{Hh - 8}" font-size="12" text-anchor="middle">Iteration</text>'
           f'<text x="14" y="{Hh / 2}" font-size="12" text-anchor="middle" transform="rotate(-90 14 {Hh / 2})">Cost</text>'
           "</svg>\n")
    with open(path, "w") as fh:
        fh.write(svg)
    return path


def _boundary_segments(V):
    """(segments of Gamma_1, segments of the rest of the boundary) as (2,2) arrays - `mesh_boundary` of the reference."""
    m = V.mesh
    on_bnd = np.flatnonzero(m.edge_cells[:, 1] < 0)
    g1 = set(int(e) for e in V.marking.gamma1)
    seg = lambda e: m.coords[m.edges[e]]
    return [seg(e) for e in on_bnd if int(e) in g1], [seg(e) for e in on_bnd if int(e) not in g1]


def save_plots(ocp, res, out_dir: str, x: Optional[np.ndarray] = None, u_values: Optional[np.ndarray] = None,
               max_buoy_plots: int = 12) -> dict:
    """Write the figures of a finished run.  ``ocp`` is the ``pipeline.OCP`` that produced ``res`` (a ``RunResult``);
    ``x`` / ``u_values`` (K,nt,2) default to the trajectories of the last gradient evaluation.  Returns
    ``{"written": [...], "skipped": [...]}``."""
    os.makedirs(out_dir, exist_ok=True)
    written: List[str] = [write_cost_svg(os.path.join(out_dir, "J.svg"), res.J_array)]
    names = ["mesh.png", "J.png", "u_field.png", "buoy_movements/frames/buoy_movement_0.png"] + \
            [f"ud_plot_buoy_{k}.png" for k in range(min(ocp.K, max_buoy_plots))]
    if not _have_matplotlib():
        return {"written": written, "skipped": names, "reason": "matplotlib is not installed"}
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    from matplotlib.tri import Triangulation
    V = ocp.V
    if x is None:
        x = ocp._to_reference_layout(ocp.d_x)
    if u_values is None:
        u_values = ocp._to_reference_layout(ocp.d_u)
    g1, rest = _boundary_segments(V)
    tri = Triangulation(V.mesh.coords[:, 0], V.mesh.coords[:, 1], V.mesh.cells)

    def boundary(ax, gamma1_color="orange", rest_color="blue"):
        for k, s in enumerate(rest):
            ax.plot(s[:, 0], s[:, 1], color=rest_color, label=r"$\\Gamma_2$" if k == 0 else None)
        for k, s in enumerate(g1):
            ax.plot(s[:, 0], s[:, 1], color=gamma1_color, label=r"$\\Gamma_1$" if k == 0 else None)

    # mesh.png (OCP_dolfin.py:455-475)
    fig, ax = plt.subplots()
    ax.triplot(tri, color="gray", linewidth=0.3)
    boundary(ax)
    ax.set_title(r"discretized domain $\\Omega_h$"), ax.set_xlabel(r"$x$"), ax.set_ylabel(r"$y$")
    ax.set_aspect("equal")
    ax.legend(loc="best", bbox_to_anchor=(1.02, 1))
    fig.savefig(os.path.join(out_dir, "mesh.png"), bbox_inches="tight"), plt.close(fig)
    # J.png (OCP_dolfin.py:515-521)
    fig, ax = plt.subplots()
    ax.plot(res.J_array, color="black")
    ax.set_xlabel("Iteration"), ax.set_ylabel("Cost"), ax.set_title(r"Reduced cost $j(q)$")
    fig.savefig(os.path.join(out_dir, "J.png")), plt.close(fig)
    # u_field.png (OCP_dolfin.py:570-575): magnitude on the vertices + arrows
    vel = V.velocity_nodal(ocp.d_w.cpu().numpy())[:V.mesh.num_vertices]
    fig, ax = plt.subplots()
    c = ax.tripcolor(tri, np.hypot(vel[:, 0], vel[:, 1]), shading="gouraud")
    ax.quiver(V.mesh.coords[:, 0], V.mesh.coords[:, 1], vel[:, 0], vel[:, 1], color="white", width=0.002)
    fig.colorbar(c), ax.set_title(r"Velocity field $u$"), ax.set_xlabel(r"$x$"), ax.set_ylabel(r"$y$")
    ax.set_aspect("equal")
    fig.savefig(os.path.join(out_dir, "u_field.png")), plt.close(fig)
    # buoy movement (OCP_dolfin.py:531-552) - first max_buoy_plots buoys of the last iteration
    os.makedirs(os.path.join(out_dir, "buoy_movements", "frames"), exist_ok=True)
    fig, ax = plt.subplots()
    for i in range(min(ocp.K, max_buoy_plots)):
        ax.scatter(ocp.xsarr[i], ocp.ysarr[i], color="red", zorder=5)
        ax.plot(x[i, :, 0], x[i, :, 1], color="b", linestyle=(0, (i + 2, (i + 2) // 2)), label=rf"$x_{{{i + 1}}}$")
    for s in g1 + rest:
        ax.plot(s[:, 0], s[:, 1], color="gray")
    ax.set_aspect("equal", adjustable="box"), ax.set_title("Buoy movement result")
    ax.set_xlabel(r"$x$"), ax.set_ylabel(r"$y$"), ax.legend(loc="best", bbox_to_anchor=(1.02, 1))
    fig.savefig(os.path.join(out_dir, "buoy_movements", "frames", "buoy_movement_0.png"), bbox_inches="tight"), plt.close(fig)
    # velocity comparison per buoy (OCP_dolfin.py:555-567); time axis = linspace(t0, T, nt) as in the reference
    t = np.linspace(ocp.params.t0, ocp.params.T, ocp.nt)
    for k in range(min(ocp.K, max_buoy_plots)):
        fig, ax = plt.subplots()
        if ocp.u_d is not None:
            ax.plot(t, ocp.u_d[k, :, 0], color="black", alpha=0.8, label=r"$u_{d,1}$")
            ax.plot(t, ocp.u_d[k, :, 1], color="black", alpha=0.8, label=r"$u_{d,2}$")
        ax.plot(t, u_values[k, :, 0], color="b", linestyle=(0, (k + 2, (k + 2) // 2)), label=r"$u_1$")
        ax.plot(t, u_values[k, :, 1], color="b", linestyle=(0, (k + 2, (k + 2) // 2)), label=r"$u_2$")
        ax.set_title(rf"Velocity comparison for buoy k={k + 1}"), ax.set_xlabel("Time"), ax.set_ylabel("Velocity")
        ax.legend(loc="best")
        fig.savefig(os.path.join(out_dir, f"ud_plot_buoy_{k}.png")), plt.close(fig)
    return {"written": written + [os.path.join(out_dir, n) for n in names], "skipped": []}
