"""Checkpoint / resume in the reference's XDMF+HDF5 layout.

Reference: the control is written every iteration and at the end (``write_checkpoint(project(f, W.sub(0).collapse()),
"f", 0, append=True)``, OCP_dolfin.py:440-441, 485-486), the final state as ``u`` / ``p`` (OCP_dolfin.py:578-588), and a
run is resumed with ``read_checkpoint(f, "f")`` (``load_q`` / ``checkpoints``, OCP_dolfin.py:151-160).
The control is a P2 field, so projecting it onto ``W.sub(0).collapse()`` is the identity on its nodal values.
"""
from __future__ import annotations

import os

import numpy as np

from . import h5lite
from .fespace import TaylorHood


def _collapsed_p2_dofs(V: TaylorHood):
    """Numbering of ``W.sub(0).collapse()`` used for files written here: node-interleaved (2n, 2n+1)."""
    cn = V.cell_nodes.astype(np.int64)
    return np.hstack([2 * cn, 2 * cn + 1]).astype(np.int32)           # (nc, 12) [u_x(6), u_y(6)]


def write_control(path_h5: str, V: TaylorHood, f_nodal: np.ndarray, mtime: int = 0, append: bool = False) -> str:
    """``q.xdmf`` / ``q.h5`` with the function named ``f`` (OCP_dolfin.py:440-441, 485-486).  ``append=True`` adds a
    group ``f_k`` per call like the reference's per-iteration ``checkpoints/q.xdmf``; a non-finite control is
    refused so that a diverged iteration cannot replace the last good resume point."""
    vec = np.ascontiguousarray(f_nodal, np.float64).reshape(-1)         # (nn,2) row-major = interleaved numbering
    if not np.all(np.isfinite(vec)):
        raise ValueError("write_control: the control holds non-finite values; checkpoint left untouched")
    return h5lite.write_checkpoint(path_h5, "f", V.mesh.cells, V.mesh.coords, _collapsed_p2_dofs(V), vec, 12, mtime,
                                   append=append)


def read_control(path_h5: str, V: TaylorHood, counter: int = -1) -> np.ndarray:
    """Control checkpoint -> P2 nodal field (nn,2), for any dof numbering of the collapsed space (the file carries its
    own ``cell_dofs``), e.g. reference_runs/u_bar_chapter_6.3.3/q_backup/q.h5.  ``counter`` = -1 reads the LAST
    appended group, which is what dolfin's ``read_checkpoint(f, "f")`` returns (OCP_dolfin.py:151-160)."""
    d = h5lite.read_checkpoint(path_h5, "f", counter)
    if not np.array_equal(d["topology"], V.mesh.cells) or not np.allclose(d["geometry"], V.mesh.coords, atol=1e-14):
        raise ValueError("checkpoint mesh differs from the space's mesh")
    cd = d["cell_dofs"].reshape(-1, 12)
    f = np.zeros((V.num_nodes, 2))
    f[V.cell_nodes, 0] = d["vector"][cd[:, :6]]
    f[V.cell_nodes, 1] = d["vector"][cd[:, 6:]]
    return f


def write_state(dir_path: str, V: TaylorHood, w: np.ndarray, mtime: int = 0):
    """``u.xdmf`` and ``p.xdmf`` of the final state (OCP_dolfin.py:578-588): both carry the full mixed vector and the
    cell dofs of their sub-space, exactly like dolfin's checkpoints of ``w.split()``."""
    os.makedirs(dir_path, exist_ok=True)
    cd = V.cell_dofs
    xu = h5lite.write_checkpoint(os.path.join(dir_path, "u.h5"), "u", V.mesh.cells, V.mesh.coords, cd[:, :12], w, 12, mtime)
    xp = h5lite.write_checkpoint(os.path.join(dir_path, "p.h5"), "p", V.mesh.cells, V.mesh.coords, cd[:, 12:], w, 3, mtime,
                                 element_degree=1, value_rank=0)
    return xu, xp


def read_state(path_h5: str, V: TaylorHood, name: str = "u") -> np.ndarray:
    """W vector from a ``u`` / ``p`` checkpoint written with the same dof numbering (dolfin's for the 32x32 square)."""
    d = h5lite.read_checkpoint(path_h5, name)
    if not np.array_equal(d["topology"], V.mesh.cells):
        raise ValueError("checkpoint mesh differs from the space's mesh")
    ncol = 12 if name == "u" else 3
    mine = V.cell_dofs[:, :12] if name == "u" else V.cell_dofs[:, 12:]
    if not np.array_equal(d["cell_dofs"].reshape(-1, ncol), mine):
        raise ValueError("checkpoint dof numbering differs from the space's numbering")
    return d["vector"].copy()
