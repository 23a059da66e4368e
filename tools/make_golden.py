#!/usr/bin/env python
"""Extract the reference's own FEniCS artefacts into small committed fixtures.

Run in the build container (needs the read-only /root/reference mount):

    python tools/make_golden.py

Writes
  <package>/data/dolfin_square32_dofmap.npz   dolfin's dof numbering of W on the 32x32 square
  tests/golden/traj_<K>_buoys.npz             x_0_array / u_d_array of reference_runs/<K>_buoys
  tests/golden/fields.npz                     FE vectors (twin-experiment fields, u_bar state, control q)
  tests/golden/scalars.json                   J_array, norms, divergence values, variables

Nothing here is reference *source*; these are data files produced by the reference's
FEniCS runs (reference_runs/*), which is the only oracle data that exists for this path.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocp_b200  # noqa: E402
from ocp_b200 import h5lite  # noqa: E402

REF = "/root/reference/reference_runs"
PKG = os.path.dirname(ocp_b200.__file__)
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(os.path.join(PKG, "data"), exist_ok=True)
    os.makedirs(GOLD, exist_ok=True)
    ub = os.path.join(REF, "u_bar_chapter_6.3.3")
    u = h5lite.read_checkpoint(os.path.join(ub, "paraview/checkpoint/u.h5"), "u")
    p = h5lite.read_checkpoint(os.path.join(ub, "paraview/checkpoint/p.h5"), "p")
    q = h5lite.read_checkpoint(os.path.join(ub, "q_backup/q.h5"), "f")
    assert np.array_equal(u["vector"], p["vector"])
    np.savez_compressed(
        os.path.join(PKG, "data", "dolfin_square32_dofmap.npz"),
        topology=u["topology"].astype(np.int32),
        geometry=u["geometry"],
        u_cell_dofs=u["cell_dofs"].astype(np.int32),
        p_cell_dofs=p["cell_dofs"].astype(np.int32),
    )

    fields = {"u_bar": u["vector"], "q_vector": q["vector"], "q_cell_dofs": q["cell_dofs"].astype(np.int32)}
    scalars = {}
    for K in (2, 4, 6, 10, 100, 400, 10000):
        d = os.path.join(REF, f"{K}_buoys")
        v = h5lite.read_checkpoint(os.path.join(d, "paraview/velocity.h5"), "u")
        for k in ("topology", "cell_dofs"):
            assert np.array_equal(v[k], u[k]), (K, k)
        fields[f"velocity_{K}"] = v["vector"]
        if K != 10000:
            np.savez_compressed(
                os.path.join(GOLD, f"traj_{K}_buoys.npz"),
                x_0_array=np.load(os.path.join(d, "x_0_array.npy")),
                u_d_array=np.load(os.path.join(d, "u_d_array.npy")),
            )
        sc = {}
        for name in ("norms.txt", "u_divergence.txt", "variables.txt"):
            with open(os.path.join(d, name)) as fh:
                sc[name] = fh.read()
        scalars[f"{K}_buoys"] = sc
    # drop duplicates (100/400/10000 share one field; 2/4/6 share one)
    keep = {}
    alias = {}
    for k, v in fields.items():
        for k2, v2 in keep.items():
            if v.shape == v2.shape and np.array_equal(v, v2):
                alias[k] = k2
                break
        else:
            keep[k] = v
    np.savez_compressed(os.path.join(GOLD, "fields.npz"), **keep)
    sc = {"J_array": np.load(os.path.join(ub, "J_array.npy")).tolist()}
    for name in ("norms.txt", "u_divergence.txt", "variables.txt"):
        with open(os.path.join(ub, name)) as fh:
            sc[name] = fh.read()
    scalars["u_bar_chapter_6.3.3"] = sc
    scalars["field_aliases"] = alias
    with open(os.path.join(GOLD, "scalars.json"), "w") as fh:
        json.dump(scalars, fh, indent=1)
    print("aliases:", alias)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
