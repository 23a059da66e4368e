"""Checkpoint writer/reader in the reference's XDMF+HDF5 layout (OCP_dolfin.py:151-160, 440-441, 485-486, 578-588)."""
import os

import numpy as np
import pytest

import helpers as H
from ocp_b200 import checkpoint, h5lite

REF = "/root/reference/reference_runs/u_bar_chapter_6.3.3"


def test_control_round_trip_and_foreign_numbering(tmp_path):
    V = H.square32()
    q = H.q_nodal(V)
    x = checkpoint.write_control(str(tmp_path / "q.h5"), V, q)
    assert os.path.exists(x)
    assert np.array_equal(checkpoint.read_control(str(tmp_path / "q.h5"), V), q)
    # a file in the reference's own (graph-reordered) numbering of the collapsed space reads to the same field
    F = H.fields()
    h5lite.write_checkpoint(str(tmp_path / "q_ref.h5"), "f", V.mesh.cells, V.mesh.coords, F["q_cell_dofs"], F["q_vector"], 12)
    assert np.array_equal(checkpoint.read_control(str(tmp_path / "q_ref.h5"), V), q)


def test_state_round_trip(tmp_path):
    V = H.square32()
    w = H.fields()["u_bar"]
    checkpoint.write_state(str(tmp_path), V, w)
    assert np.array_equal(checkpoint.read_state(str(tmp_path / "u.h5"), V, "u"), w)
    assert np.array_equal(checkpoint.read_state(str(tmp_path / "p.h5"), V, "p"), w)
    d = h5lite.read_checkpoint(str(tmp_path / "p.h5"), "p")
    assert d["cell_dofs"].size == 6144 and np.array_equal(d["x_cell_dofs"].ravel(), 3 * np.arange(2049))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference mount absent")
def test_written_files_have_the_structure_of_dolfins(tmp_path):
    """Same datasets, same xdmf text and the same object-header message set (type, size) as the reference's files."""
    ref = os.path.join(REF, "paraview/checkpoint/u.h5")
    d = h5lite.read_checkpoint(ref, "u")
    x = h5lite.write_checkpoint(str(tmp_path / "u.h5"), "u", d["topology"], d["geometry"], d["cell_dofs"], d["vector"], 12)
    assert open(x).read() == open(ref.replace(".h5", ".xdmf")).read()

    def messages(path):
        f = h5lite.H5File(path)
        out = {}

        def rec(prefix, hdr):
            out[prefix or "/"] = [(t, sz) for t, _, sz in f._messages(hdr) if t != 0]     # NIL padding ignored
            for k, v in (f._children(hdr) or {}).items():
                rec(prefix + "/" + k, v)
        rec("", f._root_header)
        return out
    assert messages(ref) == messages(str(tmp_path / "u.h5"))
    e = h5lite.read_checkpoint(str(tmp_path / "u.h5"), "u")
    for k in d:
        assert np.array_equal(np.asarray(d[k]).ravel(), np.asarray(e[k]).ravel())
    # the reference's control checkpoint reads into the nodal field used by the K4/K5 tests
    V = H.square32()
    assert np.array_equal(checkpoint.read_control(os.path.join(REF, "q_backup/q.h5"), V), H.q_nodal(V))


def test_appended_checkpoints_read_back_the_last_group(tmp_path):
    """The reference appends one group f_k per GD iteration (OCP_dolfin.py:440-441) and dolfin's
    read_checkpoint(f, "f") returns the LAST one: multi-step files, counter semantics, more than 8 links per group."""
    V = H.square32()
    q = H.q_nodal(V)
    path = str(tmp_path / "q.h5")
    n = 21                                                    # > 2 symbol-table nodes (8 links each)
    for k in range(n):
        checkpoint.write_control(path, V, q * (k + 1), append=k > 0)
    assert h5lite.checkpoint_counters(path, "f") == list(range(n))
    assert np.array_equal(checkpoint.read_control(path, V), q * n)             # default: the latest iteration
    assert np.array_equal(checkpoint.read_control(path, V, counter=0), q)
    assert np.array_equal(checkpoint.read_control(path, V, counter=9), q * 10)
    assert np.array_equal(checkpoint.read_control(path, V, counter=-2), q * (n - 1))
    with pytest.raises(h5lite.H5FormatError):
        checkpoint.read_control(path, V, counter=n)
    assert open(str(tmp_path / "q.xdmf")).read().count("<Grid Name=\"f_") == n
    # append=False truncates like dolfin
    checkpoint.write_control(path, V, q)
    assert h5lite.checkpoint_counters(path, "f") == [0]
    # a non-finite control never replaces the last good checkpoint, and no temporary file is left behind
    bad = q.copy()
    bad[3, 1] = np.nan
    with pytest.raises(ValueError):
        checkpoint.write_control(path, V, bad, append=True)
    assert np.array_equal(checkpoint.read_control(path, V), q)
    assert sorted(os.listdir(tmp_path)) == ["q.h5", "q.xdmf"]
