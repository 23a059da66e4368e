"""Minimal reader for the legacy-dolfin XDMF/HDF5 "checkpoint" files.

The reference stores its FE functions with ``XDMFFile.write_checkpoint``
(OCP_dolfin.py:440-441, 485-486, 578-588) and reads them back with
``read_checkpoint`` (OCP_dolfin.py:151-160).  h5py is not available in this
image, so this module walks the HDF5 structures those files actually use:
superblock v0, v1 object headers, v1 group B-trees + local heaps, and
contiguous little-endian integer / IEEE-float datasets.  Anything else
(chunking, filters, v2 headers) raises ``H5FormatError``.

Layout written by dolfin for a function called ``name``::

    /<name>/<name>_0/mesh/topology   (ncells, 3)  int
    /<name>/<name>_0/mesh/geometry   (nverts, 2)  f8
    /<name>/<name>_0/cell_dofs       (sum dofs per cell,) int
    /<name>/<name>_0/x_cell_dofs     (ncells + 1,) uint64
    /<name>/<name>_0/cells           (ncells,) uint64
    /<name>/<name>_0/vector          (ndofs,) f8
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(RuntimeError):
    pass


class H5File:
    """Read-only view of one HDF5 file; ``datasets()`` maps path -> ndarray."""

    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self._b = fh.read()
        b = self._b
        if b[:8] != _SIG:
            raise H5FormatError(f"{path}: not an HDF5 file")
        if b[8] != 0:
            raise H5FormatError(f"{path}: superblock version {b[8]} unsupported (need 0)")
        if b[13] != 8 or b[14] != 8:
            raise H5FormatError("only 8-byte offsets/lengths supported")
        self._base = struct.unpack_from("<Q", b, 24)[0]
        # root group symbol-table entry starts after the four 8-byte addresses
        root_ste = 24 + 32
        self._root_header = struct.unpack_from("<Q", b, root_ste + 8)[0]
        self._cache: Dict[str, np.ndarray] | None = None

    # -- object headers ---------------------------------------------------
    def _messages(self, addr: int):
        b = self._b
        addr += self._base
        version, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", b, addr)
        if version != 1:
            raise H5FormatError(f"object header version {version} unsupported")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                data = pos + 8
                if mtype == 0x10:  # continuation block
                    off, ln = struct.unpack_from("<QQ", b, data)
                    blocks.append((off + self._base, ln))
                out.append((mtype, data, msize))
                pos = data + msize
        return out

    # -- groups -----------------------------------------------------------
    def _heap_data(self, heap_addr: int) -> int:
        b = self._b
        heap_addr += self._base
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5FormatError("bad local heap signature")
        return struct.unpack_from("<Q", b, heap_addr + 24)[0] + self._base

    def _walk_btree(self, addr: int, heap_data: int, out: Dict[str, int]):
        b = self._b
        addr += self._base
        if b[addr:addr + 4] == b"TREE":
            _ntype, level, nent = struct.unpack_from("<BBH", b, addr + 4)
            pos = addr + 24
            for i in range(nent):
                child = struct.unpack_from("<Q", b, pos + 8 + 16 * i)[0]
                self._walk_btree(child, heap_data, out)
        elif b[addr:addr + 4] == b"SNOD":
            nsym = struct.unpack_from("<H", b, addr + 6)[0]
            for i in range(nsym):
                e = addr + 8 + 40 * i
                name_off, hdr = struct.unpack_from("<QQ", b, e)
                s = heap_data + name_off
                name = b[s:b.index(b"\0", s)].decode()
                out[name] = hdr
        else:
            raise H5FormatError("bad group node signature")

    def _children(self, hdr: int) -> Dict[str, int] | None:
        for mtype, data, _ in self._messages(hdr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", self._b, data)
                out: Dict[str, int] = {}
                self._walk_btree(btree, self._heap_data(heap), out)
                return out
        return None

    # -- datasets ---------------------------------------------------------
    def _dataset(self, hdr: int) -> np.ndarray:
        b = self._b
        shape = dtype = None
        addr = size = None
        for mtype, data, msize in self._messages(hdr):
            if mtype == 0x01:
                ver, rank, _fl = struct.unpack_from("<BBB", b, data)
                off = data + (8 if ver == 1 else 4)
                shape = struct.unpack_from("<%dQ" % rank, b, off)
            elif mtype == 0x03:
                cv, bits0 = b[data], b[data + 1]
                cls = cv & 0x0F
                nbytes = struct.unpack_from("<I", b, data + 4)[0]
                if bits0 & 1:
                    raise H5FormatError("big-endian datasets unsupported")
                if cls == 0:
                    dtype = np.dtype("<%s%d" % ("i" if bits0 & 0x08 else "u", nbytes))
                elif cls == 1:
                    dtype = np.dtype("<f%d" % nbytes)
                else:
                    raise H5FormatError(f"datatype class {cls} unsupported")
            elif mtype == 0x08:
                ver, lclass = b[data], b[data + 1]
                if ver != 3 or lclass != 1:
                    raise H5FormatError("only contiguous (v3 layout) datasets supported")
                addr, size = struct.unpack_from("<QQ", b, data + 2)
        if shape is None or dtype is None or addr is None:
            raise H5FormatError("incomplete dataset header")
        count = int(np.prod(shape)) if len(shape) else 1
        if addr == _UNDEF:
            return np.zeros(shape, dtype)
        arr = np.frombuffer(b, dtype=dtype, count=count, offset=addr + self._base)
        return arr.reshape(shape).copy()

    def datasets(self) -> Dict[str, np.ndarray]:
        if self._cache is None:
            out: Dict[str, np.ndarray] = {}

            def rec(prefix: str, hdr: int):
                kids = self._children(hdr)
                if kids is None:
                    out[prefix] = self._dataset(hdr)
                    return
                for name, child in sorted(kids.items()):
                    rec(prefix + "/" + name, child)

            rec("", self._root_header)
            self._cache = out
        return self._cache


def read_checkpoint(path: str, name: str | None = None) -> Dict[str, np.ndarray]:
    """Return the datasets of a dolfin checkpoint as a flat dict.

    Keys: ``topology, geometry, cell_dofs, x_cell_dofs, cells, vector``.
    ``name`` selects the function group (``u``, ``p``, ``f``); if omitted the
    file must hold exactly one function.
    """
    ds = H5File(path).datasets()
    groups = sorted({k.split("/")[1] for k in ds})
    if name is None:
        if len(groups) != 1:
            raise H5FormatError(f"{path}: several functions {groups}, pass name=")
        name = groups[0]
    pre = f"/{name}/{name}_0/"
    out = {}
    for key in ("mesh/topology", "mesh/geometry", "cell_dofs", "x_cell_dofs", "cells", "vector"):
        if pre + key not in ds:
            raise H5FormatError(f"{path}: missing dataset {pre + key}")
        out[key.split("/")[-1]] = ds[pre + key]
    out["vector"] = out["vector"].reshape(-1)
    out["cell_dofs"] = out["cell_dofs"].reshape(-1).astype(np.int64)
    out["topology"] = out["topology"].astype(np.int64)
    return out
