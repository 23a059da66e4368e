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

import os
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


def checkpoint_counters(path: str, name: str) -> list:
    """Counters k of the groups ``/<name>/<name>_k`` present in the file, ascending.  The reference appends one group
    per gradient-descent iteration (``write_checkpoint(..., append=True)``, OCP_dolfin.py:440-441)."""
    ds = H5File(path).datasets()
    ks = set()
    for key in ds:
        parts = key.split("/")
        if len(parts) > 2 and parts[1] == name and parts[2].startswith(name + "_") and parts[2][len(name) + 1:].isdigit():
            ks.add(int(parts[2][len(name) + 1:]))
    return sorted(ks)


def read_checkpoint(path: str, name: str | None = None, counter: int = -1) -> Dict[str, np.ndarray]:
    """Return the datasets of a dolfin checkpoint as a flat dict.

    Keys: ``topology, geometry, cell_dofs, x_cell_dofs, cells, vector``.
    ``name`` selects the function group (``u``, ``p``, ``f``); if omitted the
    file must hold exactly one function.  ``counter`` selects the group ``<name>_<counter>``; the default -1 is
    dolfin's ``read_checkpoint(f, name)`` (counter = -1): the LAST group written, i.e. the latest iteration of an
    appended ``checkpoints/q.h5`` (OCP_dolfin.py:151-160, 440-441).
    """
    ds = H5File(path).datasets()
    groups = sorted({k.split("/")[1] for k in ds})
    if name is None:
        if len(groups) != 1:
            raise H5FormatError(f"{path}: several functions {groups}, pass name=")
        name = groups[0]
    ks = sorted({int(k.split("/")[2][len(name) + 1:]) for k in ds
                 if k.split("/")[1] == name and k.split("/")[2].startswith(name + "_")
                 and k.split("/")[2][len(name) + 1:].isdigit()})
    if not ks:
        raise H5FormatError(f"{path}: no group /{name}/{name}_<k>")
    if counter < 0:
        if -counter > len(ks):
            raise H5FormatError(f"{path}: counter {counter} but only {len(ks)} groups of {name}")
        k = ks[counter]
    elif counter in ks:
        k = counter
    else:
        raise H5FormatError(f"{path}: no group /{name}/{name}_{counter} (present: {ks})")
    pre = f"/{name}/{name}_{k}/"
    out = {}
    for key in ("mesh/topology", "mesh/geometry", "cell_dofs", "x_cell_dofs", "cells", "vector"):
        if pre + key not in ds:
            raise H5FormatError(f"{path}: missing dataset {pre + key}")
        out[key.split("/")[-1]] = ds[pre + key]
    out["vector"] = out["vector"].reshape(-1)
    out["cell_dofs"] = out["cell_dofs"].reshape(-1).astype(np.int64)
    out["topology"] = out["topology"].astype(np.int64)
    out["counter"] = k
    return out


# ---------------------------------------------------------------------------------------------------------------
# Writer: the same file structure dolfin's ``XDMFFile.write_checkpoint`` produces through libhdf5 (superblock v0,
# v1 object headers, symbol-table groups with a v1 B-tree + local heap, contiguous datasets carrying the dataspace
# / datatype / fill-value / layout / mtime messages and dolfin's ``partition`` attribute) - OCP_dolfin.py:440-441,
# 485-486, 578-588.  h5py / libhdf5 are not available here, so the writer is validated by reading its files back
# with ``H5File`` and by comparing its object-header message set with the reference's own checkpoints
# (tests/test_checkpoint_io.py).
# ---------------------------------------------------------------------------------------------------------------
_LEAF_K, _INTERNAL_K = 4, 16


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _dtype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBIHH", 0x10, bits0, 0, 0, dt.itemsize, 0, 8 * dt.itemsize) + b"\0" * 4
    if dt == np.float64:
        return bytes.fromhex("11203f000800000000004000340b0034ff03000000000000")
    raise H5FormatError(f"cannot write dtype {dt}")


def _message(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)          # superblock placeholder

    def alloc(self, data: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, arr: np.ndarray, mtime: int) -> int:
        arr = np.ascontiguousarray(arr)
        if arr.ndim == 1:
            arr = arr.reshape(-1, 1)
        raw = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()
        data_addr = self.alloc(raw)
        rank = arr.ndim
        dims = struct.pack("<%dQ" % rank, *arr.shape)
        dspace = struct.pack("<BBB5x", 1, rank, 1) + dims + dims            # max dims = dims
        attr_dt = _dtype_message(np.dtype("<u8"))[:12]
        attr = (struct.pack("<BBHHH", 1, 0, 10, 12, 24) + _pad8(b"partition\0") + _pad8(attr_dt)
                + struct.pack("<BBB5xQQ", 1, 1, 1, 1, 1) + struct.pack("<Q", 0))
        msgs = [
            _message(0x01, dspace),
            _message(0x03, _dtype_message(arr.dtype)),
            _message(0x05, bytes.fromhex("0202020100000000")),
            _message(0x08, struct.pack("<BBQQ", 3, 1, data_addr, len(raw))),
            _message(0x12, struct.pack("<B3xI", 1, mtime)),
            _message(0x0C, attr),
        ]
        return self.alloc(_object_header(msgs))

    def group(self, children: dict) -> tuple:
        """children: name -> (object header address, (btree, heap) for groups or None). Returns (hdr, btree, heap).
        Links are spread over symbol-table nodes of at most 2 * _LEAF_K entries under ONE level-0 B-tree node, which
        holds up to 2 * _INTERNAL_K of them: 256 links per group (a gradient-descent run appends one group per
        iteration; the reference's default is 50 iterations)."""
        names = sorted(children, key=lambda n: n.encode())      # libhdf5 orders links by strcmp
        per = 2 * _LEAF_K
        if len(names) > per * 2 * _INTERNAL_K:
            raise H5FormatError(f"more than {per * 2 * _INTERNAL_K} links per group are not supported")
        heap_data = bytearray(8)                       # offset 0: the empty name
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += _pad8(n.encode() + b"\0")
        if len(heap_data) <= 88 - 16:                  # libhdf5's default data segment, rest is one free block
            free_off = len(heap_data)                  # (a free block needs 16 bytes: next offset + size)
            free = bytearray(88 - free_off)
            struct.pack_into("<QQ", free, 0, 1, len(free))
            heap_data += free
        else:
            free_off = 1                               # H5HL_FREE_NULL
        data_addr = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, data_addr))
        chunks = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        snod_addrs = []
        for chunk in chunks:
            snod = bytearray(8 + per * 40)
            snod[:4] = b"SNOD"
            struct.pack_into("<BBH", snod, 4, 1, 0, len(chunk))
            for k, n in enumerate(chunk):
                hdr, sub = children[n]
                if sub is None:
                    struct.pack_into("<QQII16x", snod, 8 + 40 * k, offs[n], hdr, 0, 0)
                else:
                    struct.pack_into("<QQIIQQ", snod, 8 + 40 * k, offs[n], hdr, 1, 0, sub[0], sub[1])
            snod_addrs.append(self.alloc(bytes(snod)))
        tree = bytearray(24 + (2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8)
        tree[:4] = b"TREE"
        struct.pack_into("<BBHQQ", tree, 4, 0, 0, len(chunks), _UNDEF, _UNDEF)
        # key[0] | child[0] | key[1] | child[1] | ... : key[i+1] = heap offset of the largest name below child i
        struct.pack_into("<Q", tree, 24, 0)
        for i, chunk in enumerate(chunks):
            struct.pack_into("<QQ", tree, 32 + 16 * i, snod_addrs[i], offs[chunk[-1]] if chunk else 0)
        btree = self.alloc(bytes(tree))
        hdr = self.alloc(_object_header([_message(0x11, struct.pack("<QQ", btree, heap))]))
        return hdr, btree, heap

    def finish(self, root) -> bytes:
        hdr, btree, heap = root
        self.buf += b"\0" * (-len(self.buf) % 8)
        sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, len(self.buf), _UNDEF)
        sb += struct.pack("<QQIIQQ", 0, hdr, 1, 0, btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_checkpoint(path_h5: str, name: str, topology, geometry, cell_dofs, vector, dofs_per_cell: int,
                     mtime: int = 0, element_degree: int = 2, value_rank: int = 1, append: bool = False) -> str:
    """Write ``<name>/<name>_k/{mesh/{topology,geometry},cell_dofs,x_cell_dofs,cells,vector}`` like dolfin's
    ``write_checkpoint(fn, name, 0, append=...)`` and the companion ``.xdmf``; returns the xdmf path.

    ``append=False`` writes a single group ``<name>_0`` (dolfin truncates the file); ``append=True`` keeps the groups
    already in the file and adds ``<name>_{k+1}`` - the reference appends the control of every gradient-descent
    iteration to ``checkpoints/q.xdmf`` (OCP_dolfin.py:440-441).  The file is rebuilt and replaced atomically
    (temporary file + ``os.replace``), so a crash in mid-write never destroys the previous checkpoint."""
    new = dict(topology=np.asarray(topology, np.int32), geometry=np.asarray(geometry, np.float64),
               cell_dofs=np.asarray(cell_dofs, np.int32).reshape(-1), vector=np.asarray(vector, np.float64).reshape(-1))
    steps = []
    if append and os.path.exists(path_h5):
        for k in checkpoint_counters(path_h5, name):
            d = read_checkpoint(path_h5, name, k)
            steps.append(dict(topology=d["topology"].astype(np.int32), geometry=d["geometry"],
                              cell_dofs=d["cell_dofs"].astype(np.int32), vector=d["vector"]))
    steps.append(new)
    w = _Writer()
    ds = lambda a: (w.dataset(a, mtime), None)
    groups = {}
    for k, st in enumerate(steps):
        nck = st["topology"].shape[0]
        mesh = w.group({"topology": ds(st["topology"]), "geometry": ds(st["geometry"])})
        inner = w.group({
            "mesh": (mesh[0], mesh[1:]),
            "cell_dofs": ds(st["cell_dofs"]),
            "x_cell_dofs": ds(np.arange(nck + 1, dtype=np.uint64) * np.uint64(dofs_per_cell)),
            "cells": ds(np.arange(nck, dtype=np.uint64)),
            "vector": ds(st["vector"]),
        })
        groups[f"{name}_{k}"] = (inner[0], inner[1:])
    outer = w.group(groups)
    root = w.group({name: (outer[0], outer[1:])})
    tmp = path_h5 + ".tmp"
    with open(tmp, "wb") as fh:
        fh.write(w.finish(root))
    os.replace(tmp, path_h5)
    base = os.path.basename(path_h5)
    atype = "Vector" if value_rank == 1 else "Scalar"
    grids = []
    for k, st in enumerate(steps):
        grp = f"{base}:{name}/{name}_{k}"
        nc, nvert = st["topology"].shape[0], st["geometry"].shape[0]
        grids.append(f"""      <Grid Name="{name}_{k}" GridType="Uniform">
        <Topology NumberOfElements="{nc}" TopologyType="Triangle" NodesPerElement="3">
          <DataItem Dimensions="{nc} 3" NumberType="UInt" Format="HDF">{grp}/mesh/topology</DataItem>
        </Topology>
        <Geometry GeometryType="XY">
          <DataItem Dimensions="{nvert} 2" Format="HDF">{grp}/mesh/geometry</DataItem>
        </Geometry>
        <Time Value="0.000000000000000e+00" />
        <Attribute ItemType="FiniteElementFunction" ElementFamily="CG" ElementDegree="{element_degree}" ElementCell="triangle" Name="{name}" Center="Other" AttributeType="{atype}">
          <DataItem Dimensions="{st['cell_dofs'].size} 1" NumberType="UInt" Format="HDF">{grp}/cell_dofs</DataItem>
          <DataItem Dimensions="{st['vector'].size} 1" NumberType="Float" Format="HDF">{grp}/vector</DataItem>
          <DataItem Dimensions="{nc + 1} 1" NumberType="UInt" Format="HDF">{grp}/x_cell_dofs</DataItem>
          <DataItem Dimensions="{nc} 1" NumberType="UInt" Format="HDF">{grp}/cells</DataItem>
        </Attribute>
      </Grid>
""")
    xdmf = f"""<?xml version="1.0"?>
<Xdmf Version="3.0">
  <Domain>
    <Grid GridType="Collection" CollectionType="Temporal" Name="{name}">
{''.join(grids)}    </Grid>
  </Domain>
</Xdmf>
"""
    path_x = os.path.splitext(path_h5)[0] + ".xdmf"
    with open(path_x + ".tmp", "w") as fh:
        fh.write(xdmf)
    os.replace(path_x + ".tmp", path_x)
    return path_x
