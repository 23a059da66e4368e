"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.  ctypes front-end of oracle/buoy_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this module.  Function-by-function citations are in buoy_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbuoy_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "buoy_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", src, "-o", _SO, "-lm"]
        )
    return _SO


class _Tables(C.Structure):
    _fields_ = [
        ("nc", C.c_int32), ("nn", C.c_int32), ("nv", C.c_int32),
        ("geom", C.c_void_p), ("cell_nodes", C.c_void_p),
        ("ox", C.c_double), ("oy", C.c_double), ("ihx", C.c_double), ("ihy", C.c_double),
        ("nbx", C.c_int32), ("nby", C.c_int32),
        ("bin_ptr", C.c_void_p), ("bin_cells", C.c_void_p),
        ("brute", C.c_int32),
    ]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class BuoyOracle:
    def __init__(self, V, brute: bool = False):
        self.lib = C.CDLL(build())
        self.lib.oracle_misfit.restype = C.c_double
        self.V = V
        self._keep = dict(
            geom=np.ascontiguousarray(V.cell_geom, np.float64),
            cn=np.ascontiguousarray(V.cell_nodes, np.int32),
            bp=np.ascontiguousarray(V.bin_ptr, np.int32),
            bc=np.ascontiguousarray(V.bin_cells, np.int32),
        )
        k = self._keep
        self.t = _Tables(
            V.mesh.num_cells, V.num_nodes, V.mesh.num_vertices, _p(k["geom"]), _p(k["cn"]),
            float(V.bin_origin[0]), float(V.bin_origin[1]), float(V.bin_inv_h[0]), float(V.bin_inv_h[1]),
            int(V.bin_dims[0]), int(V.bin_dims[1]), _p(k["bp"]), _p(k["bc"]), int(brute),
        )

    def set_threads(self, n: int):
        """1 = the reference's serial loop structure; >1 spreads the per-buoy loops over host threads."""
        self.lib.oracle_set_threads(int(n))

    def max_threads(self) -> int:
        return int(self.lib.oracle_max_threads())

    def forward(self, vel, x0, nt, h, center, mask=None):
        """solve_primal_ode: returns x, u (K,nt,2), cell (K,nt), mask (K) f8, parked (K) u1."""
        vel = np.ascontiguousarray(vel, np.float64)
        x0 = np.ascontiguousarray(x0, np.float64)
        K = x0.shape[0]
        x = np.empty((K, nt, 2))
        u = np.empty((K, nt, 2))
        cell = np.empty((K, nt), np.int32)
        mask = np.zeros(K) if mask is None else mask
        parked = np.zeros(K, np.uint8)
        ctr = np.ascontiguousarray(center, np.float64)
        self.lib.oracle_buoy_forward(C.byref(self.t), _p(vel), K, nt, C.c_double(h), _p(x0), _p(ctr),
                                     _p(x), _p(u), _p(cell), _p(mask), _p(parked))
        return x, u, cell, mask, parked

    def adjoint(self, g, x, u, ud, mask, h):
        K, nt, _ = x.shape
        mu = np.empty((K, nt, 2))
        args = [np.ascontiguousarray(a, np.float64) for a in (g, x, u, ud, mask)]
        self.lib.oracle_buoy_adjoint(C.byref(self.t), _p(args[0]), K, nt, C.c_double(h), _p(args[1]), _p(args[2]),
                                     _p(args[3]), _p(args[4]), _p(mu))
        return mu

    def point_sources(self, vel, x, ud, mu, mask, h, center):
        K, nt, _ = x.shape
        bnode = np.zeros((self.V.num_nodes, 2))
        args = [np.ascontiguousarray(a, np.float64) for a in (vel, x, ud, mu, mask, center)]
        self.lib.oracle_point_sources(C.byref(self.t), _p(args[0]), K, nt, C.c_double(h), _p(args[1]), _p(args[2]),
                                      _p(args[3]), _p(args[4]), _p(args[5]), _p(bnode))
        return bnode

    def misfit(self, u, ud, h):
        K, nt, _ = u.shape
        u, ud = np.ascontiguousarray(u, np.float64), np.ascontiguousarray(ud, np.float64)
        return float(self.lib.oracle_misfit(K, nt, C.c_double(h), _p(u), _p(ud)))
