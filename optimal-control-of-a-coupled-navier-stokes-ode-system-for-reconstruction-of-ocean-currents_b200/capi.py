"""ctypes binding of libocp_b200.so (include/ocp_b200.h).

Device buffers are torch CUDA tensors (torch is only the allocator / stream / NCCL plumbing);
the library receives raw ``data_ptr()`` addresses.  If the shared library is missing it is an
error - there is no Python or CPU fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libocp_b200.so")

OCP_OK = 0
ERRORS = {-1: "invalid argument", -2: "CUDA error", -3: "sparse solver error", -4: "Newton did not converge",
          -5: "no CUDA device", -6: "NCCL communicator error"}

# every symbol include/ocp_b200.h declares (tests/test_capi_cpu.py checks the header against this list)
SYMBOLS = [
    "ocp_version", "ocp_device_available", "ocp_create", "ocp_destroy", "ocp_last_error",
    "ocp_get_solver_stats", "ocp_reset_solver_stats", "ocp_set_viscosity", "ocp_set_profiling", "ocp_set_dirichlet",
    "ocp_set_deterministic", "ocp_get_option",
    "ocp_forward_solve", "ocp_assemble_forward", "ocp_assemble_adjoint", "ocp_project_grad",
    "ocp_velocity_nodal", "ocp_buoy_forward", "ocp_buoy_adjoint_scatter", "ocp_misfit",
    "ocp_adjoint_solve", "ocp_boundary_inner", "ocp_nodal_axpby", "ocp_field_norms", "ocp_traj_transpose",
    "ocp_solve_primal_ode_host", "ocp_solve_adjoint_ode_host", "ocp_set_observations_host", "ocp_gradient_host",
    "ocp_gradient_device", "ocp_newton_history",
    "ocp_launch_count", "ocp_get_solver_info", "ocp_selftest_fp64_peak",
    "ocp_comm_get_unique_id", "ocp_comm_init", "ocp_comm_size", "ocp_comm_nccl_version", "ocp_allreduce",
    "ocp_host_gather_probe", "ocp_host_lu_probe", "ocp_host_mf_probe", "ocp_host_mf_set_pivot_window", "ocp_selftest_cell_matrix", "ocp_selftest_facet_matrix",
]


class OcpError(RuntimeError):
    pass


class ProblemDesc(C.Structure):
    _fields_ = [
        ("nv", C.c_int32), ("nn", C.c_int32), ("nc", C.c_int32), ("ndofs", C.c_int32), ("nnz", C.c_int32),
        ("n_dirichlet", C.c_int32), ("n_g1", C.c_int32), ("nt", C.c_int32),
        ("cell_geom", C.c_void_p), ("cell_nodes", C.c_void_p), ("cell_nbr", C.c_void_p), ("node_coords", C.c_void_p),
        ("dof_ux", C.c_void_p), ("dof_uy", C.c_void_p), ("dof_p", C.c_void_p),
        ("csr_rowptr", C.c_void_p), ("csr_col", C.c_void_p), ("dirichlet_dofs", C.c_void_p),
        ("g1_nodes", C.c_void_p), ("g1_len", C.c_void_p), ("g1_normal", C.c_void_p),
        ("bin_ox", C.c_double), ("bin_oy", C.c_double), ("bin_ihx", C.c_double), ("bin_ihy", C.c_double),
        ("nbx", C.c_int32), ("nby", C.c_int32),
        ("bin_ptr", C.c_void_p), ("bin_cells", C.c_void_p),
        ("viscosity", C.c_double), ("dt", C.c_double), ("center_x", C.c_double), ("center_y", C.c_double),
    ]


class SolverStats(C.Structure):
    _fields_ = [("assemble_ms", C.c_double), ("factor_ms", C.c_double), ("solve_ms", C.c_double),
                ("analyse_ms", C.c_double), ("n_factor", C.c_int32), ("n_solve", C.c_int32),
                ("n_dense", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def load_library() -> C.CDLL:
    """dlopen the CUDA library; raises if it has not been built (``python -m ocp_b200.build``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OcpError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); "
                           "there is no CPU fallback for this path")
        lib = C.CDLL(LIB_PATH)
        lib.ocp_last_error.restype = C.c_char_p
        lib.ocp_last_error.argtypes = [C.c_void_p]
        lib.ocp_destroy.argtypes = [C.c_void_p]
        lib.ocp_destroy.restype = None
        lib.ocp_host_lu_probe.restype = C.c_int64
        lib.ocp_host_mf_probe.restype = C.c_int64
        lib.ocp_host_gather_probe.restype = C.c_int64
        lib.ocp_host_mf_set_pivot_window.restype = None
        lib.ocp_launch_count.restype = C.c_longlong
        lib.ocp_set_viscosity.argtypes = [C.c_void_p, C.c_double]
        lib.ocp_set_viscosity.restype = None
        lib.ocp_get_solver_stats.restype = None
        lib.ocp_reset_solver_stats.restype = None
        lib.ocp_set_profiling.restype = None
        lib.ocp_selftest_cell_matrix.restype = None
        lib.ocp_selftest_facet_matrix.restype = None
        lib.ocp_get_option.argtypes = [C.c_void_p, C.c_char_p]
        lib.ocp_set_deterministic.argtypes = [C.c_void_p, C.c_int]
        lib.ocp_get_solver_info.restype = None
        lib.ocp_get_solver_info.argtypes = [C.c_void_p, C.c_void_p]
        lib.ocp_allreduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.ocp_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.ocp_comm_size.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def _hp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _dp(t):
    """device pointer of a torch tensor (or None)"""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda or not t.is_contiguous():
        raise OcpError("expected a contiguous CUDA tensor")
    return C.c_void_p(t.data_ptr())


def problem_desc(V, viscosity: float, dt: float, nt: int, center):
    """(ProblemDesc, arrays that must stay alive) for a TaylorHood space."""
    m = V.mesh
    keep = dict(
        geom=np.ascontiguousarray(V.cell_geom, np.float64),
        cn=np.ascontiguousarray(V.cell_nodes, np.int32),
        nbr=np.ascontiguousarray(V.cell_nbr, np.int32),
        xy=np.ascontiguousarray(V.node_coords, np.float64),
        ux=np.ascontiguousarray(V.dof_ux, np.int32), uy=np.ascontiguousarray(V.dof_uy, np.int32),
        pp=np.ascontiguousarray(V.dof_p, np.int32),
        rp=np.ascontiguousarray(V.csr_rowptr, np.int32), ci=np.ascontiguousarray(V.csr_col, np.int32),
        dd=np.ascontiguousarray(V.dirichlet_dofs, np.int32),
        g1n=np.ascontiguousarray(V.g1_nodes, np.int32), g1l=np.ascontiguousarray(V.g1_len, np.float64),
        g1m=np.ascontiguousarray(V.g1_normal, np.float64),
        bp=np.ascontiguousarray(V.bin_ptr, np.int32), bc=np.ascontiguousarray(V.bin_cells, np.int32),
    )
    d = ProblemDesc(
        m.num_vertices, V.num_nodes, m.num_cells, V.ndofs, int(V.csr_col.size), int(V.dirichlet_dofs.size),
        int(V.g1_cell.size), int(nt),
        _hp(keep["geom"]), _hp(keep["cn"]), _hp(keep["nbr"]), _hp(keep["xy"]), _hp(keep["ux"]), _hp(keep["uy"]), _hp(keep["pp"]),
        _hp(keep["rp"]), _hp(keep["ci"]), _hp(keep["dd"]), _hp(keep["g1n"]), _hp(keep["g1l"]), _hp(keep["g1m"]),
        float(V.bin_origin[0]), float(V.bin_origin[1]), float(V.bin_inv_h[0]), float(V.bin_inv_h[1]),
        int(V.bin_dims[0]), int(V.bin_dims[1]), _hp(keep["bp"]), _hp(keep["bc"]),
        float(viscosity), float(dt), float(center[0]), float(center[1]),
    )
    return d, keep


def host_gather_probe(V, w, nu: float = 1.0):
    """Host emulation of the atomic-free assembly tables: returns (vals, res, vals_transposed, stats)."""
    lib = load_library()
    d, keep = problem_desc(V, nu, 0.005, 200, (1.0, 1.0))
    w = np.ascontiguousarray(w, np.float64)
    vals, res, vt = np.zeros(int(V.csr_col.size)), np.zeros(V.ndofs), np.zeros(int(V.csr_col.size))
    st = np.zeros(4, np.int32)
    rc = lib.ocp_host_gather_probe(C.byref(d), _hp(w), C.c_double(nu), _hp(vals), _hp(res), _hp(vt), _hp(st))
    if rc < 0:
        raise OcpError(f"gather tables not representable for this mesh ({ERRORS.get(rc, rc)})")
    return vals, res, vt, dict(ctas=int(st[0]), rounds=int(st[1]), colors=int(st[2]), smem_bytes=int(st[3]))


class Context:
    """One problem (mesh + space + parameters) on one GPU."""

    def __init__(self, V, viscosity: float, dt: float, nt: int, center, stream: int = 0):
        self.lib = load_library()
        if self.lib.ocp_device_available() != OCP_OK:
            raise OcpError("no CUDA device: libocp_b200 has no CPU fallback")
        self.V = V
        m = V.mesh
        keep = dict(
            geom=np.ascontiguousarray(V.cell_geom, np.float64),
            cn=np.ascontiguousarray(V.cell_nodes, np.int32),
            nbr=np.ascontiguousarray(V.cell_nbr, np.int32),
            xy=np.ascontiguousarray(V.node_coords, np.float64),
            ux=np.ascontiguousarray(V.dof_ux, np.int32), uy=np.ascontiguousarray(V.dof_uy, np.int32),
            pp=np.ascontiguousarray(V.dof_p, np.int32),
            rp=np.ascontiguousarray(V.csr_rowptr, np.int32), ci=np.ascontiguousarray(V.csr_col, np.int32),
            dd=np.ascontiguousarray(V.dirichlet_dofs, np.int32),
            g1n=np.ascontiguousarray(V.g1_nodes, np.int32), g1l=np.ascontiguousarray(V.g1_len, np.float64),
            g1m=np.ascontiguousarray(V.g1_normal, np.float64),
            bp=np.ascontiguousarray(V.bin_ptr, np.int32), bc=np.ascontiguousarray(V.bin_cells, np.int32),
        )
        self._keep = keep
        d = ProblemDesc(
            m.num_vertices, V.num_nodes, m.num_cells, V.ndofs, int(V.csr_col.size), int(V.dirichlet_dofs.size),
            int(V.g1_cell.size), int(nt),
            _hp(keep["geom"]), _hp(keep["cn"]), _hp(keep["nbr"]), _hp(keep["xy"]), _hp(keep["ux"]), _hp(keep["uy"]), _hp(keep["pp"]),
            _hp(keep["rp"]), _hp(keep["ci"]), _hp(keep["dd"]), _hp(keep["g1n"]), _hp(keep["g1l"]), _hp(keep["g1m"]),
            float(V.bin_origin[0]), float(V.bin_origin[1]), float(V.bin_inv_h[0]), float(V.bin_inv_h[1]),
            int(V.bin_dims[0]), int(V.bin_dims[1]), _hp(keep["bp"]), _hp(keep["bc"]),
            float(viscosity), float(dt), float(center[0]), float(center[1]),
        )
        self.nt = int(nt)
        self.nn, self.nv, self.ndofs, self.nnz = V.num_nodes, m.num_vertices, V.ndofs, int(V.csr_col.size)
        h = C.c_void_p(0)
        rc = self.lib.ocp_create(C.byref(d), C.c_void_p(stream), C.byref(h))
        self._h = h
        if rc != OCP_OK:
            msg = self._err()
            self.close()
            raise OcpError(f"ocp_create failed ({ERRORS.get(rc, rc)}): {msg}")

    # -- plumbing ---------------------------------------------------------------
    def _err(self) -> str:
        if not self._h:
            return ""
        s = self.lib.ocp_last_error(self._h)
        return s.decode() if s else ""

    def _check(self, rc: int, what: str):
        if rc != OCP_OK:
            raise OcpError(f"{what} failed ({ERRORS.get(rc, rc)}): {self._err()}")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ocp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solver_stats(self) -> dict:
        st = SolverStats()
        self.lib.ocp_get_solver_stats(self._h, C.byref(st))
        return {k: getattr(st, k) for k, _ in SolverStats._fields_}

    def reset_solver_stats(self):
        self.lib.ocp_reset_solver_stats(self._h)

    def solver_info(self) -> dict:
        out = np.zeros(8)
        self.lib.ocp_get_solver_info(self._h, _hp(out))
        keys = ["factor_flops", "factor_nnz", "levels", "max_front", "workspace_doubles", "mass_flops", "mass_nnz",
                "fronts"]
        return dict(zip(keys, out.tolist()))

    def fp64_peak_tflops(self) -> float:
        v = C.c_double(0.0)
        self._check(self.lib.ocp_selftest_fp64_peak(self._h, C.byref(v)), "ocp_selftest_fp64_peak")
        return float(v.value)

    def set_profiling(self, on: bool):
        """Line-item timing synchronises after every phase: keep it off outside profiling runs."""
        self.lib.ocp_set_profiling(self._h, int(bool(on)))

    def set_dirichlet(self, dofs: np.ndarray, vals=None):
        dofs = np.ascontiguousarray(dofs, np.int32)
        v = None if vals is None else np.ascontiguousarray(vals, np.float64)
        self._check(self.lib.ocp_set_dirichlet(self._h, _hp(dofs), _hp(v) if v is not None else C.c_void_p(0),
                                               int(dofs.size)), "ocp_set_dirichlet")

    def set_deterministic(self, on: bool):
        """Reproducible (integer fixed-point) point-source deposit instead of fp64 atomics."""
        self._check(self.lib.ocp_set_deterministic(self._h, int(bool(on))), "ocp_set_deterministic")

    def option(self, name: str) -> int:
        return int(self.lib.ocp_get_option(self._h, name.encode()))

    def set_viscosity(self, nu: float):
        self.lib.ocp_set_viscosity(self._h, float(nu))

    # -- multi-GPU exchange (NCCL inside the library) -------------------------------------
    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        """Collective: every rank calls it with the 128-byte id rank 0 obtained from ``comm_unique_id()``."""
        if len(unique_id) != COMM_ID_BYTES:
            raise OcpError("unique id must be 128 bytes")
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        self._check(self.lib.ocp_comm_init(self._h, int(nranks), int(rank), buf), "ocp_comm_init")

    def comm_size(self) -> int:
        return int(self.lib.ocp_comm_size(self._h))

    def allreduce(self, d_buf):
        self._check(self.lib.ocp_allreduce(self._h, _dp(d_buf), C.c_size_t(d_buf.numel())), "ocp_allreduce")

    # -- device entry points ----------------------------------------------------------
    def forward_solve(self, d_f, d_w, zero_init: bool = True):
        its = C.c_int(0)
        hist = (C.c_double * 64)()
        rc = self.lib.ocp_forward_solve(self._h, _dp(d_f), _dp(d_w), int(zero_init), C.byref(its), hist)
        self._check(rc, "ocp_forward_solve")
        return its.value, [hist[i] for i in range(its.value + 1)]

    def assemble_forward(self, d_w, d_f, d_vals, d_res, apply_bc: bool):
        self._check(self.lib.ocp_assemble_forward(self._h, _dp(d_w), _dp(d_f), _dp(d_vals), _dp(d_res),
                                                  int(apply_bc)), "ocp_assemble_forward")

    def assemble_adjoint(self, d_w, d_vals, apply_bc: bool):
        self._check(self.lib.ocp_assemble_adjoint(self._h, _dp(d_w), _dp(d_vals), int(apply_bc)),
                    "ocp_assemble_adjoint")

    def project_grad(self, d_w, d_g):
        self._check(self.lib.ocp_project_grad(self._h, _dp(d_w), _dp(d_g)), "ocp_project_grad")

    def velocity_nodal(self, d_w, d_vel):
        self._check(self.lib.ocp_velocity_nodal(self._h, _dp(d_w), _dp(d_vel)), "ocp_velocity_nodal")

    def buoy_forward(self, d_vel, d_x0, K, d_x, d_u, d_cell, d_mask, d_parked):
        self._check(self.lib.ocp_buoy_forward(self._h, _dp(d_vel), _dp(d_x0), int(K), _dp(d_x), _dp(d_u),
                                              _dp(d_cell), _dp(d_mask), _dp(d_parked)), "ocp_buoy_forward")

    def buoy_adjoint_scatter(self, d_vel, d_g, K, d_x, d_u, d_ud, d_mask, d_parked, d_mu, d_acc):
        self._check(self.lib.ocp_buoy_adjoint_scatter(self._h, _dp(d_vel), _dp(d_g), int(K), _dp(d_x), _dp(d_u),
                                                      _dp(d_ud), _dp(d_mask), _dp(d_parked), _dp(d_mu), _dp(d_acc)),
                    "ocp_buoy_adjoint_scatter")

    def misfit(self, K, d_u, d_ud, d_out):
        self._check(self.lib.ocp_misfit(self._h, int(K), _dp(d_u), _dp(d_ud), _dp(d_out)), "ocp_misfit")

    def adjoint_solve(self, d_w, d_bnode, d_z):
        self._check(self.lib.ocp_adjoint_solve(self._h, _dp(d_w), _dp(d_bnode), _dp(d_z)), "ocp_adjoint_solve")

    def boundary_inner(self, d_a, d_b, d_out):
        self._check(self.lib.ocp_boundary_inner(self._h, _dp(d_a), _dp(d_b), _dp(d_out)), "ocp_boundary_inner")

    def nodal_axpby(self, ca, d_a, cb, d_b, d_out):
        self._check(self.lib.ocp_nodal_axpby(self._h, C.c_double(ca), _dp(d_a), C.c_double(cb), _dp(d_b),
                                             _dp(d_out)), "ocp_nodal_axpby")

    def field_norms(self, d_w, d_out):
        self._check(self.lib.ocp_field_norms(self._h, _dp(d_w), _dp(d_out)), "ocp_field_norms")

    def traj_transpose(self, d_src, d_dst, K, to_time_major: bool):
        self._check(self.lib.ocp_traj_transpose(self._h, _dp(d_src), _dp(d_dst), int(K), int(to_time_major)),
                    "ocp_traj_transpose")

    def gradient_device(self, d_f, d_x0, d_ud, K, d_w, d_g, d_vel, d_x, d_u, d_mask, d_parked, d_acc, d_z, d_znod,
                        d_grad, alpha: float) -> int:
        """The whole "outer" block on device buffers (one CUDA graph replay once warm); returns the Newton count."""
        its = C.c_int(0)
        rc = self.lib.ocp_gradient_device(self._h, _dp(d_f), _dp(d_x0), _dp(d_ud), int(K), _dp(d_w), _dp(d_g), _dp(d_vel),
                                          _dp(d_x), _dp(d_u), _dp(d_mask), _dp(d_parked), _dp(d_acc), _dp(d_z),
                                          _dp(d_znod), _dp(d_grad), C.c_double(alpha), C.byref(its))
        self._check(rc, "ocp_gradient_device")
        return its.value

    def newton_history(self):
        hist = (C.c_double * 64)()
        n = self.lib.ocp_newton_history(self._h, hist, 64)
        return [hist[i] for i in range(n)]

    # -- host-buffer entry points (numpy in / numpy out, copies inside the call) -------------------------------
    def solve_primal_ode_host(self, w: np.ndarray, x0: np.ndarray, mask: np.ndarray):
        K = x0.shape[0]
        w = np.ascontiguousarray(w, np.float64)
        x0 = np.ascontiguousarray(x0, np.float64)
        x = np.empty((K, self.nt, 2))
        u = np.empty((K, self.nt, 2))
        self._check(self.lib.ocp_solve_primal_ode_host(self._h, _hp(w), _hp(x0), K, _hp(x), _hp(u), _hp(mask)),
                    "ocp_solve_primal_ode_host")
        return x, u

    def solve_adjoint_ode_host(self, g, x, u, ud, mask):
        K = x.shape[0]
        args = [np.ascontiguousarray(a, np.float64) for a in (g, x, u, ud, mask)]
        mu = np.empty((K, self.nt, 2))
        self._check(self.lib.ocp_solve_adjoint_ode_host(self._h, *[_hp(a) for a in args], K, _hp(mu)),
                    "ocp_solve_adjoint_ode_host")
        return mu

    def set_observations_host(self, x0: np.ndarray, ud: np.ndarray):
        x0, ud = np.ascontiguousarray(x0, np.float64), np.ascontiguousarray(ud, np.float64)
        self._obs_K = x0.shape[0]
        self._check(self.lib.ocp_set_observations_host(self._h, _hp(x0), _hp(ud), self._obs_K),
                    "ocp_set_observations_host")

    def gradient_host(self, f, out=None):
        """numpy control (nn,2) in -> w, z, mask (numpy) and scalars; `out` lets the caller reuse (pinned) buffers."""
        f = np.ascontiguousarray(f, np.float64)
        if out is None:
            out = (np.empty(self.ndofs), np.empty(self.ndofs), np.zeros(self._obs_K), np.zeros(4))
        w, z, mask, sc = out
        self._check(self.lib.ocp_gradient_host(self._h, _hp(f), _hp(w), _hp(z), _hp(mask), _hp(sc)),
                    "ocp_gradient_host")
        return w, z, mask, dict(misfit=sc[0], f_norm2=sc[1], n_masked=int(sc[2]), newton_its=int(sc[3]))


COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0); ship the bytes to the other ranks and call Context.comm_init."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = load_library().ocp_comm_get_unique_id(buf)
    if rc != OCP_OK:
        raise OcpError(f"ocp_comm_get_unique_id failed ({ERRORS.get(rc, rc)}): is libnccl.so.2 loadable?")
    return buf.raw


def nccl_version() -> int:
    return int(load_library().ocp_comm_nccl_version())


def launch_count() -> int:
    return int(load_library().ocp_launch_count())


# -- element-level self-tests and host analysis (no GPU needed) ------------------------------------------------
def selftest_cell_matrix(geom6, coef15, nu):
    lib = load_library()
    g = np.ascontiguousarray(geom6, np.float64)
    cf = np.ascontiguousarray(coef15, np.float64)
    A = np.zeros((15, 15))
    R = np.zeros(15)
    lib.ocp_selftest_cell_matrix(_hp(g), _hp(cf), C.c_double(nu), _hp(A), _hp(R))
    return A, R


def selftest_facet_matrix(length, nx, ny, uv6, f6):
    lib = load_library()
    uv = np.ascontiguousarray(uv6, np.float64)
    ff = np.ascontiguousarray(f6, np.float64)
    A = np.zeros((6, 6))
    R = np.zeros(6)
    lib.ocp_selftest_facet_matrix(C.c_double(length), C.c_double(nx), C.c_double(ny), _hp(uv), _hp(ff), _hp(A), _hp(R))
    return A, R


def host_lu_probe(rowptr, col, val, xy, rhs):
    """One-time host analysis (nested dissection + threshold-pivoted LU) on a CSR matrix: returns
    (x, nnz(L)+nnz(U), P, Q).  Exposed so that the ordering/pivoting can be tested without a GPU."""
    lib = load_library()
    n = rowptr.size - 1
    rp, ci = np.ascontiguousarray(rowptr, np.int32), np.ascontiguousarray(col, np.int32)
    v, c = np.ascontiguousarray(val, np.float64), np.ascontiguousarray(xy, np.float64)
    x = np.array(rhs, dtype=np.float64, copy=True)
    p, q = np.zeros(n, np.int32), np.zeros(n, np.int32)
    nz = lib.ocp_host_lu_probe(n, _hp(rp), _hp(ci), _hp(v), _hp(c), _hp(x), _hp(p), _hp(q))
    if nz < 0:
        raise OcpError(f"host LU analysis failed ({ERRORS.get(nz, nz)})")
    return x, int(nz), p, q


def host_mf_probe(rowptr, col, val, xy, kind, rhs, pivot_window: int = 1):
    """Multifrontal symbolic analysis + host restatement of its numeric phase: returns (x, stats dict).
    pivot_window=1 is the GPU kernels' static pivoting; a large window = restricted partial pivoting."""
    lib = load_library()
    lib.ocp_host_mf_set_pivot_window(int(pivot_window))
    n = rowptr.size - 1
    rp, ci = np.ascontiguousarray(rowptr, np.int32), np.ascontiguousarray(col, np.int32)
    v, c = np.ascontiguousarray(val, np.float64), np.ascontiguousarray(xy, np.float64)
    kd = np.ascontiguousarray(kind, np.uint8)
    x = np.array(rhs, dtype=np.float64, copy=True)
    st = np.zeros(8)
    rc = lib.ocp_host_mf_probe(n, _hp(rp), _hp(ci), _hp(v), _hp(c), _hp(kd), _hp(x), _hp(st))
    if rc < 0:
        raise OcpError(f"multifrontal host analysis failed ({ERRORS.get(rc, rc)})")
    keys = ["fronts", "levels", "max_front", "max_pivot_block", "flops", "min_pivot", "workspace"]
    return x, dict(zip(keys, st[:7]))
