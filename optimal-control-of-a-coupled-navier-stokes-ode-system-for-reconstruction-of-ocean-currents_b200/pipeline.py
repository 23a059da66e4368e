"""Host-side mirror of the reference's gradient-descent pipeline.

The reference is three scripts whose "API" is a handful of module-level functions and inline
blocks.  ``OCP`` keeps the state those scripts keep in module globals
(OCP_dolfin.py:63-196: parameters, K, alpha*K, mesh/space, u_d, start points, centre) and exposes
the same operations under the same names with the same array layouts:

    forward_solve(f)                          OCP_dolfin.py:315-325      -> State
    project_grad(w)                           OCP_dolfin.py:328-329
    solve_primal_ode(wSol, buoy_mask)         OCP_dolfin.py:201-230      -> x, u_values_array  (K,nt,2) numpy
    solve_adjoint_ode(wSol, grad_u, x, buoy_mask, u_values_array)  OCP_dolfin.py:234-252  -> mu
    adjoint_solve(w, x, u_values, mu)         OCP_dolfin.py:336-371      -> z
    J(u__, f_)                                OCP_dolfin.py:258-261
    gradient(f, z, df)                        OCP_dolfin.py:379, 388
    grad_test(...)                            OCP_dolfin.py:268-295
    run(...)                                  the loop OCP_dolfin.py:309-450 / Pipeline_limits.py:262-402

All arithmetic happens in libocp_b200.so (CUDA); torch supplies device buffers and, when buoys are
sharded over ranks, the NCCL all-reduce of the single accumulator ``[b | misfit | n_masked]``.
"""
from __future__ import annotations

import json
import os
import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np
import torch

from . import capi
from .fespace import TaylorHood


@dataclass
class Parameters:
    """parameters.json (OCP_dolfin.py:63-69)."""
    viscosity: float = 1.0
    t0: float = 0.0
    T: float = 1.0
    dt: float = 0.005
    alpha: float = 1e-6

    @property
    def nt(self) -> int:
        return int(self.T / self.dt)


def load_parameters(path: str) -> Parameters:
    with open(path, "r") as fh:
        p = json.load(fh)
    return Parameters(viscosity=p["viscosity"], t0=p["t0"], T=p["T"], dt=p["dt"], alpha=p["alpha"])


@dataclass
class Knobs:
    """The module-level constants a user edits in the reference (OCP_dolfin.py:21-48)."""
    num_steps: int = 50
    grad_check: bool = False
    use_line_search: bool = True
    tau: float = 0.5
    c: float = 1e-4
    LR_MIN: float = 1e-6
    LR_MAX: float = 5.0
    conv_crit: float = 1e-3
    exit_rule: str = "half"      # "half": sum(mask) > K/2 (OCP_dolfin.py:448); "ten": > 10 (Pipeline_limits.py:400)


@dataclass
class RunResult:
    J_array: List[float] = field(default_factory=list)
    divs_u: List[float] = field(default_factory=list)
    outer_time: List[float] = field(default_factory=list)
    inner_time: List[float] = field(default_factory=list)
    inner_iterations: List[int] = field(default_factory=list)
    newton_its: List[int] = field(default_factory=list)
    n_masked: List[int] = field(default_factory=list)
    LR: float = 0.0
    exit_reason: str = "num_steps"
    f: Optional[np.ndarray] = None
    grad_tables: Optional[dict] = None
    plots: Optional[dict] = None


class State:
    """A ``Function(W)`` of the reference: the mixed vector on the device (``.vector()`` copies it to host)."""

    def __init__(self, d_w: torch.Tensor):
        self.d_w = d_w

    def vector(self) -> np.ndarray:
        return self.d_w.cpu().numpy()


class OCP:
    def __init__(self, V: TaylorHood, params: Parameters, x0: np.ndarray, u_d: np.ndarray,
                 device: Optional[torch.device] = None, group=None, alpha_scale_K: Optional[int] = None,
                 sort_buoys: bool = True):
        """``x0`` (K,2) start points and ``u_d`` (K,nt,2) of THIS rank's buoys (``u_d=None``: zeros on the device).  With a process group the
        global buoy count is the sum over ranks; alpha is rescaled by the global K (OCP_dolfin.py:76).
        ``sort_buoys`` stores the buoys on the device in a spatially coherent order (sharding.spatial_order);
        every host-facing array keeps the caller's order."""
        if not torch.cuda.is_available():
            raise capi.OcpError("OCP needs a CUDA device: the hot path has no CPU fallback")
        self.V, self.params = V, params
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        self.group = group
        self.world = torch.distributed.get_world_size(group) if group is not None else 1
        self.h, self.nt = float(params.dt), params.nt
        self.center_of_domain = np.array([1.0, 0.5]) if V.mesh.l_shape else np.array([1.0, 1.0])
        self.ctx = capi.Context(V, params.viscosity, params.dt, self.nt, self.center_of_domain)
        if group is not None:
            from .sharding import attach_communicator
            attach_communicator(self.ctx, group)      # NCCL groups: the all-reduce runs inside the library
        x0 = np.ascontiguousarray(x0, np.float64).reshape(-1, 2)
        self.K = x0.shape[0]
        if u_d is not None and u_d.shape != (self.K, self.nt, 2):
            raise ValueError(f"u_d must be ({self.K},{self.nt},2), got {u_d.shape}")
        Kg = self.K
        if group is not None:
            t = torch.tensor([self.K], device=self.device, dtype=torch.int64)
            torch.distributed.all_reduce(t, group=group)
            Kg = int(t.item())
        self.K_global = Kg
        self.alpha = params.alpha * (alpha_scale_K if alpha_scale_K is not None else Kg)
        dev, f64 = self.device, torch.float64
        K, nt, nn, nv, nd = self.K, self.nt, V.num_nodes, V.mesh.num_vertices, V.ndofs
        self.xsarr, self.ysarr = x0[:, 0].copy(), x0[:, 1].copy()
        self.u_d = u_d
        # device order of the buoys: device index i holds the caller's buoy perm[i]
        self.perm = None
        if sort_buoys and K > 32:
            from .sharding import spatial_order
            perm = spatial_order(V, x0)
            if not np.array_equal(perm, np.arange(K)):
                self.perm = torch.from_numpy(perm).to(dev)
                self.inv_perm = torch.empty_like(self.perm)
                self.inv_perm[self.perm] = torch.arange(K, device=dev)
        self.d_x0 = torch.from_numpy(x0).to(dev)
        if self.perm is not None:
            self.d_x0 = self.d_x0[self.perm].contiguous()
        # u_d = None: synthetic sweeps fill the device array themselves (no 32 GB host array for 1e7 buoys)
        self.d_ud = (self._to_time_major(u_d) if u_d is not None
                     else torch.zeros((self.nt, self.K, 2), device=dev, dtype=torch.float64))
        self.d_x = torch.empty((nt, K, 2), device=dev, dtype=f64)
        self.d_u = torch.empty((nt, K, 2), device=dev, dtype=f64)
        self.d_x_ls = self.d_u_ls = None
        self.d_mask = torch.zeros(K, device=dev, dtype=f64)
        self.d_mask_ls = torch.zeros(K, device=dev, dtype=f64)
        self.d_parked = torch.zeros(K, device=dev, dtype=torch.uint8)
        self.d_acc = torch.zeros(2 * nn + 2, device=dev, dtype=f64)
        self.d_w = torch.zeros(nd, device=dev, dtype=f64)
        self.d_w_ls = torch.zeros(nd, device=dev, dtype=f64)
        self.d_z = torch.zeros(nd, device=dev, dtype=f64)
        self.d_g = torch.zeros((nv, 4), device=dev, dtype=f64)
        self.d_vel = torch.zeros((nn, 2), device=dev, dtype=f64)
        self.d_znod = torch.zeros((nn, 2), device=dev, dtype=f64)
        self.d_grad = torch.zeros((nn, 2), device=dev, dtype=f64)
        self.d_f = torch.zeros((nn, 2), device=dev, dtype=f64)
        self.d_f_ls = torch.zeros((nn, 2), device=dev, dtype=f64)
        self.d_sc = torch.zeros(8, device=dev, dtype=f64)
        self.last_newton_its = 0
        self.last_res_hist: List[float] = []

    # ------------------------------------------------------------------ helpers
    def _to_time_major(self, a: np.ndarray) -> torch.Tensor:
        """host (K,nt,2) in the caller's buoy order -> device (nt,K,2) in device order"""
        src = torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(self.device)
        if self.perm is not None and a.shape[0] == self.K:
            src = src[self.perm].contiguous()
        dst = torch.empty((self.nt, a.shape[0], 2), device=self.device, dtype=torch.float64)
        self.ctx.traj_transpose(src, dst, a.shape[0], True)
        return dst

    def _to_reference_layout(self, d: torch.Tensor) -> np.ndarray:
        """device (nt,K,2) in device order -> host (K,nt,2) in the caller's buoy order"""
        K = d.shape[1]
        dst = torch.empty((K, self.nt, 2), device=self.device, dtype=torch.float64)
        self.ctx.traj_transpose(d, dst, K, False)
        if self.perm is not None and K == self.K:
            dst = dst[self.inv_perm]
        return dst.cpu().numpy()

    def _buoy_vector_to_host(self, d: torch.Tensor) -> np.ndarray:
        """per-buoy device vector (mask, parked) -> host array in the caller's order"""
        if self.perm is not None and d.shape[0] == self.K:
            d = d[self.inv_perm]
        return d.cpu().numpy()

    def _cells_to_host(self, d_cell: torch.Tensor) -> np.ndarray:
        """device (nt,K) cell indices -> host (K,nt) in the caller's buoy order"""
        c = d_cell.t()
        if self.perm is not None and c.shape[0] == self.K:
            c = c[self.inv_perm]
        return c.contiguous().cpu().numpy()

    def _buoy_vector_to_dev(self, a) -> torch.Tensor:
        t = self._dev(np.asarray(a, np.float64))
        if self.perm is not None and t.shape[0] == self.K:
            t = t[self.perm].contiguous()
        return t

    def _dev(self, a) -> torch.Tensor:
        if isinstance(a, torch.Tensor):
            return a
        return torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(self.device)

    def _allreduce(self, t: torch.Tensor):
        from .sharding import allreduce_accumulator
        allreduce_accumulator(t, self.group, self.ctx)

    def set_control(self, f_nodal: np.ndarray):
        self.d_f.copy_(self._dev(f_nodal))

    # ------------------------------------------------------- reference-named API
    def forward_solve(self, f, w0: Optional[State] = None, out: Optional[torch.Tensor] = None) -> State:
        """``solve(F == 0, w, bcs)`` with ``w = Function(W)`` zero-initialised (OCP_dolfin.py:315-325);
        pass ``w0`` to warm-start and overwrite it like grad_test does (OCP_dolfin.py:274)."""
        d_f = self._dev(f)
        if w0 is not None:
            d_w, zero = w0.d_w, False
        else:
            d_w, zero = (out if out is not None else torch.empty_like(self.d_w)), True
        self.last_newton_its, self.last_res_hist = self.ctx.forward_solve(d_f, d_w, zero)
        return State(d_w)

    def project_grad(self, w: State, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``project(grad(w.sub(0)), V_vec)`` (OCP_dolfin.py:328-329) -> device tensor (nv,4)."""
        d_g = out if out is not None else torch.empty_like(self.d_g)
        self.ctx.project_grad(w.d_w, d_g)
        return d_g

    def solve_primal_ode(self, wSol: State, buoy_mask: np.ndarray):
        """OCP_dolfin.py:201-230.  ``buoy_mask`` (K) float array is mutated in place; returns numpy
        ``x, u_values_array`` of shape (K,nt,2)."""
        self._primal(wSol.d_w, self.d_x, self.d_u, self.d_mask, init_mask=buoy_mask)
        buoy_mask[:] = self._buoy_vector_to_host(self.d_mask)
        return self._to_reference_layout(self.d_x), self._to_reference_layout(self.d_u)

    def last_parked(self) -> np.ndarray:
        """Per-buoy flags of the last solve_primal_ode: 1 where only the final sample left the domain."""
        return self._buoy_vector_to_host(self.d_parked)

    def solve_adjoint_ode(self, wSol: State, grad_u: torch.Tensor, x, buoy_mask, u_values_array) -> np.ndarray:
        """OCP_dolfin.py:234-252 -> numpy mu (K,nt,2).  (The fused sweep also deposits point sources and the
        misfit into a scratch accumulator, which this reference-shaped call discards.)"""
        d_x, d_u = self._to_time_major(x), self._to_time_major(u_values_array)
        d_mask = self._buoy_vector_to_dev(buoy_mask)
        d_mu = torch.empty_like(d_x)
        acc = torch.zeros_like(self.d_acc)
        self.ctx.velocity_nodal(wSol.d_w, self.d_vel)
        parked = torch.zeros(x.shape[0], device=self.device, dtype=torch.uint8)
        self.ctx.buoy_adjoint_scatter(self.d_vel, grad_u, x.shape[0], d_x, d_u, self.d_ud, d_mask, parked, d_mu, acc)
        return self._to_reference_layout(d_mu)

    def adjoint_solve(self, w: State, x, u_values, buoy_mask, grad_u: Optional[torch.Tensor] = None,
                      parked=None) -> State:
        """The adjoint PDE block OCP_dolfin.py:336-371 for host arrays x, u_values (K,nt,2): recomputes mu,
        deposits the point sources, all-reduces, assembles aAdj, applies the BC and solves.  ``parked`` (K) flags
        the buoys whose LAST sample alone left the domain (OCP_dolfin.py:226-229; ``last_parked()`` after
        solve_primal_ode); without it the flags are inferred from the stored arrays (last sample at the centre with
        zero velocity), which is what the reference's scatter loop sees."""
        d_x, d_u = self._to_time_major(x), self._to_time_major(u_values)
        d_mask = self._buoy_vector_to_dev(buoy_mask)
        g = grad_u if grad_u is not None else self.project_grad(w)
        self.ctx.velocity_nodal(w.d_w, self.d_vel)
        if parked is not None:
            parked = self._buoy_vector_to_dev(np.asarray(parked, np.float64)).to(torch.uint8)
        else:
            parked = ((d_x[-1, :, 0] == self.center_of_domain[0]) & (d_x[-1, :, 1] == self.center_of_domain[1])
                      & (d_u[-1, :, 0] == 0) & (d_u[-1, :, 1] == 0) & (d_mask == 0)).to(torch.uint8)
        self.d_acc.zero_()
        self.ctx.buoy_adjoint_scatter(self.d_vel, g, x.shape[0], d_x, d_u, self.d_ud, d_mask, parked, None, self.d_acc)
        self._allreduce(self.d_acc)
        z = torch.empty_like(self.d_z)
        self.ctx.adjoint_solve(w.d_w, self.d_acc, z)
        return State(z)

    def J(self, u__, f_) -> float:
        """OCP_dolfin.py:258-261; ``u__`` (K,nt,2) numpy or a time-major device tensor."""
        d_u = u__ if isinstance(u__, torch.Tensor) else self._to_time_major(u__)
        return self._cost(d_u, self._dev(f_))

    def gradient(self, f, z: State, df) -> float:
        """``assemble(inner(alpha*f - zSol, df)*ds(1))`` (OCP_dolfin.py:379, 388)."""
        d_f, d_df = self._dev(f), self._dev(df)
        self.ctx.velocity_nodal(z.d_w, self.d_znod)
        self.ctx.nodal_axpby(self.alpha, d_f, -1.0, self.d_znod, self.d_grad)
        self.ctx.boundary_inner(self.d_grad, d_df, self.d_sc)
        return float(self.d_sc[0].item())

    # ----------------------------------------------------------- device-resident path
    def _primal(self, d_w, d_x, d_u, d_mask, init_mask=None, d_cell=None):
        if init_mask is None:
            d_mask.zero_()
        else:
            d_mask.copy_(self._buoy_vector_to_dev(init_mask))
        self.ctx.velocity_nodal(d_w, self.d_vel)
        self.ctx.buoy_forward(self.d_vel, self.d_x0, self.K, d_x, d_u, d_cell, d_mask, self.d_parked)

    def _cost(self, d_u: torch.Tensor, d_f: torch.Tensor) -> float:
        self.d_sc.zero_()
        self.ctx.misfit(d_u.shape[1], d_u, self.d_ud, self.d_sc)
        self._allreduce(self.d_sc[0:1])
        self.ctx.boundary_inner(d_f, d_f, self.d_sc[1:])
        sc = self.d_sc[:2].cpu().numpy()
        return float(sc[0] + 0.5 * self.alpha * sc[1])

    def gradient_step(self, d_f: torch.Tensor):
        """forward NS, projection, primal ODE, backward sweep (+ all-reduce), adjoint NS - the "outer" block
        of one gradient-descent iteration (OCP_dolfin.py:313-371).  Results stay on the device."""
        if self.group is not None and self.ctx.comm_size() <= 1 and self.world > 1:
            # process groups that are not NCCL (gloo in the CPU tests): the exchange goes through torch.distributed,
            # so the block is issued piecewise
            its, hist = self.ctx.forward_solve(d_f, self.d_w, True)
            self.last_newton_its, self.last_res_hist = its, hist
            self.ctx.project_grad(self.d_w, self.d_g)
            self._primal(self.d_w, self.d_x, self.d_u, self.d_mask)
            self.d_acc.zero_()
            self.ctx.buoy_adjoint_scatter(self.d_vel, self.d_g, self.K, self.d_x, self.d_u, self.d_ud, self.d_mask,
                                          self.d_parked, None, self.d_acc)
            self._allreduce(self.d_acc)
            self.ctx.adjoint_solve(self.d_w, self.d_acc, self.d_z)
            self.ctx.velocity_nodal(self.d_z, self.d_znod)
            self.ctx.nodal_axpby(self.alpha, d_f, -1.0, self.d_znod, self.d_grad)     # grad j = alpha f - z
            return
        # one library call = one CUDA graph replay once warm (ocp_gradient_device), collective included
        self.last_newton_its = self.ctx.gradient_device(d_f, self.d_x0, self.d_ud, self.K, self.d_w, self.d_g, self.d_vel,
                                                        self.d_x, self.d_u, self.d_mask, self.d_parked, self.d_acc,
                                                        self.d_z, self.d_znod, self.d_grad, self.alpha)
        self.last_res_hist = self.ctx.newton_history()

    def _cost_from_acc(self, d_f) -> float:
        self.ctx.boundary_inner(d_f, d_f, self.d_sc)
        nn = self.V.num_nodes
        v = torch.stack([self.d_acc[2 * nn], self.d_sc[0]]).cpu().numpy()
        return float(v[0] + 0.5 * self.alpha * v[1])

    def grad_test(self, w: State, J0: float, gradj: float, df, buoy_mask, f=None, out_dir: Optional[str] = None,
                  iter: int = 0) -> dict:
        """OCP_dolfin.py:268-295: one-sided and centred finite differences for h_ = 1e-1 .. 1e-8.  ``w`` is
        warm-started and overwritten exactly as in the reference.  Returns the two tables (and writes the
        reference's ``grad_J_error_{iter}.txt`` / ``grad_J_error_centered_{iter}.txt`` when out_dir is given)."""
        d_f = self.d_f if f is None else self._dev(f)
        d_df = self._dev(df)
        fpert = torch.empty_like(d_f)
        mask = self._buoy_vector_to_dev(buoy_mask).clone()
        rows1, rows2 = [], []

        def Jat(hh):
            self.ctx.nodal_axpby(1.0, d_f, hh, d_df, fpert)
            self.forward_solve(fpert, w0=w)
            self.ctx.velocity_nodal(w.d_w, self.d_vel)
            self.ctx.buoy_forward(self.d_vel, self.d_x0, self.K, self.d_x, self.d_u, None, mask, self.d_parked)
            return self._cost(self.d_u, fpert)

        for k in range(1, 9):
            h_ = 10 ** (-k)
            ga = (Jat(h_) - J0) / h_
            rows1.append((gradj, ga, abs(ga - gradj), h_))
        for k in range(1, 9):
            h_ = 10 ** (-k)
            jr = Jat(h_)
            jl = Jat(-h_)
            ga = (jr - jl) / (2 * h_)
            rows2.append((gradj, ga, abs(gradj - ga), h_))
        buoy_mask[:] = self._buoy_vector_to_host(mask)
        if out_dir is not None:
            os.makedirs(out_dir, exist_ok=True)
            hdr = "reduced Gradient j \t \t approximated gradient J \t Error \t \t \t h_i \n"
            for name, rows in ((f"grad_J_error_{iter}.txt", rows1), (f"grad_J_error_centered_{iter}.txt", rows2)):
                with open(os.path.join(out_dir, name), "w") as fh:
                    fh.write(hdr)
                    for r in rows:
                        fh.write(f" {r[0]} \t {r[1]} \t {r[2]} \t {r[3]} \n")
        return {"one_sided": rows1, "centered": rows2}

    def run(self, f0: np.ndarray, knobs: Knobs = Knobs(), df_check=None, LR: Optional[float] = None,
            out_dir: Optional[str] = None, callback: Optional[Callable] = None) -> RunResult:
        """The gradient-descent loop OCP_dolfin.py:309-450 (Pipeline_limits.py:262-402 with
        ``use_line_search=False, exit_rule="ten"``), quirks of SURVEY App. A.7 included."""
        kn = knobs
        res = RunResult()
        LR = kn.LR_MAX if LR is None else LR
        self.set_control(f0)
        d_f = self.d_f
        nn = self.V.num_nodes
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for i in range(kn.num_steps):
            t0 = time.time()
            ev[0].record()
            self.gradient_step(d_f)
            ev[1].record()
            torch.cuda.synchronize()
            res.outer_time.append(time.time() - t0)
            res.newton_its.append(self.last_newton_its)
            nmask = int(round(float(self.d_acc[2 * nn + 1].item())))
            res.n_masked.append(nmask)
            if kn.grad_check and i == 0:
                df = np.full((nn, 2), 0.1) if df_check is None else df_check
                d_df = self._dev(df)
                self.ctx.boundary_inner(self.d_grad, d_df, self.d_sc)
                gradj = float(self.d_sc[0].item())
                J0 = self._cost_from_acc(d_f)
                u_keep, x_keep = self.d_u.clone(), self.d_x.clone()
                wtmp = State(self.d_w.clone())
                mask = self._buoy_vector_to_host(self.d_mask)
                res.grad_tables = self.grad_test(wtmp, J0, gradj, df, mask, f=d_f, out_dir=out_dir, iter=i)
                res.grad_tables.update(J0=J0, gradj=gradj)
                self.d_u.copy_(u_keep)
                self.d_x.copy_(x_keep)
            t1 = time.time()
            inner = 0
            if kn.use_line_search:
                # df = -(alpha f - z);  gradj = int (alpha f - z).df = -||alpha f - z||^2
                self.ctx.boundary_inner(self.d_grad, self.d_grad, self.d_sc)
                gradj = -float(self.d_sc[0].item())
                cond = -kn.c * gradj
                J_old = self._cost_from_acc(d_f)
                if self.d_x_ls is None:
                    self.d_x_ls, self.d_u_ls = torch.empty_like(self.d_x), torch.empty_like(self.d_u)
                while True:
                    inner += 1
                    self.ctx.nodal_axpby(1.0, d_f, -LR, self.d_grad, self.d_f_ls)    # f + LR*df, df = -(alpha f - z)
                    self.ctx.forward_solve(self.d_f_ls, self.d_w_ls, True)
                    self._primal(self.d_w_ls, self.d_x_ls, self.d_u_ls, self.d_mask_ls)
                    J_new = self._cost(self.d_u_ls, self.d_f_ls)
                    if J_old - J_new >= LR * cond:
                        break
                    if LR <= kn.LR_MIN:
                        # the reference loops forever here (latent hazard); stop instead
                        res.exit_reason = "line_search_stalled"
                        break
                    LR = max(kn.tau * LR, kn.LR_MIN)
            torch.cuda.synchronize()
            res.inner_time.append(time.time() - t1)
            res.inner_iterations.append(inner)
            # control update f <- f - LR (alpha f - z), OCP_dolfin.py:426
            self.ctx.nodal_axpby(1.0, d_f, -LR, self.d_grad, d_f)
            # J_array uses the OLD velocities with the NEW control (OCP_dolfin.py:426-429)
            res.J_array.append(self._cost_from_acc(d_f))
            self.ctx.field_norms(self.d_w, self.d_sc)
            res.divs_u.append(float(np.sqrt(self.d_sc[0].item())))
            if out_dir is not None:
                # control checkpoint of this iteration, appended as group f_i like the reference's
                # write_checkpoint(..., append=True) (OCP_dolfin.py:440-441); resume with checkpoint.read_control,
                # which returns the last group.  A non-finite cost/control never replaces the last good checkpoint.
                from . import checkpoint
                os.makedirs(os.path.join(out_dir, "checkpoints"), exist_ok=True)
                f_host = d_f.cpu().numpy()
                if np.isfinite(res.J_array[-1]) and np.all(np.isfinite(f_host)):
                    checkpoint.write_control(os.path.join(out_dir, "checkpoints", "q.h5"), self.V, f_host, append=i > 0)
                else:
                    res.exit_reason = "non-finite cost or control"
                    break
            if callback is not None:
                callback(i, self, res)
            if res.exit_reason == "line_search_stalled":
                break
            if i > 5 and abs(res.J_array[i] - res.J_array[i - 1]) < kn.conv_crit:
                res.exit_reason = "cost small enough"
                break
            limit = self.K_global / 2 if kn.exit_rule == "half" else 10
            if nmask > limit:
                res.exit_reason = "too many buoys out of domain"
                break
        res.LR = LR
        res.f = d_f.cpu().numpy()
        if out_dir is not None:
            self.save_run(res, out_dir, kn)
        return res

    def save_run(self, res: RunResult, out_dir: str, kn: Knobs):
        """The files the reference writes at the end of a run, same names and text formats: ``timings.txt``
        (OCP_dolfin.py:476-482), ``q_backup/q.xdmf`` (485-486), ``u_divergence.txt`` (489-492), ``variables.txt``
        (495-507), ``J_array.npy`` (510-511), ``paraview/checkpoint/{u,p}.xdmf`` (578-588), and the figures through
        ``report.save_plots`` (455-575; PNGs need matplotlib, ``J.svg`` is always written)."""
        from . import checkpoint
        os.makedirs(os.path.join(out_dir, "q_backup"), exist_ok=True)
        with open(os.path.join(out_dir, "timings.txt"), "w") as fh:
            for k, it in enumerate(res.inner_iterations):
                fh.write(f"Iteration {k}:\n")
                fh.write(f"  outer loop time: {res.outer_time[k]:.6f} seconds\n")
                fh.write(f"  inner loop time: {res.inner_time[k]:.6f} seconds\n")
                fh.write(f"  inner loop iterations: {it}\n")
                fh.write("-" * 40 + "\n")
        checkpoint.write_control(os.path.join(out_dir, "q_backup", "q.h5"), self.V, res.f)
        with open(os.path.join(out_dir, "u_divergence.txt"), "w") as fh:
            for i, dv in enumerate(res.divs_u):
                fh.write("div(u) \t \t \t i  \n")
                fh.write(f" {dv} \t {i} \n")
        nx = int(round(np.sqrt(self.V.mesh.num_cells / 2))) if not self.V.mesh.l_shape else self.V.mesh.num_cells
        with open(os.path.join(out_dir, "variables.txt"), "w") as fh:
            fh.write("mesh resolution: %s \n" % nx)
            fh.write("ud type: %s \n" % ("L-shape" if self.V.mesh.l_shape else "custom_ud"))
            fh.write("t0: %s \n" % self.params.t0)
            fh.write("T: %s \n" % self.params.T)
            fh.write("dt: %s \n" % self.params.dt)
            fh.write("viscosity: %s \n" % self.params.viscosity)
            fh.write("buoy count: %s \n" % self.K_global)
            fh.write("LR: %s \n" % res.LR)
            fh.write("LR_MAX: %s \n" % kn.LR_MAX)
            fh.write("LR_MIN: %s \n" % kn.LR_MIN)
            fh.write("conv. crit.: %s \n" % kn.conv_crit)
            fh.write("gradient descent steps: %s \n" % kn.num_steps)
        np.save(os.path.join(out_dir, "J_array.npy"), np.array(res.J_array))
        checkpoint.write_state(os.path.join(out_dir, "paraview", "checkpoint"), self.V, self.d_w.cpu().numpy())
        # figures of OCP_dolfin.py:455-575 (J.svg always; the PNGs when matplotlib is installed)
        from . import report
        res.plots = report.save_plots(self, res, out_dir)

    def field_norms(self, w: State):
        """(||div u||, ||u||_L2, ||u||_H1) - OCP_dolfin.py:430, Pipeline_limits.py:433-443."""
        self.ctx.field_norms(w.d_w, self.d_sc)
        d, l2, h1 = self.d_sc[:3].cpu().numpy()
        return float(np.sqrt(d)), float(np.sqrt(l2)), float(np.sqrt(l2 + h1))

    def error_norms(self, w: State, w_ref) -> tuple:
        """L2 and H1 norm of ``u - u_ref`` - the ``norm_table.txt`` of Pipeline_limits.py:433-443 (``w_ref`` is a W
        vector, e.g. the stored u_bar of reference_runs/u_bar_chapter_6.3.3)."""
        d = w.d_w - self._dev(w_ref)
        self.ctx.field_norms(d, self.d_sc)
        _, l2, h1 = self.d_sc[:3].cpu().numpy()
        return float(np.sqrt(l2)), float(np.sqrt(l2 + h1))

    def close(self):
        self.ctx.close()


# -- the reference's own L-shape experiment --------------------------------------------------------------------
def lshape_reference_observations(params: Parameters = Parameters()):
    """``ud_type == "L-shape"`` of OCP_dolfin.py:163-196: K = 3 buoys started at (0.5, 0.5), (1.0, 0.5), (1.5, 1.0)
    with the analytic target velocities u_d(t) sampled on ``time_interval = np.linspace(t0, T, int(T / h))`` - whose
    spacing is 1/199, not h (kept).  Returns ``x0`` (3,2) and ``u_d`` (3,nt,2)."""
    nt = params.nt
    time_interval = np.linspace(params.t0, params.T, nt)
    ud1 = 0.5 * (np.cos(np.pi * (time_interval - 0.5)) - 1 - np.cos(np.pi))
    ud2 = ud1.copy()
    x0 = np.array([[0.5, 0.5], [1.0, 0.5], [1.5, 1.0]])
    u_d = np.zeros((3, nt, 2))
    u_d[0, :, 0] = ud1
    u_d[1, :, 0], u_d[1, :, 1] = ud1, ud2
    u_d[2, :, 1] = ud2
    return x0, u_d


# -- initial controls of the three pipelines -------------------------------------------------------------------
def initial_control(V: TaylorHood, pipeline: str = "OCP", case: int = 0) -> np.ndarray:
    """q_0 as a P2 nodal field.  "OCP": sinusoidal Expression(degree=1) (OCP_dolfin.py:143-145);
    "PL": constant (0.1, 0) (Pipeline_limits.py:123); "ICT": `case` 0-3 (initial_control_test.py:30-43)."""
    pi = np.pi
    sin_x = lambda x, y: -np.cos(pi * x) * np.sin(pi * y)
    sin_y = lambda x, y: np.sin(pi * x) * np.cos(pi * y)
    const = lambda v: (lambda x, y: v + 0.0 * x)
    if pipeline == "OCP":
        return V.interpolate_control(sin_x, sin_y, 1)
    if pipeline == "PL":
        return V.interpolate_control(const(0.1), const(0.0), 2)
    if pipeline == "ICT":
        return _ict_case(V, case)
    raise ValueError(pipeline)


def _ict_case(V: TaylorHood, case: int) -> np.ndarray:
    """initial_control_test.py:30-43 (all four are Expression(degree=1))."""
    pi = np.pi
    const = lambda v: (lambda x, y: v + 0.0 * x)
    a = lambda x, y: -np.cos(pi * x) * np.sin(pi * y)
    b = lambda x, y: np.sin(pi * x) * np.cos(pi * y)
    if case == 0:
        return V.interpolate_control(a, b, 1)
    if case == 1:
        return V.interpolate_control(const(0.0), const(0.0), 1)
    if case == 2:
        return V.interpolate_control(b, a, 1)
    return V.interpolate_control(const(0.1), const(0.1), 1)
