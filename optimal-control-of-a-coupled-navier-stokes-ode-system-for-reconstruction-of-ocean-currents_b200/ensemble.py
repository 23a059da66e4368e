"""Independent optimal-control cases run concurrently on ONE GPU (BASELINE.json cfg4: initial_control_test.py cases
0-3 plus the Pipeline_limits.py buoy-count sweep; the reference runs them as separate serial scripts).

A gradient-descent iteration on the 32 x 32 mesh is latency bound - the sparse LU walks a dependent chain of small
fronts and leaves most of the 148 SMs idle - so independent cases are given one context, one CUDA stream and one host
thread each and overlap on the device.  Every case goes through the C-ABI host-buffer entry point
``ocp_gradient_host`` (forward Navier-Stokes, projection, primal buoy ODE, backward sweep, adjoint Navier-Stokes); the
host side only forms ``alpha f - z`` and the cost.  No collective: cases never exchange data."""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from . import capi
from .fespace import TaylorHood


@dataclass
class Case:
    name: str
    x0: np.ndarray          # (K, 2) start points
    u_d: np.ndarray         # (K, nt, 2) observations
    f0: np.ndarray          # (nn, 2) initial control
    viscosity: float = 1.0
    alpha: float = 1e-6     # per buoy (multiplied by K like OCP_dolfin.py:76)


class Ensemble:
    def __init__(self, V: TaylorHood, cases: List[Case], dt: float = 0.005, nt: int = 200,
                 device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise capi.OcpError("Ensemble needs a CUDA device: the hot path has no CPU fallback")
        self.V, self.cases = V, cases
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        center = (1.0, 0.5) if V.mesh.l_shape else (1.0, 1.0)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in cases]
        # Programmatic dependent launches between the tree levels of the triangular solves shorten the latency of ONE
        # evaluation (2.87 vs 3.04 ms), but the next level's CTAs become resident - a full SM each - while they wait;
        # with several contexts saturating the GPU that costs throughput (measured on B200, the eight cfg4 cases:
        # 814 vs 903 GD iterations/s).  Ensemble contexts are therefore created with plain launches unless the
        # caller has set the switch (it is read when a context is created).
        pdl_prev = os.environ.get("OCP_MF_PDL")
        if pdl_prev is None:
            os.environ["OCP_MF_PDL"] = "0"
        try:
            self.ctx = [capi.Context(V, c.viscosity, dt, nt, center, stream=s.cuda_stream)
                        for c, s in zip(cases, self.streams)]
        finally:
            if pdl_prev is None:
                del os.environ["OCP_MF_PDL"]
        self.out = []
        for c, ctx in zip(cases, self.ctx):
            ctx.set_observations_host(c.x0, c.u_d)
            pin = lambda n: torch.empty(n, dtype=torch.float64).pin_memory().numpy()
            self.out.append((pin(V.ndofs), pin(V.ndofs), pin(c.x0.shape[0]), pin(4)))
        self.pool = ThreadPoolExecutor(max_workers=len(cases))

    def _one(self, i: int, f: np.ndarray, copy: bool = True):
        torch.cuda.set_device(self.device)            # the CUDA context is per host thread
        c = self.cases[i]
        w, z, mask, sc = self.ctx[i].gradient_host(f, self.out[i])
        alpha = c.alpha * c.x0.shape[0]
        grad = alpha * f - self.V.velocity_nodal(z)
        if copy:      # the pinned staging buffers are overwritten by the next call: hand out copies
            w, z = w.copy(), z.copy()
        return dict(J=sc["misfit"] + 0.5 * alpha * sc["f_norm2"], grad=grad, newton_its=sc["newton_its"],
                    n_masked=sc["n_masked"], w=w, z=z)

    def gradients(self, controls: Optional[List[np.ndarray]] = None, concurrent: bool = True) -> List[dict]:
        """One gradient evaluation per case; ``concurrent=False`` runs them one after the other (same results)."""
        fs = controls if controls is not None else [c.f0 for c in self.cases]
        if not concurrent:
            return [self._one(i, f) for i, f in enumerate(fs)]
        futs = [self.pool.submit(self._one, i, f) for i, f in enumerate(fs)]
        return [f.result() for f in futs]

    def descend(self, steps: int, lr: float = 5.0, concurrent: bool = True) -> List[List[float]]:
        """`steps` plain gradient-descent iterations per case (no line search, Pipeline_limits.py defaults);
        returns the cost history of every case."""
        fs = [c.f0.copy() for c in self.cases]
        hist: List[List[float]] = [[] for _ in self.cases]

        def run(i):
            torch.cuda.set_device(self.device)
            for _ in range(steps):
                r = self._one(i, fs[i], copy=False)
                hist[i].append(r["J"])
                fs[i] = fs[i] - lr * r["grad"]

        if concurrent:
            for f in [self.pool.submit(run, i) for i in range(len(self.cases))]:
                f.result()
        else:
            for i in range(len(self.cases)):
                run(i)
        return hist

    def close(self):
        self.pool.shutdown(wait=True)
        for c in self.ctx:
            c.close()
        self.ctx = []
