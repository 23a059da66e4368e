"""B200-native hot path of the Navier-Stokes / buoy-ODE optimal-control loop.

Scope: the forward / adjoint Taylor-Hood assembly, the buoy ODE
sweeps with point location, the point-source scatter and the cost / gradient
reductions of OCP_dolfin.py:201-295, 309-450, as CUDA (sm_100a) kernels behind the
C ABI declared in include/ocp_b200.h, plus the host-side mirror of the reference's
function-level API (``pipeline``).
"""
from . import h5lite, mesh, fespace  # noqa: F401

__all__ = ["h5lite", "mesh", "fespace"]
