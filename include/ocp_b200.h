/*
 * ocp_b200.h - C ABI of the B200-native hot path of the Navier-Stokes / buoy-ODE
 * optimal-control loop (libocp_b200.so, sm_100a).
 *
 * The reference (legacy-FEniCS scripts) has no FFI; its "interface" for this path is
 * the set of module-level functions and inline blocks of OCP_dolfin.py (identical in
 * Pipeline_limits.py / initial_control_test.py).  Every entry point below names the
 * reference lines it replaces.  Conventions:
 *   - plain C, opaque context, int return code (0 = OK, negative = error;
 *     ocp_last_error() gives the text), no exceptions cross the boundary;
 *   - all `d_*` pointers are DEVICE pointers owned by the caller, `h_*` are HOST
 *     pointers owned by the caller; nothing is retained after a call returns
 *     except what ocp_create copied;
 *   - one context = one GPU = one stream; a context is not thread-safe;
 *   - all arithmetic is fp64, all indices int32.
 *
 * Device array layouts
 *   W vector        (ndofs)        the mixed [P2]^2 x P1 coefficient vector in the caller's dof numbering
 *   nodal field     (nn, 2)        P2 nodal vector field, node-major (x, y) pairs: velocity, control f, adjoint z
 *   grad field      (nv, 4)        continuous P1 tensor, vertex-major [g00 g01 g10 g11]
 *   trajectories    (nt, K, 2)     TIME-major (coalesced across buoys).  The reference's (K, nt, 2)
 *                                  layout exists only in the *_host entry points / ocp_traj_transpose.
 *   accumulator     (2*nn + 2)     [nodal point-source vector b (nn,2) | misfit | n_masked]  - the single
 *                                  buffer that is all-reduced across GPUs when buoys are sharded.
 */
#ifndef OCP_B200_H_
#define OCP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCP_OK 0
#define OCP_ERR_INVALID (-1)
#define OCP_ERR_CUDA (-2)
#define OCP_ERR_SOLVER (-3)
#define OCP_ERR_NOT_CONVERGED (-4)
#define OCP_ERR_NO_DEVICE (-5)
#define OCP_ERR_COMM (-6)

typedef struct ocp_ctx ocp_ctx;

/* Problem tables: everything the reference sets up at OCP_dolfin.py:79-148 (mesh, W, boundary
 * markers, Dirichlet BC, ds(1)) plus parameters.json (OCP_dolfin.py:63-69). All pointers are HOST. */
typedef struct {
    int32_t nv;          /* mesh vertices (= P1 dofs)                        */
    int32_t nn;          /* P2 nodes = vertices + edges                      */
    int32_t nc;          /* cells                                            */
    int32_t ndofs;       /* 2*nn + nv                                        */
    int32_t nnz;         /* CSR non-zeros of the W x W pattern               */
    int32_t n_dirichlet; /* constrained dofs of DirichletBC(W.sub(0),(0,0))  */
    int32_t n_g1;        /* facets of Gamma_1 (marker 1)                     */
    int32_t nt;          /* int(T/dt) samples per trajectory (200)           */
    const double *cell_geom;      /* (nc,6) x0 y0 a1 b1 a2 b2: lambda_1 = a1 (x-x0) + b1 (y-y0), lambda_2 likewise */
    const int32_t *cell_nodes;    /* (nc,6) v0 v1 v2 e0 e1 e2 (vertices sorted ascending, e_i opposite v_i)         */
    const int32_t *cell_nbr;      /* (nc,3) cell across the edge opposite local vertex i, -1 on the boundary      */
    const double *node_coords;    /* (nn,2)                                   */
    const int32_t *dof_ux;        /* (nn) W dof of u_x at a node               */
    const int32_t *dof_uy;        /* (nn)                                      */
    const int32_t *dof_p;         /* (nv)                                      */
    const int32_t *csr_rowptr;    /* (ndofs+1)                                 */
    const int32_t *csr_col;       /* (nnz) sorted within each row              */
    const int32_t *dirichlet_dofs;/* (n_dirichlet)                             */
    const int32_t *g1_nodes;      /* (n_g1,3) [va vb mid] node ids             */
    const double *g1_len;         /* (n_g1)                                    */
    const double *g1_normal;      /* (n_g1,2) outward unit normal              */
    double bin_ox, bin_oy, bin_ihx, bin_ihy; /* point-location bins            */
    int32_t nbx, nby;
    const int32_t *bin_ptr;       /* (nbx*nby+1)                               */
    const int32_t *bin_cells;     /* candidates per bin, ascending             */
    double viscosity;             /* parameters.json "viscosity"               */
    double dt;                    /* parameters.json "dt" (h)                  */
    double center_x, center_y;    /* center_of_domain, OCP_dolfin.py:81, 98    */
} ocp_problem_desc;

/* Per-phase device time of the last call sequence, milliseconds (CUDA events). */
typedef struct {
    double assemble_ms;   /* cell + facet assembly, BC rows, residual norm          */
    double factor_ms;     /* sparse LU numeric refactorisation                      */
    double solve_ms;      /* triangular solves                                      */
    double analyse_ms;    /* one-time host symbolic analysis + first factorisation  */
    int32_t n_factor, n_solve;
    int32_t n_dense;      /* solves replaced by a precomputed dense operator (Stokes step, mass matrix)   */
    int32_t reserved;
} ocp_solver_stats;

int ocp_version(void);
/* 0 if a CUDA device is usable, OCP_ERR_NO_DEVICE otherwise (never falls back to the CPU). */
int ocp_device_available(void);

/* `stream` is a cudaStream_t (NULL = legacy default stream). */
int ocp_create(const ocp_problem_desc *desc, void *stream, ocp_ctx **out);
void ocp_destroy(ocp_ctx *ctx);
const char *ocp_last_error(const ocp_ctx *ctx);
void ocp_get_solver_stats(const ocp_ctx *ctx, ocp_solver_stats *out);
void ocp_reset_solver_stats(ocp_ctx *ctx);
void ocp_set_viscosity(ocp_ctx *ctx, double viscosity);
/* Replace the Dirichlet set of the context: rows h_dofs (n) are constrained to h_vals (NULL = 0).  The OCP
 * pipelines use the homogeneous velocity BC given to ocp_create (OCP_dolfin.py:136); the twin experiment that produced
 * the reference's u_d data constrains the whole boundary and one pressure line with non-zero data
 * (plotting/ud_construction_pipeline.py:95-106). */
int ocp_set_dirichlet(ocp_ctx *ctx, const int32_t *h_dofs, const double *h_vals, int n);
/* Reproducible point-source deposit (also: environment OCP_DETERMINISTIC=1).  Off (default): the deposits of
 * OCP_dolfin.py:353-366 are summed with fp64 atomics, so the last bits of b change from run to run with the arrival
 * order of the threads.  On: every deposit is split exactly into fixed-point digits that are accumulated with 64-bit
 * INTEGER atomics (associative), which makes b - hence z, the gradient and the control update - bit-identical from
 * run to run, independent of launch geometry; it agrees with the atomic variant to round-off (1e-12 tested). */
int ocp_set_deterministic(ocp_ctx *ctx, int on);
/* Current value of an option: "deterministic", "buoy_staged" (buoy kernels with the mesh tables staged in shared
 * memory by TMA bulk copies; opt-in with environment OCP_BUOY_STAGED=1 when the tables fit an SM - measured slower
 * than the default global-table kernels on B200), "adj_reuse"; -1 unknown. */
int ocp_get_option(const ocp_ctx *ctx, const char *name);
/* Per-phase line-item timing (ocp_get_solver_stats) synchronises the stream after every phase; it is therefore off
 * by default and switched on only for profiling runs (also: environment OCP_PROFILE=1). */
void ocp_set_profiling(ocp_ctx *ctx, int on);

/* ---- forward Navier-Stokes: `solve(F == 0, w, bcs)`, OCP_dolfin.py:315-325 (406, 274, 284, 289) ----------
 * d_f: control as P2 nodal field (nn,2) (only its Gamma_1 trace is used).  d_w: in = initial guess when
 * zero_init == 0 (grad_test warm start, OCP_dolfin.py:274), out = solution.  Newton with dolfin's defaults
 * (atol 1e-10, rtol 1e-9, max 50, full step).  h_res_hist (host, >= 52 doubles or NULL) receives ||F||_2
 * per iterate. */
int ocp_forward_solve(ocp_ctx *ctx, const double *d_f, double *d_w, int zero_init, int *newton_its,
                      double *h_res_hist);

/* One assembly of the forward residual F(w; f) (BC rows = w_d) and/or the Newton matrix dF/dw on the CSR
 * pattern (BC rows -> identity when apply_bc).  `assemble(F)`, `assemble(J)` inside dolfin's NewtonSolver. */
int ocp_assemble_forward(ocp_ctx *ctx, const double *d_w, const double *d_f, double *d_vals, double *d_res,
                         int apply_bc);

/* `A = assemble(aAdj)` (+ `bcs[0].apply(A)` when apply_bc), OCP_dolfin.py:344-350, 368. */
int ocp_assemble_adjoint(ocp_ctx *ctx, const double *d_w, double *d_vals, int apply_bc);

/* `project(grad(w.sub(0)), V_vec)`, OCP_dolfin.py:328-329 -> (nv,4). */
int ocp_project_grad(ocp_ctx *ctx, const double *d_w, double *d_g);

/* `w.sub(0)` as a nodal field (nn,2) and back (`W` vector with zero pressure block). */
int ocp_velocity_nodal(ocp_ctx *ctx, const double *d_w, double *d_vel);

/* ---- buoy ODE: solve_primal_ode, OCP_dolfin.py:201-230 ----------------------------------------------------
 * d_x0 (K,2) start points; d_x, d_u (nt,K,2) time-major; d_cell (nt,K) int32 or NULL: cell used for each
 * velocity sample (-1 where the reference stored none); d_mask (K) fp64, set to 1.0 for buoys that left the
 * domain (never cleared - the caller zeroes it per iteration like `np.zeros(K)`, OCP_dolfin.py:311);
 * d_parked (K) uint8: 1 where only the last sample left the domain (OCP_dolfin.py:226-229). */
int ocp_buoy_forward(ocp_ctx *ctx, const double *d_vel, const double *d_x0, int K, double *d_x, double *d_u,
                     int32_t *d_cell, double *d_mask, uint8_t *d_parked);

/* ---- adjoint ODE + point sources + misfit in one backward sweep ---------------------------------------------
 * solve_adjoint_ode (OCP_dolfin.py:234-252), the PointSource loop (OCP_dolfin.py:353-366) and partA of J
 * (OCP_dolfin.py:259).  d_mu (nt,K,2) or NULL.  d_acc (2*nn+2) is ADDED to:
 * [b nodal | 0.5*sum h|u-u_d|^2 | number of masked buoys]. */
int ocp_buoy_adjoint_scatter(ocp_ctx *ctx, const double *d_vel, const double *d_g, int K, const double *d_x,
                             const double *d_u, const double *d_ud, const double *d_mask,
                             const uint8_t *d_parked, double *d_mu, double *d_acc);

/* partA of J only (line search / grad_test evaluate J without an adjoint), added to d_out[0]. */
int ocp_misfit(ocp_ctx *ctx, int K, const double *d_u, const double *d_ud, double *d_out);

/* ---- adjoint Navier-Stokes: OCP_dolfin.py:336-371 -----------------------------------------------------------
 * assembles aAdj at d_w, builds b from the nodal point-source vector (first 2*nn entries of d_acc), applies
 * the Dirichlet rows to A and b and solves.  d_z (ndofs). */
int ocp_adjoint_solve(ocp_ctx *ctx, const double *d_w, const double *d_bnode, double *d_z);

/* ---- boundary forms on ds(1) --------------------------------------------------------------------------------
 * int_{Gamma_1} a . b ds for P2 nodal fields: `assemble(alpha*0.5*inner(f,f)*ds(1))` (OCP_dolfin.py:260) and
 * `assemble(inner(alpha*f - zSol, df)*ds(1))` (OCP_dolfin.py:379, 388).  d_out[0] is overwritten. */
int ocp_boundary_inner(ocp_ctx *ctx, const double *d_a, const double *d_b, double *d_out);

/* d_out (n,2) = ca * a + cb * b on nodal fields: gradient alpha*f - z (OCP_dolfin.py:379) and the update
 * f <- f - LR (alpha f - z) (OCP_dolfin.py:426).  d_out may alias an input. */
int ocp_nodal_axpby(ocp_ctx *ctx, double ca, const double *d_a, double cb, const double *d_b, double *d_out);

/* sqrt(assemble(div(u)*div(u)*dx)) (OCP_dolfin.py:430), velocity L2 and H1 norms (Pipeline_limits.py:433-443);
 * d_out[0..2] = ||div u||^2, ||u||^2_L2, |u|^2_H1. */
int ocp_field_norms(ocp_ctx *ctx, const double *d_w, double *d_out);

/* (K,nt,2) <-> (nt,K,2) on the device.  to_time_major != 0: src is the reference layout. */
int ocp_traj_transpose(ocp_ctx *ctx, const double *d_src, double *d_dst, int K, int to_time_major);

/* ---- host-buffer entry points (what a reference-side binding calls; copies are part of the call) -----------
 * solve_primal_ode(wSol, buoy_mask) -> x, u_values_array in the reference's (K,nt,2) layout. */
int ocp_solve_primal_ode_host(ocp_ctx *ctx, const double *h_w, const double *h_x0, int K, double *h_x,
                              double *h_u, double *h_mask);
/* solve_adjoint_ode(wSol, grad_u, x, buoy_mask, u_values_array) -> mu (K,nt,2); h_g is (nv,4). */
int ocp_solve_adjoint_ode_host(ocp_ctx *ctx, const double *h_g, const double *h_x, const double *h_u,
                               const double *h_ud, const double *h_mask, int K, double *h_mu);
/* The reference loads u_d / x_0 once into module globals (OCP_dolfin.py:176-183); this uploads them once:
 * h_x0 (K,2), h_ud (K,nt,2) in the reference layout. */
int ocp_set_observations_host(ocp_ctx *ctx, const double *h_x0, const double *h_ud, int K);
/* One full gradient evaluation for control h_f (nn,2) on the resident observations: forward solve, projection,
 * primal ODE, adjoint sweep, adjoint solve (the "outer" block OCP_dolfin.py:313-371).  Outputs: h_w (ndofs),
 * h_z (ndofs), h_mask (K), h_scalars[0..3] = misfit, int_G1 |f|^2, n_masked, newton iterations. */
int ocp_gradient_host(ocp_ctx *ctx, const double *h_f, double *h_w, double *h_z, double *h_mask,
                      double *h_scalars);

/* ---- multi-GPU: buoys sharded over the GPUs of one box -------------------------------------------
 * The reference is serial; its sum over buoys (the PointSource loop OCP_dolfin.py:353-366 and partA of J,
 * OCP_dolfin.py:259) is what gets partitioned: every rank holds a shard of the buoys and a replica of mesh, state and
 * factorisation, and the accumulator [b | misfit | n_masked] is summed across ranks by ONE NCCL all-reduce (fp64)
 * per gradient evaluation, on the context's stream.  Set-up: rank 0 calls ocp_comm_get_unique_id, the host ships the
 * OCP_COMM_ID_BYTES bytes to the other ranks (MPI, a file, torch.distributed ...), every rank calls ocp_comm_init
 * (collective, blocks until all ranks joined).  libnccl.so.2 is resolved with dlopen at the first call
 * (environment OCP_NCCL_LIB overrides the name); without it these calls return OCP_ERR_COMM.
 * With a communicator, ocp_gradient_host all-reduces the accumulator between the backward sweep and the adjoint
 * solve, so every rank returns the gradient of the WHOLE buoy set; h_scalars[0,2] are then global sums. */
#define OCP_COMM_ID_BYTES 128
int ocp_comm_get_unique_id(void *id128);
int ocp_comm_init(ocp_ctx *ctx, int nranks, int rank, const void *id128);
int ocp_comm_size(const ocp_ctx *ctx);          /* 1 without a communicator */
int ocp_comm_nccl_version(void);                /* e.g. 22809; 0 when NCCL cannot be loaded */
/* in-place sum over the ranks of d_buf[0..n) (device, fp64) on the context's stream; no-op for a single rank */
int ocp_allreduce(ocp_ctx *ctx, double *d_buf, size_t n);

/* ---- measurement helpers (bench.py) ---------------------------------------------------------------------------------
 * out8 = [flops of one numeric factorisation of the W x W matrix (from the symbolic analysis), non-zeros of L+U,
 * tree levels, largest front, front workspace in doubles, flops and non-zeros of the P1 mass factorisation, fronts]. */
void ocp_get_solver_info(const ocp_ctx *ctx, double *out8);
/* fp64 FMA-pipe peak of the device, TFlop/s: register-resident DFMA loop on every SM, best of 5 launches (CUDA
 * events).  The denominator of the sparse-LU roofline; MEASURED_PEAKS.json holds no fp64 figure. */
int ocp_selftest_fp64_peak(ocp_ctx *ctx, double *tflops);

/* ---- one gradient evaluation on DEVICE buffers: the "outer" block OCP_dolfin.py:313-371 in one call ---------------
 * forward solve from the zero guess with control d_f (nn,2), projection, primal ODE from d_x0 (K,2), backward sweep
 * against d_ud (nt,K,2), all-reduce (when a communicator is set), adjoint solve; d_mask (K) and d_acc (2 nn + 2) are
 * zeroed inside.  Outputs: d_w, d_g (nv,4), d_vel (nn,2), d_x, d_u (nt,K,2), d_mask, d_parked, d_acc, d_z and, when both
 * d_znod and d_grad are given, grad j = alpha f - z on the nodes (OCP_dolfin.py:379).  Once a plain evaluation has
 * run, the whole call is replayed as ONE CUDA graph per (buffers, expected Newton count); the host decisions of
 * the block (Newton converged, adjoint residual gate, pivot flags) are verified after the replay and the plain path
 * re-runs the evaluation if they do not hold (OCP_STEP_GRAPH=0 disables).  The call returns with the stream idle. */
int ocp_gradient_device(ocp_ctx *ctx, const double *d_f, const double *d_x0, const double *d_ud, int K, double *d_w,
                        double *d_g, double *d_vel, double *d_x, double *d_u, double *d_mask, uint8_t *d_parked,
                        double *d_acc, double *d_z, double *d_znod, double *d_grad, double alpha, int *newton_its);

/* ||F||_2 per Newton iterate of the LAST completed forward solve (the history lives on the device and is copied to the
 * host at the solve's one synchronisation); returns the number of entries written (iterates + 1). */
int ocp_newton_history(const ocp_ctx *ctx, double *h_hist, int max_entries);

/* Number of CUDA kernels this library has launched in this process (bench.py reports it as gpu_launches). */
long long ocp_launch_count(void);

/* ---- one-time host symbolic analysis (exported for CPU-side tests of the ordering / pivoting) --------------
 * LU of the CSR matrix with nested-dissection column order and threshold partial pivoting:
 * A[p,:][:,q] = L U, L unit lower.  Returns nnz(L)+nnz(U) or a negative error. Arrays p,q are (n). */
int64_t ocp_host_lu_probe(int32_t n, const int32_t *rowptr, const int32_t *col, const double *val,
                          const double *xy, double *rhs_inout, int32_t *p, int32_t *q);

/* Same for the multifrontal solver: host symbolic analysis (nested-dissection tree, fronts, extend-add maps) and
 * a host restatement of its numeric phase, to test the analysis without a GPU.  kind[i] = 1 for pressure dofs.
 * stats8 = [#fronts, #levels, largest front, largest pivot block, flops, min |pivot|, workspace doubles, 0]. */
int64_t ocp_host_mf_probe(int32_t n, const int32_t *rowptr, const int32_t *col, const double *val,
                          const double *xy, const uint8_t *kind, double *rhs_inout, double *stats8);
/* Pivot-search window of that host restatement: 1 (default) = static pivoting exactly like the GPU kernels; a large
 * value = partial pivoting restricted to the front's fully-summed rows (used by tests to validate static pivoting). */
void ocp_host_mf_set_pivot_window(int rows);

/* Host emulation of the atomic-free (gather) assembly on the tables ocp_create builds for it: cell part of the Newton
 * matrix (vals, nnz) and residual (res, ndofs) at w, and - when vals_t is given - its transpose through the pattern's
 * transposition permutation.  stats4 = [#CTAs, max rounds, facet colours, shared-memory bytes].  For CPU-only tests. */
int64_t ocp_host_gather_probe(const ocp_problem_desc *desc, const double *w, double nu, double *vals, double *res,
                              double *vals_t, int32_t *stats4);

/* ---- element-level self-tests (host evaluation of the kernels' __host__ __device__ element arithmetic for one
 * element; used by CPU-only unit tests, never by a compute path).  coef15 = [u_x(6) u_y(6) p(3)];
 * uv6 = [u_x(va,vb,mid) u_y(va,vb,mid)], f6 likewise for the control. */
void ocp_selftest_cell_matrix(const double *geom6, const double *coef15, double nu, double *A225, double *R15);
void ocp_selftest_facet_matrix(double len, double nx, double ny, const double *uv6, const double *f6, double *A36,
                               double *R6);

#ifdef __cplusplus
}
#endif
#endif /* OCP_B200_H_ */
