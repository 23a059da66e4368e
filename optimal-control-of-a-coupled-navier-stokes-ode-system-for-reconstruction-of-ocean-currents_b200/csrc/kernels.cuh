// Kernel launchers of libocp_b200 (sm_100a).  All launches go to the stream passed in.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

namespace ocp {

// process-wide count of kernels launched by this library (reported by bench.py as gpu_launches)
extern std::atomic<long long> g_launch_count;

struct DeviceTables {
    int nc, nn, nv;
    const double *geom;        // (nc,6)
    const int *cell_nodes;     // (nc,6)
    const int *cell_nbr;       // (nc,3)
    double ox, oy, ihx, ihy;
    int nbx, nby;
    const int *bin_ptr, *bin_cells;
};

// ---- buoy sweeps (buoy_kernels.cu) ---------------------------------------------------------------------------
// per-cell coefficient records: cellvel (nc,6,2) from the nodal velocity, cellg (nc,3,4) from the vertex gradients
// (either output may be null)
void launch_cell_records(const DeviceTables &t, const double *vel, double *cellvel, const double *g, double *cellg,
                         cudaStream_t s);
// `staged`: the mesh tables fit an SM's shared memory (buoy_tables_fit_shared) and are staged there by TMA bulk copies;
// `field*` are then the NODAL fields (velocity (nn,2), projected gradient (nv,4)).  Otherwise `field*` are the per-cell
// records built by launch_cell_records.
bool buoy_tables_fit_shared(int nc, int nn, int nv);
void launch_buoy_forward(const DeviceTables &t, bool staged, const double *field, const double *x0, int K, int nt,
                         double h, double cx, double cy, double *x, double *u, int *cell, double *mask,
                         uint8_t *parked, cudaStream_t s);
// scratch: >= 2*buoy_max_blocks(K)+2 doubles, counter: 1 unsigned (zero on entry, left zero).
// digits != null: reproducible deposit (integer digit sums, buoy_exact_digits(nn, nrep) long longs) instead of fp64
// atomics; bpriv (nrep > 1): private copies of b.
void launch_buoy_adjoint_scatter(const DeviceTables &t, bool staged, const double *fieldv, const double *fieldg, int K,
                                 int nt, double h, double cx, double cy, const double *x, const double *u,
                                 const double *ud, const double *mask, const uint8_t *parked, double *mu, double *acc,
                                 double *scratch, unsigned *counter, double *bpriv, int nrep, long long *digits,
                                 cudaStream_t s);
int buoy_private_copies(int K, int nc, int nn);   // number of private copies of b worth using (1 = none)
size_t buoy_exact_digits(int nn, int nrep);       // long longs of the digit accumulator (deterministic mode)
void launch_misfit(int K, int nt, double h, const double *u, const double *ud, double *out, double *scratch,
                   unsigned *counter, cudaStream_t s);
void launch_traj_transpose(const double *src, double *dst, int K, int nt, int to_time_major, cudaStream_t s);
int buoy_max_blocks(int K);

// ---- FE kernels (fe_kernels.cu) -------------------------------------------------------------------------------
void launch_assemble_cells(int nc, const double *geom, const int *cell_dofs, const int *slots, const double *w,
                           double nu, bool transpose, double *vals, double *res, cudaStream_t s);
void launch_assemble_facets(int n_g1, const int *g1_nodes, const int *g1_dofs, const int *g1_slots,
                            const double *g1_len, const double *g1_normal, const int *dof_ux, const int *dof_uy,
                            const double *w, const double *f, bool transpose, double *vals, double *res,
                            cudaStream_t s);
// ---- atomic-free assembly (fe_kernels.cu): device tables built once by ocp_create (capi.cu: build_gather_tables)
struct GatherTables {
    int nctas = 0, ncolors = 0;
    size_t smem_bytes = 0;             // largest CTA: (entries + rows) doubles
    const void *ctas = nullptr;        // GatherCta[nctas]
    const int *pair_cell = nullptr;    // per (row, cell) pair, ordered by row then cell
    const unsigned *pair_meta = nullptr;   // row offset in the CTA's entry range << 16 | local row index << 8 | element row << 4 | round
    const void *pair_pos = nullptr;    // 16 bytes per pair: position of element column j inside the CSR row
    const int *color_ptr = nullptr, *color_facets = nullptr;   // Gamma_1 facets grouped by colour
    const int *tperm = nullptr;        // CSR position of the transposed entry
};
// vals / res are WRITTEN (every entry exactly once), not accumulated: no memset beforehand
void launch_assemble_gather(const GatherTables &gt, const double *geom, const int *cell_dofs, const double *w, double nu,
                            double *vals, double *res, cudaStream_t s);
void launch_assemble_facets_ordered(const GatherTables &gt, int n_g1, const int *g1_nodes, const int *g1_dofs,
                                    const int *g1_slots, const double *g1_len, const double *g1_normal,
                                    const int *dof_ux, const int *dof_uy, const double *w, const double *f,
                                    bool transpose, double *vals, double *res, cudaStream_t s);
void launch_permute_values(int n, const int *perm, const double *in, double *out, cudaStream_t s);
// rows -> identity on the CSR values; res[d] = w[d] - dirval[d] (w null -> res[d] = 0; dirval null -> 0)
void launch_dirichlet(int n_dir, const int *dir, const int *rowptr, const int *col, double *vals, double *res,
                      const double *w, const double *dirval, cudaStream_t s);
void launch_sumsq(int n, const double *v, double *out, double *scratch, unsigned *counter, cudaStream_t s);
void launch_axpy(int n, double a, const double *x, double *y, cudaStream_t s);           // y += a x
void launch_axpby(int n, double a, const double *x, double b, const double *y, double *out, cudaStream_t s);
void launch_velocity_nodal(int nn, const int *dof_ux, const int *dof_uy, const double *w, double *vel, cudaStream_t s);
void launch_rhs_from_nodal(int nn, int nv, const int *dof_ux, const int *dof_uy, const int *dof_p,
                           const double *bnode, double *b, cudaStream_t s);
void launch_gradproj_rhs(int nc, int nv, const double *geom, const int *cell_nodes, const int *cell_dofs,
                         const double *w, double *rhs4, cudaStream_t s);
void launch_transpose4(int nv, const double *src4, double *dst, cudaStream_t s);         // (4,nv) -> (nv,4)
void launch_boundary_inner(int n_g1, const int *g1_nodes, const double *g1_len, const double *a, const double *b,
                           double *out, cudaStream_t s);
void launch_field_norms(int nc, const double *geom, const int *cell_dofs, const double *w, double *out3,
                        double *scratch, unsigned *counter, cudaStream_t s);
void launch_spmv_residual(int n, const int *rowptr, const int *col, const double *vals, const double *x,
                          const double *b, double *r, cudaStream_t s);                   // r = b - A x

// out[r n + i] = sum_j R[i + j n] x[r ldx + (idx ? idx[j] : j)], r < nrhs (1 or 4); R column-major n x ncol;
// part: dense_apply_scratch(n, ncol, nrhs) doubles of work space (partial sums of the column splits)
size_t dense_apply_scratch(int n, int ncol, int nrhs);
void launch_dense_apply(int n, int ncol, int nrhs, const double *R, const double *x, int ldx, const int *idx,
                        double *out, double *part, cudaStream_t s);
// register-resident DFMA loop (8 independent chains per thread): the fp64 FMA-pipe peak the LU roofline is quoted
// against; returns the flop count of the launch
double launch_fp64_peak(double *out, int blocks, int threads, int iters, cudaStream_t s);

}  // namespace ocp
