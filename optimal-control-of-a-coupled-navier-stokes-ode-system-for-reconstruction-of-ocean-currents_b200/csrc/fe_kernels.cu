// Finite-element kernels: cell-parallel assembly of the Taylor-Hood Newton matrix / residual and of the
// adjoint operator (atomic fp64 scatter into the precomputed CSR pattern through per-cell slot tables),
// Gamma_1 facet blocks, Dirichlet rows, the grad(u) projection right-hand side, boundary forms and norms.
#include <algorithm>

#include "element_math.cuh"
#include "kernels.cuh"

namespace ocp {

std::atomic<long long> g_launch_count{0};

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic grid reduction of up to three values (see buoy_kernels.cu for the two-value variant)
__device__ __forceinline__ void block_finish3(double a, double b, double c, double *scratch, unsigned *counter,
                                              double *out, int nout, bool accumulate) {
    __shared__ double sm[3][32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    if (lane == 0) {
        sm[0][wid] = a;
        sm[1][wid] = b;
        sm[2][wid] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[3] = {0.0, 0.0, 0.0};
        for (int i = 0; i < nw; ++i)
            for (int j = 0; j < 3; ++j) t[j] += sm[j][i];
        for (int j = 0; j < 3; ++j) scratch[3 * blockIdx.x + j] = t[j];
        __threadfence();
        last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < 32) {
        __threadfence();
        double t[3] = {0.0, 0.0, 0.0};
        for (unsigned i = threadIdx.x; i < gridDim.x; i += 32)
            for (int j = 0; j < 3; ++j) t[j] += __ldcg(scratch + 3 * i + j);
        for (int j = 0; j < 3; ++j) t[j] = warp_sum(t[j]);
        if (threadIdx.x == 0) {
            for (int j = 0; j < nout; ++j) out[j] = accumulate ? out[j] + t[j] : t[j];
            *counter = 0u;
        }
    }
}

// 16 lanes per cell, lane r < 15 owns row r of the 15x15 element matrix.
template <bool kTranspose>
__global__ void __launch_bounds__(256, 2)
assemble_cells_kernel(int nc, const double *__restrict__ geom, const int *__restrict__ cell_dofs,
                      const int *__restrict__ slots, const double *__restrict__ w, double nu,
                      double *__restrict__ vals, double *__restrict__ res) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int cell = t >> 4, row = t & 15;
    if (cell >= nc || row == 15) return;
    double g[6], U[6], V[6], P[3];
    const double2 *gp = reinterpret_cast<const double2 *>(geom) + 3 * (size_t)cell;
    const double2 ga = __ldg(gp), gb = __ldg(gp + 1), gc = __ldg(gp + 2);
    g[0] = ga.x; g[1] = ga.y; g[2] = gb.x; g[3] = gb.y; g[4] = gc.x; g[5] = gc.y;
    const int *dofs = cell_dofs + 15 * (size_t)cell;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        U[i] = __ldg(w + __ldg(dofs + i));
        V[i] = __ldg(w + __ldg(dofs + 6 + i));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) P[i] = __ldg(w + __ldg(dofs + 12 + i));
    double A[15], R;
    cell_row(g, U, V, P, nu, row, A, R);
    if (vals) {
        const int *sl = slots + 225 * (size_t)cell;
        const int ncol = row < 12 ? 15 : 12;       // the p-p block is structurally present but zero
#pragma unroll
        for (int j = 0; j < 15; ++j)
            if (j < ncol) atomicAdd(vals + __ldg(sl + (kTranspose ? j * 15 + row : row * 15 + j)), A[j]);
    }
    if (res) atomicAdd(res + __ldg(dofs + row), R);
}

// ---- atomic-free (gather) assembly ----------------------------------------------------------------------------------
// The scatter kernel above adds 213 fp64 REDs per cell into the CSR values: on refined meshes it is bound by the
// chip's fp64 atomic rate, reads a 900-byte slot table per cell, and the matrix changes in its last bits from run to
// run.  Here the work is organised by ROW instead: a CTA owns a run of consecutive dof rows, hence a contiguous range
// of CSR entries, which it accumulates in shared memory and writes exactly once (no memset, no atomics).
//   thread  = one (row, adjacent cell) pair: it evaluates that row of the cell's 15 x 15 element matrix - every element
//             row is still computed exactly once overall;
//   rounds  = pairs of the same row add into the row's entries one after the other (round q = the q-th cell of every
//             row; within a round all rows are disjoint), so the summation order is fixed: bit-reproducible matrices;
//   tables  = per pair 24 bytes (cell, packed offsets, 15 one-byte positions inside the row): 360 B per cell instead
//             of the 900 B slot table.
struct GatherCta {
    int first_pair, npairs, first_entry, nentries, first_row, nrows, rounds, pad;
};

__global__ void __launch_bounds__(256, 2)
assemble_gather_kernel(const GatherCta *__restrict__ ctas, const int *__restrict__ pair_cell,
                       const unsigned *__restrict__ pair_meta, const uint4 *__restrict__ pair_pos,
                       const double *__restrict__ geom, const int *__restrict__ cell_dofs, const double *__restrict__ w,
                       double nu, double *__restrict__ vals, double *__restrict__ res) {
    extern __shared__ double buf[];                     // nentries values, then nrows residual entries
    const GatherCta info = ctas[blockIdx.x];
    double *rbuf = buf + info.nentries;
    const int tid = threadIdx.x;
    for (int e = tid; e < info.nentries + info.nrows; e += blockDim.x) buf[e] = 0.0;
    double A[15], R = 0.0;
    unsigned meta = 0;
    uint4 pos = make_uint4(0, 0, 0, 0);
    const bool active = tid < info.npairs;
    if (active) {
        const int p = info.first_pair + tid;
        const int cell = __ldg(pair_cell + p);
        meta = __ldg(pair_meta + p);
        pos = __ldg(pair_pos + p);
        double g[6], U[6], V[6], Pr[3];
        const double2 *gp = reinterpret_cast<const double2 *>(geom) + 3 * (size_t)cell;
        const double2 ga = __ldg(gp), gb = __ldg(gp + 1), gc = __ldg(gp + 2);
        g[0] = ga.x; g[1] = ga.y; g[2] = gb.x; g[3] = gb.y; g[4] = gc.x; g[5] = gc.y;
        const int *dofs = cell_dofs + 15 * (size_t)cell;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            U[i] = __ldg(w + __ldg(dofs + i));
            V[i] = __ldg(w + __ldg(dofs + 6 + i));
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) Pr[i] = __ldg(w + __ldg(dofs + 12 + i));
        cell_row(g, U, V, Pr, nu, (int)((meta >> 4) & 15u), A, R);
    }
    __syncthreads();
    const int rowoff = (int)(meta >> 16), rowlocal = (int)((meta >> 8) & 255u), myq = (int)(meta & 15u);
    const int ncol = ((meta >> 4) & 15u) < 12u ? 15 : 12;       // the p-p block is structurally present but zero
    const unsigned char *pb = reinterpret_cast<const unsigned char *>(&pos);
    for (int q = 0; q < info.rounds; ++q) {
        if (active && myq == q) {
            if (vals) {
#pragma unroll
                for (int j = 0; j < 15; ++j)
                    if (j < ncol) buf[rowoff + pb[j]] += A[j];
            }
            rbuf[rowlocal] += R;
        }
        __syncthreads();
    }
    if (vals)
        for (int e = tid; e < info.nentries; e += blockDim.x) vals[info.first_entry + e] = buf[e];
    if (res)
        for (int r = tid; r < info.nrows; r += blockDim.x) res[info.first_row + r] = rbuf[r];
}

// Gamma_1 facet blocks without atomics: ONE CTA walks the facets colour by colour (facets of a colour share no node,
// so their read-modify-writes of the CSR values never collide; a block barrier separates the colours).
template <bool kTranspose>
__global__ void __launch_bounds__(512)
assemble_facets_ordered_kernel(int ncolors, const int *__restrict__ color_ptr, const int *__restrict__ color_facets,
                               const int *__restrict__ g1_nodes, const int *__restrict__ g1_dofs,
                               const int *__restrict__ g1_slots, const double *__restrict__ g1_len,
                               const double *__restrict__ g1_normal, const int *__restrict__ dof_ux,
                               const int *__restrict__ dof_uy, const double *__restrict__ w, const double *__restrict__ f,
                               double *__restrict__ vals, double *__restrict__ res) {
    for (int col = 0; col < ncolors; ++col) {
        const int f0 = color_ptr[col], nf = color_ptr[col + 1] - f0;
        for (int t = threadIdx.x; t < nf * 8; t += blockDim.x) {
            const int fct = color_facets[f0 + (t >> 3)], r = t & 7;
            if (r >= 6) continue;
            double U[3], V[3], Fx[3] = {0.0, 0.0, 0.0}, Fy[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int n = __ldg(g1_nodes + 3 * fct + a);
                U[a] = __ldg(w + __ldg(dof_ux + n));
                V[a] = __ldg(w + __ldg(dof_uy + n));
                if (f) {
                    Fx[a] = __ldg(f + 2 * (size_t)n);
                    Fy[a] = __ldg(f + 2 * (size_t)n + 1);
                }
            }
            double A[6], R;
            facet_row(__ldg(g1_len + fct), __ldg(g1_normal + 2 * fct), __ldg(g1_normal + 2 * fct + 1), U, V, Fx, Fy, r, A, R);
            // all read-modify-writes of this (facet, row) are issued together: one L2 round trip, not seven
            double *dst[7];
            double cur[7];
#pragma unroll
            for (int j = 0; j < 6; ++j)
                dst[j] = vals ? vals + __ldg(g1_slots + 36 * fct + (kTranspose ? j * 6 + r : r * 6 + j)) : nullptr;
            dst[6] = res ? res + __ldg(g1_dofs + 6 * fct + r) : nullptr;
#pragma unroll
            for (int j = 0; j < 7; ++j) cur[j] = dst[j] ? __ldcg(dst[j]) : 0.0;
#pragma unroll
            for (int j = 0; j < 6; ++j)
                if (dst[j]) __stcg(dst[j], cur[j] + A[j]);
            if (dst[6]) __stcg(dst[6], cur[6] + R);
        }
        __syncthreads();
    }
}

// out[k] = in[perm[k]]: the adjoint operator is the transpose of the nu = 1 Newton matrix on the same (symmetric) pattern
__global__ void permute_values_kernel(int n, const int *__restrict__ perm, const double *__restrict__ in,
                                      double *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = in[__ldg(perm + k)];
}

// 8 lanes per facet, lane r < 6 owns row r of the 6x6 facet block.
template <bool kTranspose>
__global__ void __launch_bounds__(128)
assemble_facets_kernel(int n_g1, const int *__restrict__ g1_nodes, const int *__restrict__ g1_dofs,
                       const int *__restrict__ g1_slots, const double *__restrict__ g1_len,
                       const double *__restrict__ g1_normal, const int *__restrict__ dof_ux,
                       const int *__restrict__ dof_uy, const double *__restrict__ w, const double *__restrict__ f,
                       double *__restrict__ vals, double *__restrict__ res) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int fct = t >> 3, r = t & 7;
    if (fct >= n_g1 || r >= 6) return;
    double U[3], V[3], Fx[3] = {0.0, 0.0, 0.0}, Fy[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int n = __ldg(g1_nodes + 3 * fct + a);
        U[a] = __ldg(w + __ldg(dof_ux + n));
        V[a] = __ldg(w + __ldg(dof_uy + n));
        if (f) {
            Fx[a] = __ldg(f + 2 * (size_t)n);
            Fy[a] = __ldg(f + 2 * (size_t)n + 1);
        }
    }
    double A[6], R;
    facet_row(__ldg(g1_len + fct), __ldg(g1_normal + 2 * fct), __ldg(g1_normal + 2 * fct + 1), U, V, Fx, Fy, r, A, R);
    if (vals) {
#pragma unroll
        for (int j = 0; j < 6; ++j)
            atomicAdd(vals + __ldg(g1_slots + 36 * fct + (kTranspose ? j * 6 + r : r * 6 + j)), A[j]);
    }
    if (res) atomicAdd(res + __ldg(g1_dofs + 6 * fct + r), R);
}

__global__ void dirichlet_kernel(int n_dir, const int *__restrict__ dir, const int *__restrict__ rowptr,
                                 const int *__restrict__ col, double *__restrict__ vals, double *__restrict__ res,
                                 const double *__restrict__ w, const double *__restrict__ dirval) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_dir) return;
    const int d = dir[warp];
    if (vals)
        for (int p = rowptr[d] + lane; p < rowptr[d + 1]; p += 32) vals[p] = (col[p] == d) ? 1.0 : 0.0;
    if (res && lane == 0) res[d] = w ? (w[d] - (dirval ? dirval[warp] : 0.0)) : 0.0;
}

__global__ void __launch_bounds__(256)
sumsq_kernel(int n, const double *__restrict__ v, double *out, double *scratch, unsigned *counter) {
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += v[i] * v[i];
    block_finish3(s, 0.0, 0.0, scratch, counter, out, 1, false);
}

__global__ void axpy_kernel(int n, double a, const double *__restrict__ x, double *__restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += a * x[i];
}

__global__ void axpby_kernel(int n, double a, const double *x, double b, const double *y, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a * x[i] + b * y[i];
}

__global__ void velocity_nodal_kernel(int nn, const int *__restrict__ dof_ux, const int *__restrict__ dof_uy,
                                      const double *__restrict__ w, double2 *__restrict__ vel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nn) vel[i] = make_double2(w[dof_ux[i]], w[dof_uy[i]]);
}

__global__ void rhs_from_nodal_kernel(int nn, int nv, const int *__restrict__ dof_ux, const int *__restrict__ dof_uy,
                                      const int *__restrict__ dof_p, const double2 *__restrict__ bnode,
                                      double *__restrict__ b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nn) {
        const double2 v = bnode[i];
        b[dof_ux[i]] = v.x;
        b[dof_uy[i]] = v.y;
        if (i < nv) b[dof_p[i]] = 0.0;
    }
}

// Right-hand side of the L2 projection of grad(u) onto continuous P1 (OCP_dolfin.py:328-329):
// grad(u) is linear per cell, so int (du_i/dx_j) psi_a = area/12 (sum_b D_b + D_a) with D_b its vertex values.
__global__ void __launch_bounds__(128)
gradproj_rhs_kernel(int nc, int nv, const double *__restrict__ geom, const int *__restrict__ cell_nodes,
                    const int *__restrict__ cell_dofs, const double *__restrict__ w, double *__restrict__ rhs4) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= nc) return;
    const double *g = geom + 6 * (size_t)cell;
    const double g1x = g[2], g1y = g[3], g2x = g[4], g2y = g[5];
    const double g0x = -(g1x + g2x), g0y = -(g1y + g2y);
    const double area = 0.5 / fabs(g1x * g2y - g2x * g1y);
    const int *dofs = cell_dofs + 15 * (size_t)cell;
    double U[6], V[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        U[i] = w[dofs[i]];
        V[i] = w[dofs[6 + i]];
    }
    double D[3][4];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const double l0 = b == 0 ? 1.0 : 0.0, l1 = b == 1 ? 1.0 : 0.0, l2 = b == 2 ? 1.0 : 0.0;
        const double d0 = 4.0 * l0 - 1.0, d1 = 4.0 * l1 - 1.0, d2 = 4.0 * l2 - 1.0;
        double gx[6], gy[6];
        gx[0] = d0 * g0x; gy[0] = d0 * g0y;
        gx[1] = d1 * g1x; gy[1] = d1 * g1y;
        gx[2] = d2 * g2x; gy[2] = d2 * g2y;
        gx[3] = 4.0 * (l2 * g1x + l1 * g2x); gy[3] = 4.0 * (l2 * g1y + l1 * g2y);
        gx[4] = 4.0 * (l2 * g0x + l0 * g2x); gy[4] = 4.0 * (l2 * g0y + l0 * g2y);
        gx[5] = 4.0 * (l1 * g0x + l0 * g1x); gy[5] = 4.0 * (l1 * g0y + l0 * g1y);
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            a0 += U[k] * gx[k];
            a1 += U[k] * gy[k];
            a2 += V[k] * gx[k];
            a3 += V[k] * gy[k];
        }
        D[b][0] = a0; D[b][1] = a1; D[b][2] = a2; D[b][3] = a3;
    }
    const double s = area / 12.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double tot = D[0][j] + D[1][j] + D[2][j];
#pragma unroll
        for (int a = 0; a < 3; ++a)
            atomicAdd(rhs4 + (size_t)j * nv + cell_nodes[6 * (size_t)cell + a], s * (tot + D[a][j]));
    }
}

__global__ void transpose4_kernel(int nv, const double *__restrict__ src4, double *__restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nv) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[4 * (size_t)i + j] = src4[(size_t)j * nv + i];
    }
}

// int_{Gamma_1} a.b ds, exact P2 x P2 facet mass  L/30 [[4,-1,2],[-1,4,2],[2,2,16]]  (single block, deterministic)
__global__ void __launch_bounds__(256)
boundary_inner_kernel(int n_g1, const int *__restrict__ g1_nodes, const double *__restrict__ g1_len,
                      const double2 *__restrict__ a, const double2 *__restrict__ b, double *out) {
    double s = 0.0;
    for (int f = threadIdx.x; f < n_g1; f += blockDim.x) {
        const int n0 = g1_nodes[3 * f], n1 = g1_nodes[3 * f + 1], n2 = g1_nodes[3 * f + 2];
        const double2 a0 = a[n0], a1 = a[n1], a2 = a[n2], b0 = b[n0], b1 = b[n1], b2 = b[n2];
        const double mx = a0.x * (4.0 * b0.x - b1.x + 2.0 * b2.x) + a1.x * (-b0.x + 4.0 * b1.x + 2.0 * b2.x) +
                          a2.x * (2.0 * b0.x + 2.0 * b1.x + 16.0 * b2.x);
        const double my = a0.y * (4.0 * b0.y - b1.y + 2.0 * b2.y) + a1.y * (-b0.y + 4.0 * b1.y + 2.0 * b2.y) +
                          a2.y * (2.0 * b0.y + 2.0 * b1.y + 16.0 * b2.y);
        s += g1_len[f] * (mx + my) / 30.0;
    }
    __shared__ double sm[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += sm[i];
        out[0] = t;
    }
}

// ||div u||^2, ||u||^2_{L2}, |u|^2_{H1} (OCP_dolfin.py:430, Pipeline_limits.py:433-443)
__global__ void __launch_bounds__(128)
field_norms_kernel(int nc, const double *__restrict__ geom, const int *__restrict__ cell_dofs,
                   const double *__restrict__ w, double *out3, double *scratch, unsigned *counter) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    double sdiv = 0.0, sl2 = 0.0, sh1 = 0.0;
    if (cell < nc) {
        const double *g = geom + 6 * (size_t)cell;
        const double g1x = g[2], g1y = g[3], g2x = g[4], g2y = g[5];
        const double g0x = -(g1x + g2x), g0y = -(g1y + g2y);
        const double area = 0.5 / fabs(g1x * g2y - g2x * g1y);
        const int *dofs = cell_dofs + 15 * (size_t)cell;
        double U[6], V[6];
        for (int i = 0; i < 6; ++i) {
            U[i] = w[dofs[i]];
            V[i] = w[dofs[6 + i]];
        }
        for (int q = 0; q < 7; ++q) {
            double l0, l1, l2, wq;
            tri_qpoint(q, l0, l1, l2, wq);
            double phi[6], gx[6], gy[6];
            phi[0] = l0 * (2.0 * l0 - 1.0); phi[1] = l1 * (2.0 * l1 - 1.0); phi[2] = l2 * (2.0 * l2 - 1.0);
            phi[3] = 4.0 * l1 * l2; phi[4] = 4.0 * l0 * l2; phi[5] = 4.0 * l0 * l1;
            const double d0 = 4.0 * l0 - 1.0, d1 = 4.0 * l1 - 1.0, d2 = 4.0 * l2 - 1.0;
            gx[0] = d0 * g0x; gy[0] = d0 * g0y;
            gx[1] = d1 * g1x; gy[1] = d1 * g1y;
            gx[2] = d2 * g2x; gy[2] = d2 * g2y;
            gx[3] = 4.0 * (l2 * g1x + l1 * g2x); gy[3] = 4.0 * (l2 * g1y + l1 * g2y);
            gx[4] = 4.0 * (l2 * g0x + l0 * g2x); gy[4] = 4.0 * (l2 * g0y + l0 * g2y);
            gx[5] = 4.0 * (l1 * g0x + l0 * g1x); gy[5] = 4.0 * (l1 * g0y + l0 * g1y);
            double ux = 0, uy = 0, uxx = 0, uxy = 0, uyx = 0, uyy = 0;
            for (int b = 0; b < 6; ++b) {
                ux += U[b] * phi[b]; uy += V[b] * phi[b];
                uxx += U[b] * gx[b]; uxy += U[b] * gy[b];
                uyx += V[b] * gx[b]; uyy += V[b] * gy[b];
            }
            const double wa = wq * area;
            sdiv += wa * (uxx + uyy) * (uxx + uyy);
            sl2 += wa * (ux * ux + uy * uy);
            sh1 += wa * (uxx * uxx + uxy * uxy + uyx * uyx + uyy * uyy);
        }
    }
    block_finish3(sdiv, sl2, sh1, scratch, counter, out3, 3, false);
}

// r = b - A x, one warp per row
__global__ void __launch_bounds__(256)
spmv_residual_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ col,
                     const double *__restrict__ vals, const double *__restrict__ x, const double *__restrict__ b,
                     double *__restrict__ r) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    double s = 0.0;
    for (int p = rowptr[row] + lane; p < rowptr[row + 1]; p += 32) s += vals[p] * x[col[p]];
    s = warp_sum(s);
    if (lane == 0) r[row] = b[row] - s;
}

__global__ void __launch_bounds__(256)
fp64_peak_kernel(double *out, int iters) {
    const double y = 0.99999999, x = 1e-9 * (threadIdx.x + 1);
    double a0 = x, a1 = 2 * x, a2 = 3 * x, a3 = 4 * x, a4 = 5 * x, a5 = 6 * x, a6 = 7 * x, a7 = 8 * x;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, y, x); a1 = fma(a1, y, x); a2 = fma(a2, y, x); a3 = fma(a3, y, x);
        a4 = fma(a4, y, x); a5 = fma(a5, y, x); a6 = fma(a6, y, x); a7 = fma(a7, y, x);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// part[(sp NR + r) n + i] = sum_{j in split sp} R[i + j n] x[r ldx + (idx ? idx[j] : j)],  r < NR  (R column-major
// n x ncol): dense response operators that replace a sparse triangular solve when they are small enough to precompute
// (see capi.cu).  grid = (row blocks of 128, column splits): ~600 CTAs keep the whole chip streaming R; one thread per
// row, eight independent loads in flight; the column splits are summed in a fixed order by dense_reduce_kernel.
template <int NR>
__global__ void __launch_bounds__(128)
dense_apply_kernel(int n, int ncol, int cols_per_split, const double *__restrict__ R, const double *__restrict__ x,
                   int ldx, const int *__restrict__ idx, double *__restrict__ part) {
    extern __shared__ double xs[];          // NR x cols_per_split
    const int j0 = blockIdx.y * cols_per_split, nj = min(cols_per_split, ncol - j0);
    for (int e = threadIdx.x; e < NR * nj; e += blockDim.x) {
        const int r = e / nj, j = e - r * nj;
        xs[r * cols_per_split + j] = x[(size_t)r * ldx + (idx ? idx[j0 + j] : j0 + j)];
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) a[r] = 0.0;
    const double *Rp = R + i + (size_t)j0 * n;
    int j = 0;
    for (; j + 8 <= nj; j += 8) {
        double e[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) e[q] = __ldcs(Rp + (size_t)(j + q) * n);
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int r = 0; r < NR; ++r) a[r] = fma(e[q], xs[r * cols_per_split + j + q], a[r]);
    }
    for (; j < nj; ++j) {
        const double e0 = __ldcs(Rp + (size_t)j * n);
#pragma unroll
        for (int r = 0; r < NR; ++r) a[r] = fma(e0, xs[r * cols_per_split + j], a[r]);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) part[((size_t)blockIdx.y * NR + r) * n + i] = a[r];
}

// out[e] = sum over the splits (fixed order), e < NR n
__global__ void dense_reduce_kernel(int total, int nsplit, const double *__restrict__ part, double *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    double s = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) s += part[(size_t)sp * total + e];
    out[e] = s;
}

inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

}  // namespace

void launch_assemble_cells(int nc, const double *geom, const int *cell_dofs, const int *slots, const double *w,
                           double nu, bool transpose, double *vals, double *res, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    const int blocks = cdiv((long long)nc * 16, 256);
    if (transpose)
        assemble_cells_kernel<true><<<blocks, 256, 0, s>>>(nc, geom, cell_dofs, slots, w, nu, vals, res);
    else
        assemble_cells_kernel<false><<<blocks, 256, 0, s>>>(nc, geom, cell_dofs, slots, w, nu, vals, res);
}

void launch_assemble_facets(int n_g1, const int *g1_nodes, const int *g1_dofs, const int *g1_slots,
                            const double *g1_len, const double *g1_normal, const int *dof_ux, const int *dof_uy,
                            const double *w, const double *f, bool transpose, double *vals, double *res,
                            cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    if (n_g1 <= 0) return;
    const int blocks = cdiv((long long)n_g1 * 8, 128);
    if (transpose)
        assemble_facets_kernel<true><<<blocks, 128, 0, s>>>(n_g1, g1_nodes, g1_dofs, g1_slots, g1_len, g1_normal,
                                                            dof_ux, dof_uy, w, f, vals, res);
    else
        assemble_facets_kernel<false><<<blocks, 128, 0, s>>>(n_g1, g1_nodes, g1_dofs, g1_slots, g1_len, g1_normal,
                                                             dof_ux, dof_uy, w, f, vals, res);
}

void launch_assemble_gather(const GatherTables &gt, const double *geom, const int *cell_dofs, const double *w, double nu,
                            double *vals, double *res, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    assemble_gather_kernel<<<gt.nctas, 256, gt.smem_bytes, s>>>(reinterpret_cast<const GatherCta *>(gt.ctas), gt.pair_cell,
                                                                gt.pair_meta, reinterpret_cast<const uint4 *>(gt.pair_pos),
                                                                geom, cell_dofs, w, nu, vals, res);
}

void launch_assemble_facets_ordered(const GatherTables &gt, int n_g1, const int *g1_nodes, const int *g1_dofs,
                                    const int *g1_slots, const double *g1_len, const double *g1_normal,
                                    const int *dof_ux, const int *dof_uy, const double *w, const double *f,
                                    bool transpose, double *vals, double *res, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    if (n_g1 <= 0) return;
    if (transpose)
        assemble_facets_ordered_kernel<true><<<1, 512, 0, s>>>(gt.ncolors, gt.color_ptr, gt.color_facets, g1_nodes, g1_dofs,
                                                               g1_slots, g1_len, g1_normal, dof_ux, dof_uy, w, f, vals, res);
    else
        assemble_facets_ordered_kernel<false><<<1, 512, 0, s>>>(gt.ncolors, gt.color_ptr, gt.color_facets, g1_nodes, g1_dofs,
                                                                g1_slots, g1_len, g1_normal, dof_ux, dof_uy, w, f, vals, res);
}

void launch_permute_values(int n, const int *perm, const double *in, double *out, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    permute_values_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, perm, in, out);
}

void launch_dirichlet(int n_dir, const int *dir, const int *rowptr, const int *col, double *vals, double *res,
                      const double *w, const double *dirval, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    if (n_dir <= 0) return;
    dirichlet_kernel<<<cdiv((long long)n_dir * 32, 256), 256, 0, s>>>(n_dir, dir, rowptr, col, vals, res, w, dirval);
}

void launch_sumsq(int n, const double *v, double *out, double *scratch, unsigned *counter, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    int blocks = cdiv(n, 256);
    if (blocks > 296) blocks = 296;
    sumsq_kernel<<<blocks, 256, 0, s>>>(n, v, out, scratch, counter);
}

void launch_axpy(int n, double a, const double *x, double *y, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    axpy_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, a, x, y);
}

void launch_axpby(int n, double a, const double *x, double b, const double *y, double *out, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    axpby_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, a, x, b, y, out);
}

void launch_velocity_nodal(int nn, const int *dof_ux, const int *dof_uy, const double *w, double *vel,
                           cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    velocity_nodal_kernel<<<cdiv(nn, 256), 256, 0, s>>>(nn, dof_ux, dof_uy, w, reinterpret_cast<double2 *>(vel));
}

void launch_rhs_from_nodal(int nn, int nv, const int *dof_ux, const int *dof_uy, const int *dof_p,
                           const double *bnode, double *b, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    rhs_from_nodal_kernel<<<cdiv(nn, 256), 256, 0, s>>>(nn, nv, dof_ux, dof_uy, dof_p,
                                                        reinterpret_cast<const double2 *>(bnode), b);
}

void launch_gradproj_rhs(int nc, int nv, const double *geom, const int *cell_nodes, const int *cell_dofs,
                         const double *w, double *rhs4, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    gradproj_rhs_kernel<<<cdiv(nc, 128), 128, 0, s>>>(nc, nv, geom, cell_nodes, cell_dofs, w, rhs4);
}

void launch_transpose4(int nv, const double *src4, double *dst, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    transpose4_kernel<<<cdiv(nv, 256), 256, 0, s>>>(nv, src4, dst);
}

void launch_boundary_inner(int n_g1, const int *g1_nodes, const double *g1_len, const double *a, const double *b,
                           double *out, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    boundary_inner_kernel<<<1, 256, 0, s>>>(n_g1, g1_nodes, g1_len, reinterpret_cast<const double2 *>(a),
                                            reinterpret_cast<const double2 *>(b), out);
}

void launch_field_norms(int nc, const double *geom, const int *cell_dofs, const double *w, double *out3,
                        double *scratch, unsigned *counter, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    field_norms_kernel<<<cdiv(nc, 128), 128, 0, s>>>(nc, geom, cell_dofs, w, out3, scratch, counter);
}

size_t dense_apply_scratch(int n, int ncol, int nrhs) {
    const int rb = cdiv(n, 128);
    const int nsplit = std::max(1, std::min(600 / rb, cdiv(ncol, 16)));
    return (size_t)nsplit * nrhs * n;
}

void launch_dense_apply(int n, int ncol, int nrhs, const double *R, const double *x, int ldx, const int *idx,
                        double *out, double *part, cudaStream_t s) {
    g_launch_count.fetch_add(2, std::memory_order_relaxed);
    const int rb = cdiv(n, 128);
    const int nsplit = std::max(1, std::min(600 / rb, cdiv(ncol, 16)));
    const int cps = cdiv(ncol, nsplit);
    const int ns = cdiv(ncol, cps);
    dim3 grid(rb, ns);
    const size_t smem = sizeof(double) * cps * nrhs;
    if (nrhs == 4)
        dense_apply_kernel<4><<<grid, 128, smem, s>>>(n, ncol, cps, R, x, ldx, idx, part);
    else
        dense_apply_kernel<1><<<grid, 128, smem, s>>>(n, ncol, cps, R, x, ldx, idx, part);
    const int total = n * (nrhs == 4 ? 4 : 1);
    dense_reduce_kernel<<<cdiv(total, 256), 256, 0, s>>>(total, ns, part, out);
}

double launch_fp64_peak(double *out, int blocks, int threads, int iters, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    fp64_peak_kernel<<<blocks, threads, 0, s>>>(out, iters);
    return 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
}

void launch_spmv_residual(int n, const int *rowptr, const int *col, const double *vals, const double *x,
                          const double *b, double *r, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    spmv_residual_kernel<<<cdiv((long long)n * 32, 256), 256, 0, s>>>(n, rowptr, col, vals, x, b, r);
}

}  // namespace ocp
