// Buoy sweeps: one thread per buoy, trajectories time-major so that a warp's loads/stores are
// 512-byte contiguous runs.  The FE tables (cell geometry 48 B, cell nodes 24 B, nodal velocity
// 16 B, nodal gradient 32 B) are read through the read-only path and live in L1/L2.
//
//   buoy_forward_kernel          solve_primal_ode            OCP_dolfin.py:201-230
//   buoy_adjoint_scatter_kernel  solve_adjoint_ode           OCP_dolfin.py:234-252
//                                + PointSource loop          OCP_dolfin.py:353-366
//                                + partA of J                OCP_dolfin.py:259
#include <algorithm>

#include "element_math.cuh"
#include "kernels.cuh"

namespace ocp {

namespace {

constexpr int kBuoyThreads = 128;

__device__ __forceinline__ void load_geom(const DeviceTables &t, int c, double g[6]) {
    const double2 *p = reinterpret_cast<const double2 *>(t.geom) + 3 * (size_t)c;
    const double2 a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);
    g[0] = a.x; g[1] = a.y; g[2] = b.x; g[3] = b.y; g[4] = d.x; g[5] = d.y;
}

__device__ __forceinline__ void load_nodes(const DeviceTables &t, int c, int n[6]) {
    const int2 *p = reinterpret_cast<const int2 *>(t.cell_nodes) + 3 * (size_t)c;
    const int2 a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);
    n[0] = a.x; n[1] = a.y; n[2] = b.x; n[3] = b.y; n[4] = d.x; n[5] = d.y;
}

// Lowest-index cell whose barycentrics are all >= -tol (the oracle's definition of dolfin's "first colliding cell").
//   1. `hint` (previous cell): accepted when the point is strictly inside (margin 1e-9) - then no other cell can
//      contain it, so the answer equals the definition.
//   2. neighbour walk: leave through the edge with the most negative barycentric, at most 4 hops, again accepting
//      only strictly-inside hits.  This is what a buoy crossing into the next cell costs (one or two hops).
//   3. anything ambiguous (on an edge/vertex within the margin, outside the mesh, far jump): the definition itself,
//      ascending scan of the bin's candidate list.
// -1 = dolfin's "point outside" error.
__device__ __forceinline__ int locate(const DeviceTables &t, double x, double y, int hint, double &l0, double &l1,
                                      double &l2) {
    if (!(x == x) || !(y == y)) return -1;
    double g[6];
    if (hint >= 0) {
        int c = hint;
#pragma unroll 1
        for (int hop = 0; hop < 5; ++hop) {
            load_geom(t, c, g);
            bary(g, x, y, l0, l1, l2);
            if (l0 > kLocateMargin && l1 > kLocateMargin && l2 > kLocateMargin) return c;
            const double lm = fmin(l0, fmin(l1, l2));
            if (lm >= -kLocateTol) break;                       // within the tie zone of an edge: use the definition
            const int e = (l0 == lm) ? 0 : ((l1 == lm) ? 1 : 2);
            c = __ldg(t.cell_nbr + 3 * (size_t)c + e);
            if (c < 0) break;
        }
    }
    const double fx = floor(OCP_MUL(OCP_SUB(x, t.ox), t.ihx));
    const double fy = floor(OCP_MUL(OCP_SUB(y, t.oy), t.ihy));
    const int ix = fx < 0.0 ? 0 : (fx > (double)(t.nbx - 1) ? t.nbx - 1 : (int)fx);
    const int iy = fy < 0.0 ? 0 : (fy > (double)(t.nby - 1) ? t.nby - 1 : (int)fy);
    const int b = iy * t.nbx + ix;
    const int j1 = __ldg(t.bin_ptr + b + 1);
    for (int j = __ldg(t.bin_ptr + b); j < j1; ++j) {
        const int c = __ldg(t.bin_cells + j);
        load_geom(t, c, g);
        bary(g, x, y, l0, l1, l2);
        if (l0 >= -kLocateTol && l1 >= -kLocateTol && l2 >= -kLocateTol) return c;
    }
    return -1;
}

__device__ __forceinline__ void eval_p2(const DeviceTables &t, const double2 *__restrict__ vel, int c, double l0,
                                        double l1, double l2, double &ux, double &uy) {
    double phi[6];
    int n[6];
    p2_basis(l0, l1, l2, phi);
    load_nodes(t, c, n);
    double2 v = __ldg(vel + n[0]);
    double sx = OCP_MUL(phi[0], v.x), sy = OCP_MUL(phi[0], v.y);
#pragma unroll
    for (int i = 1; i < 6; ++i) {
        v = __ldg(vel + n[i]);
        sx = OCP_FMA(phi[i], v.x, sx);
        sy = OCP_FMA(phi[i], v.y, sy);
    }
    ux = sx;
    uy = sy;
}

// P2 velocity from the per-cell coefficient record (12 doubles: (u_x, u_y) of the cell's six nodes, contiguous), same
// operation order as eval_p2 - the record only replaces the node-index indirection by one contiguous 96-byte read.
__device__ __forceinline__ void eval_p2_cell(const double2 *__restrict__ cellvel, int c, double l0, double l1,
                                             double l2, double &ux, double &uy) {
    double phi[6];
    p2_basis(l0, l1, l2, phi);
    const double2 *r = cellvel + 6 * (size_t)c;
    double2 v = __ldg(r);
    double sx = OCP_MUL(phi[0], v.x), sy = OCP_MUL(phi[0], v.y);
#pragma unroll
    for (int i = 1; i < 6; ++i) {
        v = __ldg(r + i);
        sx = OCP_FMA(phi[i], v.x, sx);
        sy = OCP_FMA(phi[i], v.y, sy);
    }
    ux = sx;
    uy = sy;
}

// cellvel[c][i] = vel[cell_nodes[c][i]], cellg[c][a] = g[cell_nodes[c][a]] (a < 3): rebuilt whenever the state changes
__global__ void cell_records_kernel(int nc, const int *__restrict__ cell_nodes, const double2 *__restrict__ vel,
                                    double2 *__restrict__ cellvel, const double2 *__restrict__ g,
                                    double2 *__restrict__ cellg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc * 6) return;
    const int c = i / 6, a = i % 6;
    const int n = __ldg(cell_nodes + i);
    if (cellvel) cellvel[i] = __ldg(vel + n);
    if (cellg && a < 3) {
        cellg[6 * (size_t)c + 2 * a] = __ldg(g + 2 * (size_t)n);
        cellg[6 * (size_t)c + 2 * a + 1] = __ldg(g + 2 * (size_t)n + 1);
    }
}

__global__ void __launch_bounds__(kBuoyThreads)
buoy_forward_kernel(DeviceTables t, const double2 *__restrict__ vel /* per-cell records */, const double2 *__restrict__ x0, int K, int nt,
                    double h, double cx, double cy, double2 *__restrict__ x, double2 *__restrict__ u,
                    int *__restrict__ cell, double *__restrict__ mask, uint8_t *__restrict__ parked) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= K) return;
    double2 p = x0[b];
    int hint = -1, kfail = -1;
    double l0, l1, l2;
    for (int k = 0; k < nt - 1; ++k) {
        const int c = locate(t, p.x, p.y, hint, l0, l1, l2);
        if (c < 0) {
            kfail = k;
            break;
        }
        double ux, uy;
        eval_p2_cell(vel, c, l0, l1, l2, ux, uy);
        const size_t o = (size_t)k * K + b;
        x[o] = p;
        u[o] = make_double2(ux, uy);
        if (cell) cell[o] = c;
        p.x = OCP_ADD(p.x, OCP_MUL(h, ux));      // two roundings, as numpy at OCP_dolfin.py:212
        p.y = OCP_ADD(p.y, OCP_MUL(h, uy));
        hint = c;
    }
    if (kfail < 0) {
        // trailing evaluation at the last sample, OCP_dolfin.py:223-229
        const size_t o = (size_t)(nt - 1) * K + b;
        const int c = locate(t, p.x, p.y, hint, l0, l1, l2);
        if (c >= 0) {
            double ux, uy;
            eval_p2_cell(vel, c, l0, l1, l2, ux, uy);
            x[o] = p;
            u[o] = make_double2(ux, uy);
            if (cell) cell[o] = c;
            parked[b] = 0;
        } else {
            x[o] = make_double2(cx, cy);
            u[o] = make_double2(0.0, 0.0);
            if (cell) cell[o] = -1;
            parked[b] = 1;
        }
        return;
    }
    // the `except` branch, OCP_dolfin.py:213-221: park the whole trajectory at the centre, mask the buoy;
    // samples 0..kfail-1 keep their velocities, sample kfail stays 0, sample kfail+1 gets u(centre)
    mask[b] = 1.0;
    parked[b] = 0;
    const int cc = locate(t, cx, cy, -1, l0, l1, l2);
    double ucx = 0.0, ucy = 0.0;
    if (cc >= 0) eval_p2_cell(vel, cc, l0, l1, l2, ucx, ucy);
    for (int k = 0; k < nt; ++k) {
        const size_t o = (size_t)k * K + b;
        x[o] = make_double2(cx, cy);
        if (k == kfail + 1) {
            u[o] = make_double2(ucx, ucy);
            if (cell) cell[o] = cc;
        } else if (k >= kfail) {
            u[o] = make_double2(0.0, 0.0);
            if (cell) cell[o] = -1;
        }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sums of two values -> partial[2*blockIdx.x .. +1]; the last block to finish adds all partials
// in block order (deterministic) into out0/out1 and resets the counter.
__device__ __forceinline__ void block_finish2(double a, double b, double *scratch, unsigned *counter, double *out0,
                                              double *out1, double scale0) {
    __shared__ double sa[32], sb[32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
        sa[wid] = a;
        sb[wid] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int i = 0; i < nw; ++i) {
            ta += sa[i];
            tb += sb[i];
        }
        scratch[2 * blockIdx.x] = ta;
        scratch[2 * blockIdx.x + 1] = tb;
        __threadfence();
        last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < 32) {
        __threadfence();
        double ta = 0.0, tb = 0.0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) {   // fixed assignment -> deterministic
            ta += __ldcg(scratch + 2 * i);
            tb += __ldcg(scratch + 2 * i + 1);
        }
        ta = warp_sum(ta);
        tb = warp_sum(tb);
        if (threadIdx.x == 0) {
            if (out0) *out0 += scale0 * ta;
            if (out1) *out1 += tb;
            *counter = 0u;
        }
    }
}

__device__ __forceinline__ void flush_sources(double *__restrict__ bnode, const DeviceTables &t, int c,
                                              double acc[12]) {
    int n[6];
    load_nodes(t, c, n);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        atomicAdd(bnode + 2 * (size_t)n[i], acc[i]);
        atomicAdd(bnode + 2 * (size_t)n[i] + 1, acc[6 + i]);
        acc[i] = 0.0;
        acc[6 + i] = 0.0;
    }
}

// Backward sweep k = nt-1 .. 0.  Per sample: gamma_k = h((u_d - u(x_k)) + mu_k) is deposited as
// gamma_c phi_i(x_k); deposits are accumulated in registers while the buoy stays in one cell and
// flushed with 12 fp64 atomics when it changes cell (a buoy crosses a handful of cells per trajectory),
// then mu_{k-1} = mu_k - h G(x_k)^T ((u_k - u_d,k) - mu_k).
__global__ void __launch_bounds__(kBuoyThreads, 5)
buoy_adjoint_scatter_kernel(DeviceTables t, const double2 *__restrict__ vel /* per-cell records */,
                            const double2 *__restrict__ g /* per-cell vertex gradients */, int K,
                            int nt, double h, double cx, double cy, const double2 *__restrict__ x,
                            const double2 *__restrict__ u, const double2 *__restrict__ ud,
                            const double *__restrict__ mask, const uint8_t *__restrict__ parked,
                            double2 *__restrict__ mu, double *__restrict__ acc_out, double *scratch,
                            unsigned *counter, double *__restrict__ bpriv, int nrep) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    double misfit = 0.0, nmasked = 0.0;
    // point sources go to one of `nrep` private copies of b (selected by SM id) so that fp64 atomics of the
    // thousands of buoys sharing a cell do not serialise on 12 addresses; the copies are summed afterwards
    double *bdst = acc_out;
    if (nrep > 1) {
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        bdst = bpriv + (size_t)(smid % (unsigned)nrep) * (2 * (size_t)t.nn);
    }
    if (b < K) {
        const bool masked = mask[b] != 0.0;
        const bool park = parked[b] != 0;
        nmasked = masked ? 1.0 : 0.0;
        double mux = 0.0, muy = 0.0;
        double acc[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) acc[i] = 0.0;
        int acc_cell = -1, hint = -1;
        // the three streams are read one sample ahead so that their HBM latency overlaps the arithmetic
        double2 pn = __ldcs(x + (size_t)(nt - 1) * K + b), Un = __ldcs(u + (size_t)(nt - 1) * K + b),
                Dn = __ldcs(ud + (size_t)(nt - 1) * K + b);
        for (int k = nt - 1; k >= 0; --k) {
            const size_t o = (size_t)k * K + b;
            double2 p = pn;
            const double2 U = Un, D = Dn;
            if (k > 0) {
                pn = __ldcs(x + o - K);
                Un = __ldcs(u + o - K);
                Dn = __ldcs(ud + o - K);
            }
            const double ex = U.x - D.x, ey = U.y - D.y;
            misfit += ex * ex + ey * ey;
            if (masked) {
                if (mu) __stcs(mu + o, make_double2(0.0, 0.0));
                continue;
            }
            double l0, l1, l2;
            double ukx = U.x, uky = U.y;
            int c = locate(t, p.x, p.y, hint, l0, l1, l2);
            if (c < 0) {            // `except` of OCP_dolfin.py:359-361: u_x = 0, point = centre
                ukx = 0.0;
                uky = 0.0;
                p = make_double2(cx, cy);
                c = locate(t, cx, cy, -1, l0, l1, l2);
            } else if (park && k == nt - 1) {
                // the stored velocity of a parked last sample is 0, the scatter loop re-evaluates u(centre)
                eval_p2_cell(vel, c, l0, l1, l2, ukx, uky);
            }
            if (mu) __stcs(mu + o, make_double2(mux, muy));
            if (c >= 0) {
                hint = c;
                if (c != acc_cell) {
                    if (acc_cell >= 0) flush_sources(bdst, t, acc_cell, acc);
                    acc_cell = c;
                }
                const double gx = h * ((D.x - ukx) + mux), gy = h * ((D.y - uky) + muy);
                double phi[6];
                p2_basis(l0, l1, l2, phi);
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    acc[i] = fma(gx, phi[i], acc[i]);
                    acc[6 + i] = fma(gy, phi[i], acc[6 + i]);
                }
                if (k > 0) {
                    // continuous P1 tensor at the cell's three vertices, [g00 g01 | g10 g11] per vertex
                    const double2 *gr = g + 6 * (size_t)c;
                    const double2 a0 = __ldg(gr), a1 = __ldg(gr + 1), b0 = __ldg(gr + 2), b1 = __ldg(gr + 3),
                                  c0 = __ldg(gr + 4), c1 = __ldg(gr + 5);
                    const double G0 = l0 * a0.x + l1 * b0.x + l2 * c0.x;
                    const double G1 = l0 * a0.y + l1 * b0.y + l2 * c0.y;
                    const double G2 = l0 * a1.x + l1 * b1.x + l2 * c1.x;
                    const double G3 = l0 * a1.y + l1 * b1.y + l2 * c1.y;
                    const double rx = ex - mux, ry = ey - muy;
                    mux = mux - h * (G0 * rx + G2 * ry);
                    muy = muy - h * (G1 * rx + G3 * ry);
                }
            }
        }
        if (acc_cell >= 0) flush_sources(bdst, t, acc_cell, acc);
    }
    block_finish2(misfit, nmasked, scratch, counter, acc_out + 2 * (size_t)t.nn, acc_out + 2 * (size_t)t.nn + 1,
                  0.5 * h);
}

// b += sum over the private copies, in copy order (deterministic given the copies)
__global__ void reduce_private_kernel(int n, int nrep, const double *__restrict__ bpriv, double *__restrict__ b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int r = 0; r < nrep; ++r) s += bpriv[(size_t)r * n + i];
    b[i] += s;
}

__global__ void __launch_bounds__(256)
misfit_kernel(size_t n, const double2 *__restrict__ u, const double2 *__restrict__ ud, double h, double *out,
              double *scratch, unsigned *counter) {
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double2 a = __ldcs(u + i), d = __ldcs(ud + i);
        const double ex = a.x - d.x, ey = a.y - d.y;
        s += ex * ex + ey * ey;
    }
    block_finish2(s, 0.0, scratch, counter, out, nullptr, 0.5 * h);
}

// (K,nt) <-> (nt,K) transpose of double2 elements through a padded shared tile
// (grid.x always runs over the long buoy dimension; `rows_on_x` says whether that is the row index)
__global__ void transpose_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, int rows, int cols,
                                 int rows_on_x) {
    __shared__ double2 tile[32][33];
    const int bx = (rows_on_x ? blockIdx.y : blockIdx.x) * 32, by = (rows_on_x ? blockIdx.x : blockIdx.y) * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = by + j, c = bx + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = src[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = bx + j, r = by + threadIdx.x;
        if (r < rows && c < cols) dst[(size_t)c * rows + r] = tile[threadIdx.x][j];
    }
}

}  // namespace

int buoy_max_blocks(int K) { return (K + kBuoyThreads - 1) / kBuoyThreads; }

int buoy_private_copies(int K, int nc, int nn) {
    if ((long long)K < 8LL * nc) return 1;                        // few buoys per cell: no contention to avoid
    const long long cap = (256LL << 20) / (16LL * nn);            // at most 256 MiB of private copies
    return (int)std::max(1LL, std::min(148LL, cap));
}

void launch_cell_records(const DeviceTables &t, const double *vel, double *cellvel, const double *g, double *cellg,
                         cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    cell_records_kernel<<<(t.nc * 6 + 255) / 256, 256, 0, s>>>(
        t.nc, t.cell_nodes, reinterpret_cast<const double2 *>(vel), reinterpret_cast<double2 *>(cellvel),
        reinterpret_cast<const double2 *>(g), reinterpret_cast<double2 *>(cellg));
}

void launch_buoy_forward(const DeviceTables &t, const double *cellvel, const double *x0, int K, int nt, double h,
                         double cx, double cy, double *x, double *u, int *cell, double *mask, uint8_t *parked,
                         cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    if (K <= 0) return;
    buoy_forward_kernel<<<buoy_max_blocks(K), kBuoyThreads, 0, s>>>(
        t, reinterpret_cast<const double2 *>(cellvel), reinterpret_cast<const double2 *>(x0), K, nt, h, cx, cy,
        reinterpret_cast<double2 *>(x), reinterpret_cast<double2 *>(u), cell, mask, parked);
}

void launch_buoy_adjoint_scatter(const DeviceTables &t, const double *vel, const double *g, int K, int nt, double h,
                                 double cx, double cy, const double *x, const double *u, const double *ud,
                                 const double *mask, const uint8_t *parked, double *mu, double *acc,
                                 double *scratch, unsigned *counter, double *bpriv, int nrep, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    if (K <= 0) return;
    if (nrep > 1) cudaMemsetAsync(bpriv, 0, sizeof(double) * 2 * (size_t)t.nn * nrep, s);
    buoy_adjoint_scatter_kernel<<<buoy_max_blocks(K), kBuoyThreads, 0, s>>>(
        t, reinterpret_cast<const double2 *>(vel), reinterpret_cast<const double2 *>(g), K, nt, h, cx, cy,
        reinterpret_cast<const double2 *>(x), reinterpret_cast<const double2 *>(u),
        reinterpret_cast<const double2 *>(ud), mask, parked, reinterpret_cast<double2 *>(mu), acc, scratch, counter,
        bpriv, nrep);
    if (nrep > 1) {
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        reduce_private_kernel<<<(2 * t.nn + 255) / 256, 256, 0, s>>>(2 * t.nn, nrep, bpriv, acc);
    }
}

void launch_misfit(int K, int nt, double h, const double *u, const double *ud, double *out, double *scratch,
                   unsigned *counter, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    const size_t n = (size_t)K * nt;
    if (n == 0) return;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    misfit_kernel<<<blocks, 256, 0, s>>>(n, reinterpret_cast<const double2 *>(u),
                                          reinterpret_cast<const double2 *>(ud), h, out, scratch, counter);
}

void launch_traj_transpose(const double *src, double *dst, int K, int nt, int to_time_major, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    if (K <= 0) return;
    const int rows = to_time_major ? K : nt, cols = to_time_major ? nt : K;
    dim3 grid((K + 31) / 32, (nt + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(reinterpret_cast<const double2 *>(src), reinterpret_cast<double2 *>(dst),
                                            rows, cols, to_time_major ? 1 : 0);
}

}  // namespace ocp
