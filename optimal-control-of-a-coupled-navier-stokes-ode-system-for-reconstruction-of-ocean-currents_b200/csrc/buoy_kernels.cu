// Buoy sweeps: one thread per buoy, trajectories time-major so that a warp's loads/stores are
// 512-byte contiguous runs.
//
//   buoy_forward_kernel          solve_primal_ode            OCP_dolfin.py:201-230
//   buoy_adjoint_scatter_kernel  solve_adjoint_ode           OCP_dolfin.py:234-252
//                                + PointSource loop          OCP_dolfin.py:353-366
//                                + partA of J                OCP_dolfin.py:259
//
// Two table paths, same arithmetic (bit-identical trajectories, cells, masks and mu).  GLOBAL is the default: measured
// on B200 it is the faster one at every size (profiles/README.md, round 2: 2^20 buoys backward 2.05 ms vs 3.04 ms
// staged - the staged kernel is limited to one 512-thread CTA per SM by its 182-215 KB of tables, and the tables
// were L1-resident anyway; at K = 10^4 both take 0.157 ms because the launch is bound by the 200-step dependent
// chain of a single warp per scheduler, not by table latency).  STAGED stays selectable (OCP_BUOY_STAGED=1).
//   * STAGED (kStaged = true): the mesh tables a sweep reads per sample - cell geometry (48 B / cell), cell -> node
//     map (24 B / cell), nodal velocity (16 B / node) or nodal projected gradient (32 B / vertex) - are copied ONCE per
//     CTA into shared memory with TMA bulk copies (cp.async.bulk + mbarrier transaction count) by a persistent
//     grid of one CTA per SM; every per-sample table access is then a shared-memory load (~25 cycles) instead of an
//     L1/L2 round trip on the dependent chain locate -> evaluate -> advance.  Possible whenever the tables fit the
//     227 KB of an SM (the reference's 32 x 32 mesh: 215 KB forward, 182 KB backward).  Its point location falls
//     back to a WARP-COOPERATIVE bin search (locate_bins_warp).
//   * GLOBAL (kStaged = false): per-cell coefficient records (cellvel / cellg, rebuilt whenever the state changes) read
//     through the read-only path; refined meshes (cfg5) whose tables exceed shared memory.
#include <algorithm>
#include <cstdlib>

#include "element_math.cuh"
#include "kernels.cuh"

namespace ocp {

namespace {

constexpr int kBuoyThreads = 128;
constexpr int kStagedMaxThreads = 512;      // 16 warps x 128 registers fill one SM's register file (no spills)
constexpr int kDepthLarge = 2;              // samples of the input streams in flight per thread in large launches
constexpr int kDepthSmall = 4;              // ... in small launches (one warp per scheduler)
constexpr size_t kSmemBudget = 227 * 1024 - 1024;   // dynamic shared memory we allow ourselves per CTA

// ---- TMA bulk copy (global -> shared) completing on an mbarrier ------------------------------------------------
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// bytes: multiple of 16, src / dst 16-byte aligned.  Issued in pieces of 32 KiB so that several are in flight.
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    for (unsigned off = 0; off < bytes; off += 32768u) {
        const unsigned n = min(32768u, bytes - off);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(static_cast<char *>(dst) + off)),
                     "l"(static_cast<const char *>(src) + off), "r"(n), "r"(smem_addr(bar))
                     : "memory");
    }
}

// ---- table access: shared memory (staged) or the read-only global path ------------------------------------------
template <bool S>
__device__ __forceinline__ double2 ld2(const double2 *p) {
    if (S) return *p;
    return __ldg(p);
}
template <bool S>
__device__ __forceinline__ int2 ldi2(const int2 *p) {
    if (S) return *p;
    return __ldg(p);
}

template <bool S>
__device__ __forceinline__ void load_geom(const double *geom, int c, double g[6]) {
    const double2 *p = reinterpret_cast<const double2 *>(geom) + 3 * (size_t)c;
    const double2 a = ld2<S>(p), b = ld2<S>(p + 1), d = ld2<S>(p + 2);
    g[0] = a.x; g[1] = a.y; g[2] = b.x; g[3] = b.y; g[4] = d.x; g[5] = d.y;
}

template <bool S>
__device__ __forceinline__ void load_nodes(const int *cell_nodes, int c, int n[6]) {
    const int2 *p = reinterpret_cast<const int2 *>(cell_nodes) + 3 * (size_t)c;
    const int2 a = ldi2<S>(p), b = ldi2<S>(p + 1), d = ldi2<S>(p + 2);
    n[0] = a.x; n[1] = a.y; n[2] = b.x; n[3] = b.y; n[4] = d.x; n[5] = d.y;
}

// Definition of the point location (the oracle's restatement of dolfin's "first colliding cell"): the lowest-index
// cell of the point's bin whose barycentrics are all >= -tol; -1 = dolfin's "point outside" error.
// WARP-COOPERATIVE: the lanes of a warp that need this slow path are served one after the other, and for each of them
// all 32 lanes test 32 candidate cells of its bin at once (ballot + find-first-set keeps "lowest index"); the lane that
// asked then recomputes its barycentrics for the winning cell, so the result is bit-identical to a serial scan.
// `need` must be warp-uniform-convergent: every lane of the (converged) warp calls this, lanes with need == false
// only help.
template <bool S>
__device__ __forceinline__ int locate_bins_warp(const DeviceTables &t, const double *geom, bool need, double x, double y,
                                                double &l0, double &l1, double &l2) {
    const unsigned active = 0xffffffffu;        // call sites keep the warp converged (uniform trip counts)
    const int lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(active, need);
    int result = -1;
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const double sx = __shfl_sync(active, x, src), sy = __shfl_sync(active, y, src);
        const double fx = floor(OCP_MUL(OCP_SUB(sx, t.ox), t.ihx));
        const double fy = floor(OCP_MUL(OCP_SUB(sy, t.oy), t.ihy));
        const int ix = fx < 0.0 ? 0 : (fx > (double)(t.nbx - 1) ? t.nbx - 1 : (int)fx);
        const int iy = fy < 0.0 ? 0 : (fy > (double)(t.nby - 1) ? t.nby - 1 : (int)fy);
        const int bin = iy * t.nbx + ix;
        const int j0 = __ldg(t.bin_ptr + bin), j1 = __ldg(t.bin_ptr + bin + 1);
        int found = -1;
        const int rank = lane, width = 32;
        for (int jb = j0; jb < j1 && found < 0; jb += width) {
            const int j = jb + rank;
            bool hit = false;
            int c = -1;
            if (j < j1) {
                c = __ldg(t.bin_cells + j);
                double g[6], a0, a1, a2;
                load_geom<S>(geom, c, g);
                bary(g, sx, sy, a0, a1, a2);
                hit = a0 >= -kLocateTol && a1 >= -kLocateTol && a2 >= -kLocateTol;
            }
            const unsigned hits = __ballot_sync(active, hit);
            if (hits) found = __shfl_sync(active, c, __ffs(hits) - 1);     // candidates ascend with the lane rank
        }
        if (lane == src) result = found;
    }
    if (need && result >= 0) {
        double g[6];
        load_geom<S>(geom, result, g);
        bary(g, x, y, l0, l1, l2);
    }
    return result;
}

// Lowest-index cell whose barycentrics are all >= -tol.
//   1. `hint` (previous cell): accepted when the point is strictly inside (margin 1e-9) - then no other cell can
//      contain it, so the answer equals the definition.
//   2. neighbour walk: leave through the edge with the most negative barycentric, at most 4 hops, again accepting
//      only strictly-inside hits.  This is what a buoy crossing into the next cell costs (one or two hops).
//   3. anything ambiguous (on an edge/vertex within the margin, outside the mesh, far jump): the definition itself,
//      searched warp-cooperatively (locate_bins_warp).
// `alive`: lanes whose buoy has already finished pass false and only help their warp in step 3.
template <bool S>
__device__ __forceinline__ int locate(const DeviceTables &t, const double *geom, bool alive, double x, double y, int hint,
                                      double &l0, double &l1, double &l2) {
    bool need = alive;
    int c = -1;
    if (alive && (!(x == x) || !(y == y))) need = false;      // NaN position: "outside"
    if (need && hint >= 0) {
        double g[6];
        c = hint;
#pragma unroll 1
        for (int hop = 0; hop < 5; ++hop) {
            load_geom<S>(geom, c, g);
            bary(g, x, y, l0, l1, l2);
            if (l0 > kLocateMargin && l1 > kLocateMargin && l2 > kLocateMargin) {
                need = false;
                break;
            }
            const double lm = fmin(l0, fmin(l1, l2));
            if (lm >= -kLocateTol) break;                       // within the tie zone of an edge: use the definition
            const int e = (l0 == lm) ? 0 : ((l1 == lm) ? 1 : 2);
            c = __ldg(t.cell_nbr + 3 * (size_t)c + e);
            if (c < 0) break;
        }
        if (need) c = -1;
    }
    if (__any_sync(0xffffffffu, need)) {
        const int r = locate_bins_warp<S>(t, geom, need, x, y, l0, l1, l2);
        if (need) c = r;
    }
    return c;
}

// P2 velocity at barycentrics (l0,l1,l2) of cell c.  Staged: nodal field + cell -> node map in shared memory; global:
// the per-cell coefficient record (12 doubles, contiguous).  Same operation order either way.
template <bool S>
__device__ __forceinline__ void eval_p2(const double2 *__restrict__ vel, const int *cell_nodes, int c, double l0,
                                        double l1, double l2, double &ux, double &uy) {
    double phi[6];
    p2_basis(l0, l1, l2, phi);
    double2 v[6];
    if (S) {
        int n[6];
        load_nodes<true>(cell_nodes, c, n);
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = vel[n[i]];
    } else {
        const double2 *r = vel + 6 * (size_t)c;
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = __ldg(r + i);
    }
    double sx = OCP_MUL(phi[0], v[0].x), sy = OCP_MUL(phi[0], v[0].y);
#pragma unroll
    for (int i = 1; i < 6; ++i) {
        sx = OCP_FMA(phi[i], v[i].x, sx);
        sy = OCP_FMA(phi[i], v[i].y, sy);
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

__device__ __forceinline__ void load_geom_g(const DeviceTables &t, int c, double g[6]) {
    const double2 *p = reinterpret_cast<const double2 *>(t.geom) + 3 * (size_t)c;
    const double2 a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);
    g[0] = a.x; g[1] = a.y; g[2] = b.x; g[3] = b.y; g[4] = d.x; g[5] = d.y;
}

__device__ __forceinline__ void load_nodes_g(const DeviceTables &t, int c, int n[6]) {
    const int2 *p = reinterpret_cast<const int2 *>(t.cell_nodes) + 3 * (size_t)c;
    const int2 a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);
    n[0] = a.x; n[1] = a.y; n[2] = b.x; n[3] = b.y; n[4] = d.x; n[5] = d.y;
}

// ---- GLOBAL-table path (per-thread point location, per-cell coefficient records) ---------------------------------
// Lowest-index cell whose barycentrics are all >= -tol (the oracle's definition of dolfin's "first colliding cell").
//   1. `hint` (previous cell): accepted when the point is strictly inside (margin 1e-9) - then no other cell can
//      contain it, so the answer equals the definition.
//   2. neighbour walk: leave through the edge with the most negative barycentric, at most 4 hops, again accepting
//      only strictly-inside hits.  This is what a buoy crossing into the next cell costs (one or two hops).
//   3. anything ambiguous (on an edge/vertex within the margin, outside the mesh, far jump): the definition itself,
//      ascending scan of the bin's candidate list.
// -1 = dolfin's "point outside" error.
__device__ __forceinline__ int locate_g(const DeviceTables &t, double x, double y, int hint, double &l0, double &l1,
                                      double &l2) {
    if (!(x == x) || !(y == y)) return -1;
    double g[6];
    if (hint >= 0) {
        int c = hint;
#pragma unroll 1
        for (int hop = 0; hop < 5; ++hop) {
            load_geom_g(t, c, g);
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
        load_geom_g(t, c, g);
        bary(g, x, y, l0, l1, l2);
        if (l0 >= -kLocateTol && l1 >= -kLocateTol && l2 >= -kLocateTol) return c;
    }
    return -1;
}

__device__ __forceinline__ void eval_p2_nodal_g(const DeviceTables &t, const double2 *__restrict__ vel, int c, double l0,
                                        double l1, double l2, double &ux, double &uy) {
    double phi[6];
    int n[6];
    p2_basis(l0, l1, l2, phi);
    load_nodes_g(t, c, n);
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
__device__ __forceinline__ void eval_p2_cell_g(const double2 *__restrict__ cellvel, int c, double l0, double l1,
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

// Small launches (K ~ 10^4: one or two warps per SM walking 199 dependent Euler steps): a buoy stays ~60 steps in a
// cell, so the cell's geometry (6 doubles) and velocity record (12 doubles) are kept in REGISTERS while it does - a
// step inside the cell is then pure arithmetic (barycentrics, basis, 12 FMAs, Euler) instead of two dependent table
// reads.  The fast test is exactly the first hop of locate_g (same bary(), same strict-inside margin) and the
// velocity sum runs in the order of eval_p2_cell_g, so x, u and the cells stay bit-identical to the table kernel.
__global__ void __launch_bounds__(kBuoyThreads)
buoy_forward_cached_kernel(DeviceTables t, const double2 *__restrict__ vel /* per-cell records */, const double2 *__restrict__ x0, int K, int nt,
                    double h, double cx, double cy, double2 *__restrict__ x, double2 *__restrict__ u,
                    int *__restrict__ cell, double *__restrict__ mask, uint8_t *__restrict__ parked) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= K) return;
    double2 p = x0[b];
    int cur = -1, kfail = -1;
    double cg[6];
    double2 cv[6];
    double l0, l1, l2;
    auto locate = [&]() -> int {
        if (cur >= 0) {
            bary(cg, p.x, p.y, l0, l1, l2);
            if (l0 > kLocateMargin && l1 > kLocateMargin && l2 > kLocateMargin) return cur;
        }
        const int c = locate_g(t, p.x, p.y, cur, l0, l1, l2);
        if (c >= 0 && c != cur) {
            load_geom_g(t, c, cg);
            const double2 *r = vel + 6 * (size_t)c;
#pragma unroll
            for (int i = 0; i < 6; ++i) cv[i] = __ldg(r + i);
            cur = c;
        }
        return c;
    };
    auto eval = [&](double &ux, double &uy) {
        double phi[6];
        p2_basis(l0, l1, l2, phi);
        double sx = OCP_MUL(phi[0], cv[0].x), sy = OCP_MUL(phi[0], cv[0].y);
#pragma unroll
        for (int i = 1; i < 6; ++i) {
            sx = OCP_FMA(phi[i], cv[i].x, sx);
            sy = OCP_FMA(phi[i], cv[i].y, sy);
        }
        ux = sx;
        uy = sy;
    };
    for (int k = 0; k < nt - 1; ++k) {
        const int c = locate();
        if (c < 0) {
            kfail = k;
            break;
        }
        double ux, uy;
        eval(ux, uy);
        const size_t o = (size_t)k * K + b;
        x[o] = p;
        u[o] = make_double2(ux, uy);
        if (cell) cell[o] = c;
        p.x = OCP_ADD(p.x, OCP_MUL(h, ux));      // two roundings, as numpy at OCP_dolfin.py:212
        p.y = OCP_ADD(p.y, OCP_MUL(h, uy));
    }
    if (kfail < 0) {
        // trailing evaluation at the last sample, OCP_dolfin.py:223-229
        const size_t o = (size_t)(nt - 1) * K + b;
        const int c = locate();
        if (c >= 0) {
            double ux, uy;
            eval(ux, uy);
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
    // the `except` branch, OCP_dolfin.py:213-221 (see buoy_forward_global_kernel)
    mask[b] = 1.0;
    parked[b] = 0;
    const int cc = locate_g(t, cx, cy, -1, l0, l1, l2);
    double ucx = 0.0, ucy = 0.0;
    if (cc >= 0) eval_p2_cell_g(vel, cc, l0, l1, l2, ucx, ucy);
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

__global__ void __launch_bounds__(kBuoyThreads)
buoy_forward_global_kernel(DeviceTables t, const double2 *__restrict__ vel /* per-cell records */, const double2 *__restrict__ x0, int K, int nt,
                    double h, double cx, double cy, double2 *__restrict__ x, double2 *__restrict__ u,
                    int *__restrict__ cell, double *__restrict__ mask, uint8_t *__restrict__ parked) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= K) return;
    double2 p = x0[b];
    int hint = -1, kfail = -1;
    double l0, l1, l2;
    for (int k = 0; k < nt - 1; ++k) {
        const int c = locate_g(t, p.x, p.y, hint, l0, l1, l2);
        if (c < 0) {
            kfail = k;
            break;
        }
        double ux, uy;
        eval_p2_cell_g(vel, c, l0, l1, l2, ux, uy);
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
        const int c = locate_g(t, p.x, p.y, hint, l0, l1, l2);
        if (c >= 0) {
            double ux, uy;
            eval_p2_cell_g(vel, c, l0, l1, l2, ux, uy);
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
    const int cc = locate_g(t, cx, cy, -1, l0, l1, l2);
    double ucx = 0.0, ucy = 0.0;
    if (cc >= 0) eval_p2_cell_g(vel, cc, l0, l1, l2, ucx, ucy);
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

// Shared-memory image of the mesh tables of a staged sweep.  Every region starts 16-byte aligned; the bulk of each
// table arrives through TMA bulk copies, a tail shorter than 16 bytes through plain loads.
struct StagedTables {
    const double *geom;
    const int *nodes;
    const double2 *field;      // nodal velocity (forward) or nodal projected gradient, 2 x double2 per vertex (backward)
};

__host__ __device__ inline size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }

__host__ __device__ inline size_t staged_bytes(int nc, size_t field_bytes) {
    return align16(48 * (size_t)nc) + align16(field_bytes) + align16(24 * (size_t)nc);
}

__device__ __forceinline__ void stage_one(void *dst, const void *src, size_t bytes, unsigned long long *bar) {
    const unsigned bulk = (unsigned)(bytes & ~(size_t)15);
    if (bulk) bulk_copy_g2s(dst, src, bulk, bar);
    for (size_t i = bulk; i < bytes; i += 4)      // tail (tables are arrays of 4- or 8-byte items)
        *reinterpret_cast<int *>(static_cast<char *>(dst) + i) = *reinterpret_cast<const int *>(static_cast<const char *>(src) + i);
}

// One elected thread arms the mbarrier with the byte count and issues the bulk copies; everybody waits on the barrier.
__device__ __forceinline__ StagedTables stage_tables(unsigned char *smem, unsigned long long *bar, const DeviceTables &t,
                                                    const void *field, size_t field_bytes) {
    const size_t gb = 48 * (size_t)t.nc, nb = 24 * (size_t)t.nc;
    unsigned char *s_geom = smem, *s_field = s_geom + align16(gb), *s_nodes = s_field + align16(field_bytes);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, (unsigned)((gb & ~(size_t)15) + (field_bytes & ~(size_t)15) + (nb & ~(size_t)15)));
        stage_one(s_geom, t.geom, gb, bar);
        stage_one(s_field, field, field_bytes, bar);
        stage_one(s_nodes, t.cell_nodes, nb, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    StagedTables st;
    st.geom = reinterpret_cast<const double *>(s_geom);
    st.field = reinterpret_cast<const double2 *>(s_field);
    st.nodes = reinterpret_cast<const int *>(s_nodes);
    return st;
}

// solve_primal_ode.  S = staged tables (`field` = NODAL velocity (nn), persistent grid, `per_block` buoys per CTA pass)
// or global tables (`field` = per-cell records).  All lanes of a warp run the same trip counts (finished or padding
// lanes idle) so that the warp-cooperative slow path of the point location always sees a converged warp.
template <bool S>
__global__ void __launch_bounds__(S ? kStagedMaxThreads : kBuoyThreads, 1)
buoy_forward_kernel(DeviceTables t, const double2 *__restrict__ field, const double2 *__restrict__ x0, int K,
                    int per_block, int nt, double h, double cx, double cy, double2 *__restrict__ x,
                    double2 *__restrict__ u, int *__restrict__ cell, double *__restrict__ mask,
                    uint8_t *__restrict__ parked) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned long long bar;
    const double *geom = t.geom;
    const int *nodes = t.cell_nodes;
    const double2 *vel = field;
    if (S) {
        const StagedTables st = stage_tables(smem_raw, &bar, t, field, 16 * (size_t)t.nn);
        geom = st.geom;
        nodes = st.nodes;
        vel = st.field;
    }
    for (long long base = (long long)blockIdx.x * per_block; base < K; base += (long long)gridDim.x * per_block) {
        const long long bl = base + threadIdx.x;
        const bool alive = (int)threadIdx.x < per_block && bl < K;
        const int b = alive ? (int)bl : 0;
        double2 p = alive ? x0[b] : make_double2(cx, cy);
        int hint = -1, kfail = -1;
        double l0, l1, l2;
        for (int k = 0; k < nt - 1; ++k) {
            const bool run = alive && kfail < 0;
            const int c = locate<S>(t, geom, run, p.x, p.y, hint, l0, l1, l2);
            if (run) {
                if (c < 0) {
                    kfail = k;
                } else {
                    double ux, uy;
                    eval_p2<S>(vel, nodes, c, l0, l1, l2, ux, uy);
                    const size_t o = (size_t)k * K + b;
                    x[o] = p;
                    u[o] = make_double2(ux, uy);
                    if (cell) cell[o] = c;
                    p.x = OCP_ADD(p.x, OCP_MUL(h, ux));      // two roundings, as numpy at OCP_dolfin.py:212
                    p.y = OCP_ADD(p.y, OCP_MUL(h, uy));
                    hint = c;
                }
            }
        }
        {
            // trailing evaluation at the last sample, OCP_dolfin.py:223-229
            const bool run = alive && kfail < 0;
            const int c = locate<S>(t, geom, run, p.x, p.y, hint, l0, l1, l2);
            if (run) {
                const size_t o = (size_t)(nt - 1) * K + b;
                if (c >= 0) {
                    double ux, uy;
                    eval_p2<S>(vel, nodes, c, l0, l1, l2, ux, uy);
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
            }
        }
        // the `except` branch, OCP_dolfin.py:213-221: park the whole trajectory at the centre, mask the buoy;
        // samples 0..kfail-1 keep their velocities, sample kfail stays 0, sample kfail+1 gets u(centre)
        const bool failed = alive && kfail >= 0;
        const int cc = locate<S>(t, geom, failed, cx, cy, -1, l0, l1, l2);
        if (failed) {
            mask[b] = 1.0;
            parked[b] = 0;
            double ucx = 0.0, ucy = 0.0;
            if (cc >= 0) eval_p2<S>(vel, nodes, cc, l0, l1, l2, ucx, ucy);
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

template <bool S>
__device__ __forceinline__ void flush_sources(double *__restrict__ bnode, const int *cell_nodes, int c, double acc[12]) {
    int n[6];
    load_nodes<S>(cell_nodes, c, n);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        atomicAdd(bnode + 2 * (size_t)n[i], acc[i]);
        atomicAdd(bnode + 2 * (size_t)n[i] + 1, acc[6 + i]);
        acc[i] = 0.0;
        acc[6 + i] = 0.0;
    }
}

// ---- reproducible deposit (deterministic mode) ---------------------------------------------------------------------
// A double is split EXACTLY into four signed digits of weights 2^-4, 2^-36, 2^-68, 2^-100 (values below 2^-100 ~ 8e-31
// are dropped, magnitudes must stay below 2^27); the digits are accumulated with 64-bit INTEGER atomics, which are
// associative and commutative - the sum does not depend on the order in which threads, CTAs, SMs or GPUs arrive, so
// two runs (and sharded vs unsharded runs) give bit-identical point-source vectors.  finalize_exact_kernel turns the
// digit sums back into doubles with one rounding.
constexpr double kW3 = 16.0, kW2 = 68719476736.0 /* 2^36 */, kW1 = 295147905179352825856.0 /* 2^68 */,
                 kW0 = 1267650600228229401496703205376.0 /* 2^100 */;

__device__ __forceinline__ void exact_add(long long *slot /* 4 digits */, double v, long long *overflow) {
    if (!(fabs(v) < 134217728.0)) {                  // NaN, Inf or beyond the fixed-point range (2^27)
        atomicAdd(reinterpret_cast<unsigned long long *>(overflow), 1ull);
        return;
    }
    const double d3 = rint(v * kW3);
    double r = v - d3 * (1.0 / kW3);                 // exact: d3 / 16 is v rounded to a multiple of 2^-4
    const double d2 = rint(r * kW2);
    r -= d2 * (1.0 / kW2);
    const double d1 = rint(r * kW1);
    r -= d1 * (1.0 / kW1);
    const double d0 = rint(r * kW0);
    unsigned long long *u = reinterpret_cast<unsigned long long *>(slot);
    if (d0 != 0.0) atomicAdd(u + 0, (unsigned long long)(long long)d0);
    if (d1 != 0.0) atomicAdd(u + 1, (unsigned long long)(long long)d1);
    if (d2 != 0.0) atomicAdd(u + 2, (unsigned long long)(long long)d2);
    if (d3 != 0.0) atomicAdd(u + 3, (unsigned long long)(long long)d3);
}

template <bool S>
__device__ __forceinline__ void flush_sources_exact(long long *__restrict__ digits, long long *overflow,
                                                    const int *cell_nodes, int c, double acc[12]) {
    int n[6];
    load_nodes<S>(cell_nodes, c, n);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        exact_add(digits + 4 * (2 * (size_t)n[i]), acc[i], overflow);
        exact_add(digits + 4 * (2 * (size_t)n[i] + 1), acc[6 + i], overflow);
        acc[i] = 0.0;
        acc[6 + i] = 0.0;
    }
}

// out[i] += value of the digit sums of entry i.  The carries are propagated in integers first, so that the four
// digits are non-overlapping 32-bit fields and the conversion rounds exactly once.
__global__ void finalize_exact_kernel(int n, const long long *__restrict__ digits, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (digits[4 * (size_t)n] != 0) {       // a deposit was NaN / Inf / >= 2^27: do not hide it
        out[i] = nan("");
        return;
    }
    long long d0 = digits[4 * (size_t)i], d1 = digits[4 * (size_t)i + 1], d2 = digits[4 * (size_t)i + 2],
              d3 = digits[4 * (size_t)i + 3];
    // floor-division carries (arithmetic shift), digit in [0, 2^32)
    long long c = d0 >> 32; d0 -= c << 32; d1 += c;
    c = d1 >> 32; d1 -= c << 32; d2 += c;
    c = d2 >> 32; d2 -= c << 32; d3 += c;
    // value = d3 2^-4 + d2 2^-36 + d1 2^-68 + d0 2^-100 with non-overlapping digits: `hi` is exact for sums below
    // 2^17, `lo` and the final addition round - a fixed function of the integer sums, error <= 1 ulp
    const double hi = (double)d3 * (1.0 / kW3) + (double)d2 * (1.0 / kW2);
    const double lo = (double)d1 * (1.0 / kW1) + (double)d0 * (1.0 / kW0);
    out[i] += hi + lo;
}

// Backward sweep k = nt-1 .. 0.  Per sample: gamma_k = h((u_d - u(x_k)) + mu_k) is deposited as
// gamma_c phi_i(x_k); deposits are accumulated in registers while the buoy stays in one cell and
// flushed with 12 fp64 atomics when it changes cell (a buoy crosses a handful of cells per trajectory),
// then mu_{k-1} = mu_k - h G(x_k)^T ((u_k - u_d,k) - mu_k).
//   S: staged tables (fieldv = NODAL velocity, fieldg = NODAL projected gradient (nv,4); persistent grid) or global
//      tables (per-cell records cellvel / cellg).
//   D: samples of the three input streams requested ahead (register ring): 2 for bandwidth-bound launches (16 warps
//      per SM hide the rest), 4 for small launches where one warp per scheduler has to cover the HBM latency itself.
//   X: reproducible integer deposit (deterministic mode) instead of fp64 atomics.
template <bool S, int D, bool X>
__global__ void __launch_bounds__(S ? (D == kDepthSmall ? 128 : kStagedMaxThreads) : kBuoyThreads, S ? 1 : 4)
buoy_adjoint_scatter_kernel(DeviceTables t, const double2 *__restrict__ fieldv, const double2 *__restrict__ fieldg,
                            int K, int per_block, int nt, double h, double cx, double cy,
                            const double2 *__restrict__ x, const double2 *__restrict__ u,
                            const double2 *__restrict__ ud, const double *__restrict__ mask,
                            const uint8_t *__restrict__ parked, double2 *__restrict__ mu, double *__restrict__ acc_out,
                            double *scratch, unsigned *counter, double *__restrict__ bpriv, int nrep,
                            long long *__restrict__ digits) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned long long bar;
    const double *geom = t.geom;
    const int *nodes = t.cell_nodes;
    const double2 *gtab = fieldg;
    if (S) {
        const StagedTables st = stage_tables(smem_raw, &bar, t, fieldg, 32 * (size_t)t.nv);
        geom = st.geom;
        nodes = st.nodes;
        gtab = st.field;
    }
    double misfit = 0.0, nmasked = 0.0;
    // point sources go to one of `nrep` private copies of b (selected by SM id) so that fp64 atomics of the
    // thousands of buoys sharing a cell do not serialise on 12 addresses; the copies are summed afterwards
    double *bdst = acc_out;
    long long *ddst = digits;
    if (nrep > 1) {
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        const size_t copy = smid % (unsigned)nrep;
        bdst = bpriv + copy * (2 * (size_t)t.nn);
        if (X) ddst = digits + copy * (8 * (size_t)t.nn);
    }
    for (long long base = (long long)blockIdx.x * per_block; base < K; base += (long long)gridDim.x * per_block) {
        const long long bl = base + threadIdx.x;
        const bool alive = (int)threadIdx.x < per_block && bl < K;
        const int b = alive ? (int)bl : 0;
        const bool masked = alive && mask[b] != 0.0;
        const bool park = alive && parked[b] != 0;
        if (masked) nmasked += 1.0;
        const bool run = alive && !masked;
        double mux = 0.0, muy = 0.0;
        double acc[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) acc[i] = 0.0;
        int acc_cell = -1, hint = -1;
        auto flush = [&](int c) {
            if (X)
                flush_sources_exact<S>(ddst, digits + 8 * (size_t)t.nn * (nrep > 1 ? nrep : 1), nodes, c, acc);
            else
                flush_sources<S>(bdst, nodes, c, acc);
        };
        // one sample; every lane of the warp calls it (padding / masked lanes only help the cooperative search)
        auto sample = [&](int k, double2 p, const double2 U, const double2 Dd) {
            const size_t o = (size_t)k * K + b;
            const double ex = U.x - Dd.x, ey = U.y - Dd.y;
            if (alive) misfit += ex * ex + ey * ey;
            if (masked && mu) __stcs(mu + o, make_double2(0.0, 0.0));
            double l0, l1, l2;
            double ukx = U.x, uky = U.y;
            int c = locate<S>(t, geom, run, p.x, p.y, hint, l0, l1, l2);
            // `except` of OCP_dolfin.py:359-361: u_x = 0, point = centre
            const bool lost = run && c < 0;
            {
                double m0, m1, m2;
                const int cc = locate<S>(t, geom, lost, cx, cy, -1, m0, m1, m2);
                if (lost) {
                    ukx = 0.0;
                    uky = 0.0;
                    p = make_double2(cx, cy);
                    c = cc;
                    l0 = m0; l1 = m1; l2 = m2;
                }
            }
            if (!run) return;
            if (!lost && park && k == nt - 1 && c >= 0) {
                // the stored velocity of a parked last sample is 0, the scatter loop re-evaluates u(centre)
                if (S)
                    eval_p2<true>(fieldv, nodes, c, l0, l1, l2, ukx, uky);      // nodal field (global memory), rare
                else
                    eval_p2<false>(fieldv, nodes, c, l0, l1, l2, ukx, uky);
            }
            if (mu) __stcs(mu + o, make_double2(mux, muy));
            if (c >= 0) {
                hint = c;
                if (c != acc_cell) {
                    if (acc_cell >= 0) flush(acc_cell);
                    acc_cell = c;
                }
                const double gx = h * ((Dd.x - ukx) + mux), gy = h * ((Dd.y - uky) + muy);
                double phi[6];
                p2_basis(l0, l1, l2, phi);
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    acc[i] = fma(gx, phi[i], acc[i]);
                    acc[6 + i] = fma(gy, phi[i], acc[6 + i]);
                }
                if (k > 0) {
                    // continuous P1 tensor at the cell's three vertices, [g00 g01 | g10 g11] per vertex
                    double2 a0, a1, b0, b1, c0, c1;
                    if (S) {
                        int n[6];
                        load_nodes<true>(nodes, c, n);
                        a0 = gtab[2 * n[0]]; a1 = gtab[2 * n[0] + 1];
                        b0 = gtab[2 * n[1]]; b1 = gtab[2 * n[1] + 1];
                        c0 = gtab[2 * n[2]]; c1 = gtab[2 * n[2] + 1];
                    } else {
                        const double2 *gr = gtab + 6 * (size_t)c;
                        a0 = __ldg(gr); a1 = __ldg(gr + 1); b0 = __ldg(gr + 2); b1 = __ldg(gr + 3);
                        c0 = __ldg(gr + 4); c1 = __ldg(gr + 5);
                    }
                    const double G0 = l0 * a0.x + l1 * b0.x + l2 * c0.x;
                    const double G1 = l0 * a0.y + l1 * b0.y + l2 * c0.y;
                    const double G2 = l0 * a1.x + l1 * b1.x + l2 * c1.x;
                    const double G3 = l0 * a1.y + l1 * b1.y + l2 * c1.y;
                    const double rx = ex - mux, ry = ey - muy;
                    mux = mux - h * (G0 * rx + G2 * ry);
                    muy = muy - h * (G1 * rx + G3 * ry);
                }
            }
        };
        // the three streams are read D samples ahead so that their HBM latency overlaps the arithmetic
        const double2 zero2 = make_double2(0.0, 0.0);
        double2 P[D], U[D], Q[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const int k = nt - 1 - j;
            const bool ok = alive && k >= 0;
            const size_t o = (size_t)(ok ? k : 0) * K + b;
            P[j] = ok ? __ldcs(x + o) : zero2;
            U[j] = ok ? __ldcs(u + o) : zero2;
            Q[j] = ok ? __ldcs(ud + o) : zero2;
        }
        for (int k0 = nt - 1; k0 >= 0; k0 -= D) {
            double2 Pn[D], Un[D], Qn[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const int k = k0 - D - j;
                const bool ok = alive && k >= 0;
                const size_t o = (size_t)(ok ? k : 0) * K + b;
                Pn[j] = ok ? __ldcs(x + o) : zero2;
                Un[j] = ok ? __ldcs(u + o) : zero2;
                Qn[j] = ok ? __ldcs(ud + o) : zero2;
            }
#pragma unroll
            for (int j = 0; j < D; ++j)
                if (k0 - j >= 0) sample(k0 - j, P[j], U[j], Q[j]);      // (warp-uniform condition)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                P[j] = Pn[j];
                U[j] = Un[j];
                Q[j] = Qn[j];
            }
        }
        if (run && acc_cell >= 0) flush(acc_cell);
    }
    block_finish2(misfit, nmasked, scratch, counter, acc_out + 2 * (size_t)t.nn, acc_out + 2 * (size_t)t.nn + 1,
                  0.5 * h);
}

__device__ __forceinline__ void flush_sources_g(double *__restrict__ bnode, const DeviceTables &t, int c,
                                                double acc[12]) {
    int n[6];
    load_nodes_g(t, c, n);
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
template <bool X>
__global__ void __launch_bounds__(kBuoyThreads, 5)
buoy_adjoint_scatter_global_kernel(DeviceTables t, const double2 *__restrict__ vel /* per-cell records */,
                            const double2 *__restrict__ g /* per-cell vertex gradients */, int K,
                            int nt, double h, double cx, double cy, const double2 *__restrict__ x,
                            const double2 *__restrict__ u, const double2 *__restrict__ ud,
                            const double *__restrict__ mask, const uint8_t *__restrict__ parked,
                            double2 *__restrict__ mu, double *__restrict__ acc_out, double *scratch,
                            unsigned *counter, double *__restrict__ bpriv, int nrep, long long *__restrict__ digits) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    double misfit = 0.0, nmasked = 0.0;
    // point sources go to one of `nrep` private copies of b (selected by SM id) so that fp64 atomics of the
    // thousands of buoys sharing a cell do not serialise on 12 addresses; the copies are summed afterwards
    double *bdst = acc_out;
    long long *ddst = digits;
    if (nrep > 1) {
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        bdst = bpriv + (size_t)(smid % (unsigned)nrep) * (2 * (size_t)t.nn);
        if (X) ddst = digits + (size_t)(smid % (unsigned)nrep) * (8 * (size_t)t.nn);
    }
    long long *const dovf = X ? digits + 8 * (size_t)t.nn * (nrep > 1 ? nrep : 1) : nullptr;
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
            int c = locate_g(t, p.x, p.y, hint, l0, l1, l2);
            if (c < 0) {            // `except` of OCP_dolfin.py:359-361: u_x = 0, point = centre
                ukx = 0.0;
                uky = 0.0;
                p = make_double2(cx, cy);
                c = locate_g(t, cx, cy, -1, l0, l1, l2);
            } else if (park && k == nt - 1) {
                // the stored velocity of a parked last sample is 0, the scatter loop re-evaluates u(centre)
                eval_p2_cell_g(vel, c, l0, l1, l2, ukx, uky);
            }
            if (mu) __stcs(mu + o, make_double2(mux, muy));
            if (c >= 0) {
                hint = c;
                if (c != acc_cell) {
                    if (acc_cell >= 0) {
                        if (X)
                            flush_sources_exact<false>(ddst, dovf, t.cell_nodes, acc_cell, acc);
                        else
                            flush_sources_g(bdst, t, acc_cell, acc);
                    }
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
        if (acc_cell >= 0) {
            if (X)
                flush_sources_exact<false>(ddst, dovf, t.cell_nodes, acc_cell, acc);
            else
                flush_sources_g(bdst, t, acc_cell, acc);
        }
    }
    block_finish2(misfit, nmasked, scratch, counter, acc_out + 2 * (size_t)t.nn, acc_out + 2 * (size_t)t.nn + 1,
                  0.5 * h);
}

// TIME-PARALLEL backward sweep for small launches.  At K ~ 10^4 one thread per buoy leaves one warp per scheduler
// walking 200 dependent samples (~0.8 us each): the launch is bound by that chain, not by bandwidth.  The adjoint
// recursion is AFFINE in mu,
//     mu_{k-1} = (I + h G_k^T) mu_k - h G_k^T e_k,      G_k = G(x_k), e_k = u_k - u_d,k,
// so T lanes share one buoy: lane j owns a contiguous chunk of samples and
//   pass 1  composes the affine map of its chunk (point location + G per sample, no deposits),
//   scan    the incoming mu of every chunk follows from the chunks after it (T-1 shuffle steps),
//   pass 2  runs the ordinary per-sample sweep over its chunk from that incoming mu (deposits, misfit, mu output).
// Measured on B200 (cfg3, K = 10^4, in-step, L2 flushed): serial sweep 0.157 ms, this kernel 0.102 ms (ncu: 23 M warp
// instructions, issue-active 28 %, L1/TEX 86 % - the eight sample rows a warp touches per load cost wavefronts); the
// alternative mapping WARP = CHUNK (32 neighbouring buoys per warp, chunks of a buoy in different warps of a CTA, scan
// through shared memory) keeps every access coalesced but couples eight warps by a block barrier between the passes
// and measured 0.131 ms, so the lane mapping stays.
// The dependent chain shrinks from nt to 2 nt / T samples and T times as many warps are in flight.  mu at the chunk
// boundaries is formed through the composed maps, i.e. with a different (equally valid) rounding sequence than the
// serial sweep: mu and b agree with it to ~1e-15 relative (tests: 1e-12 against the oracle).
// CR (opt-in): the geometry and the projected-gradient record of the cell a lane is in stay in REGISTERS while it
// stays there (~60 samples): the fast test is the first hop of locate_g (same bary(), same strict-inside margin), so
// cells and barycentrics - hence mu and b - are bit-identical to the table path, and the nine table reads per sample
// only happen on a cell change.  Measured on B200 and NOT the default: 154 instead of 96 registers, 12 instead of 20
// warps per SM, in-step launch 0.132 vs 0.100 ms (4 lanes per buoy: 0.114, 16: 0.139).
template <int T, bool X, bool CR>
__global__ void __launch_bounds__(kBuoyThreads)
buoy_adjoint_scatter_tp_kernel(DeviceTables t, const double2 *__restrict__ vel /* per-cell records */,
                               const double2 *__restrict__ g /* per-cell vertex gradients */, int K, int nt, double h,
                               double cx, double cy, const double2 *__restrict__ x, const double2 *__restrict__ u,
                               const double2 *__restrict__ ud, const double *__restrict__ mask,
                               const uint8_t *__restrict__ parked, double2 *__restrict__ mu,
                               double *__restrict__ acc_out, double *scratch, unsigned *counter,
                               long long *__restrict__ digits) {
    constexpr int G = 32 / T;                                  // buoys per warp
    const int lane = threadIdx.x & 31, j = lane % T;           // j = chunk index, chunk T-1 holds the last samples
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long bl = warp * G + lane / T;
    const bool alive = bl < K;
    const int b = alive ? (int)bl : 0;
    const int klo = (int)(((long long)j * nt) / T), khi = (int)(((long long)(j + 1) * nt) / T) - 1;
    const bool masked = alive && mask[b] != 0.0;
    const bool park = alive && parked[b] != 0;
    const bool run = alive && !masked;
    long long *const dovf = X ? digits + 8 * (size_t)t.nn : nullptr;
    double misfit = 0.0, nmasked = (masked && j == 0) ? 1.0 : 0.0;
    // G(x_k)^T-step coefficients of one sample: returns false when the sample leaves mu unchanged
    double cgeo[6];                                            // CR: geometry / gradient record of cell `hint`
    double2 cgr[6];
    auto locate_sample = [&](int k, double2 &p, int &hint, double &l0, double &l1, double &l2, bool &lost) -> int {
        if (CR && hint >= 0) {
            bary(cgeo, p.x, p.y, l0, l1, l2);
            if (l0 > kLocateMargin && l1 > kLocateMargin && l2 > kLocateMargin) {
                lost = false;
                return hint;
            }
        }
        int c = locate_g(t, p.x, p.y, hint, l0, l1, l2);
        lost = c < 0;
        if (lost) {                                            // `except` of OCP_dolfin.py:359-361: point = centre
            p = make_double2(cx, cy);
            c = locate_g(t, cx, cy, -1, l0, l1, l2);
        }
        if (c >= 0) {
            if (CR && c != hint) {
                load_geom_g(t, c, cgeo);
                const double2 *gr = g + 6 * (size_t)c;
#pragma unroll
                for (int i = 0; i < 6; ++i) cgr[i] = __ldg(gr + i);
            }
            hint = c;
        }
        return c;
    };
    auto grad_at = [&](int c, double l0, double l1, double l2, double &G0, double &G1, double &G2, double &G3) {
        const double2 *gr = g + 6 * (size_t)c;
        // (CR: c is the cell the record in registers belongs to - locate_sample has just returned it)
        const double2 a0 = CR ? cgr[0] : __ldg(gr), a1 = CR ? cgr[1] : __ldg(gr + 1), b0 = CR ? cgr[2] : __ldg(gr + 2),
                      b1 = CR ? cgr[3] : __ldg(gr + 3), c0 = CR ? cgr[4] : __ldg(gr + 4), c1 = CR ? cgr[5] : __ldg(gr + 5);
        G0 = l0 * a0.x + l1 * b0.x + l2 * c0.x;
        G1 = l0 * a0.y + l1 * b0.y + l2 * c0.y;
        G2 = l0 * a1.x + l1 * b1.x + l2 * c1.x;
        G3 = l0 * a1.y + l1 * b1.y + l2 * c1.y;
    };
    // ---- pass 1: affine map of the chunk, mu_{klo-1} = M mu_{khi} + v
    double m00 = 1.0, m01 = 0.0, m10 = 0.0, m11 = 1.0, v0 = 0.0, v1 = 0.0;
    if (run && j > 0) {                    // (chunk 0's map is never needed)
        int hint = -1;
        // (the three streams are read one sample ahead, as in the serial sweep)
        double2 pn = __ldg(x + (size_t)khi * K + b), Un = __ldg(u + (size_t)khi * K + b), Dn = __ldg(ud + (size_t)khi * K + b);
        for (int k = khi; k >= klo; --k) {
            const size_t o = (size_t)k * K + b;
            double2 p = pn;
            const double2 U = Un, D = Dn;
            if (k > klo) {
                pn = __ldg(x + o - K);
                Un = __ldg(u + o - K);
                Dn = __ldg(ud + o - K);
            }
            double l0, l1, l2;
            bool lost;
            const int c = locate_sample(k, p, hint, l0, l1, l2, lost);
            if (c >= 0 && k > 0) {
                double G0, G1, G2, G3;
                grad_at(c, l0, l1, l2, G0, G1, G2, G3);
                const double ex = U.x - D.x, ey = U.y - D.y;
                // A = I + h G^T (row 0: [1 + h G0, h G2], row 1: [h G1, 1 + h G3]),  c = -h G^T e
                const double a00 = 1.0 + h * G0, a01 = h * G2, a10 = h * G1, a11 = 1.0 + h * G3;
                const double c0 = -h * (G0 * ex + G2 * ey), c1 = -h * (G1 * ex + G3 * ey);
                const double n00 = a00 * m00 + a01 * m10, n01 = a00 * m01 + a01 * m11;
                const double n10 = a10 * m00 + a11 * m10, n11 = a10 * m01 + a11 * m11;
                const double w0 = a00 * v0 + a01 * v1 + c0, w1 = a10 * v0 + a11 * v1 + c1;
                m00 = n00; m01 = n01; m10 = n10; m11 = n11; v0 = w0; v1 = w1;
            }
        }
    }
    // ---- scan over the chunks of a buoy, last chunk first: incoming mu of chunk s-1 = map of chunk s applied to its own
    double mux = 0.0, muy = 0.0;
#pragma unroll 1
    for (int s_ = T - 1; s_ >= 1; --s_) {
        const double ox = m00 * mux + m01 * muy + v0, oy = m10 * mux + m11 * muy + v1;
        const int src = (lane / T) * T + s_;
        const double rx = __shfl_sync(0xffffffffu, ox, src), ry = __shfl_sync(0xffffffffu, oy, src);
        if (j == s_ - 1) {
            mux = rx;
            muy = ry;
        }
    }
    // ---- pass 2: the ordinary sweep over the chunk
    int pending_cell = -1;
    double pending[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) pending[i] = 0.0;
    if (alive) {
        double acc[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) acc[i] = 0.0;
        int acc_cell = -1, hint = -1;
        double2 pn = __ldg(x + (size_t)khi * K + b), Un = __ldg(u + (size_t)khi * K + b), Dn = __ldg(ud + (size_t)khi * K + b);
        for (int k = khi; k >= klo; --k) {
            const size_t o = (size_t)k * K + b;
            double2 p = pn;
            const double2 U = Un, D = Dn;
            if (k > klo) {
                pn = __ldg(x + o - K);
                Un = __ldg(u + o - K);
                Dn = __ldg(ud + o - K);
            }
            const double ex = U.x - D.x, ey = U.y - D.y;
            misfit += ex * ex + ey * ey;
            if (masked) {
                if (mu) __stcs(mu + o, make_double2(0.0, 0.0));
                continue;
            }
            double l0, l1, l2;
            double ukx = U.x, uky = U.y;
            bool lost;
            const int c = locate_sample(k, p, hint, l0, l1, l2, lost);
            if (lost) {
                ukx = 0.0;
                uky = 0.0;
            } else if (park && k == nt - 1) {
                // the stored velocity of a parked last sample is 0, the scatter loop re-evaluates u(centre)
                eval_p2_cell_g(vel, c, l0, l1, l2, ukx, uky);
            }
            if (mu) __stcs(mu + o, make_double2(mux, muy));
            if (c >= 0) {
                if (c != acc_cell) {
                    if (acc_cell >= 0) {
                        if (X)
                            flush_sources_exact<false>(digits, dovf, t.cell_nodes, acc_cell, acc);
                        else
                            flush_sources_g(acc_out, t, acc_cell, acc);
                    }
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
                    double G0, G1, G2, G3;
                    grad_at(c, l0, l1, l2, G0, G1, G2, G3);
                    const double rx = ex - mux, ry = ey - muy;
                    mux = mux - h * (G0 * rx + G2 * ry);
                    muy = muy - h * (G1 * rx + G3 * ry);
                }
            }
        }
        // The T lanes of a buoy mostly end in the same cell: their pending deposits are combined by a shuffle tree
        // (lane j takes over lane j + d when both hold the same cell) so that a buoy flushes once instead of T times.
        // [all lanes of the warp reach this point: `alive` lanes only differ in the trip counts above]
        pending_cell = acc_cell;
#pragma unroll
        for (int i = 0; i < 12; ++i) pending[i] = acc[i];
    }
    for (int dlt = 1; dlt < T; dlt <<= 1) {
        const int other_cell = __shfl_down_sync(0xffffffffu, pending_cell, dlt);
        const int below_cell = __shfl_up_sync(0xffffffffu, pending_cell, dlt);
        const bool take = (j % (2 * dlt) == 0) && (j + dlt < T) && pending_cell >= 0 && other_cell == pending_cell;
        const bool give = (j % (2 * dlt) == dlt) && pending_cell >= 0 && below_cell == pending_cell;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const double v = __shfl_down_sync(0xffffffffu, pending[i], dlt);
            if (take) pending[i] += v;
        }
        if (give) pending_cell = -1;
    }
    if (pending_cell >= 0) {
        if (X)
            flush_sources_exact<false>(digits, dovf, t.cell_nodes, pending_cell, pending);
        else
            flush_sources_g(acc_out, t, pending_cell, pending);
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

// digit sums of the private copies into copy 0 (integer additions: exact, order-free)
__global__ void reduce_digits_kernel(size_t n, int nrep, long long *__restrict__ digits) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long s = digits[i];
    for (int r = 1; r < nrep; ++r) s += digits[(size_t)r * n + i];
    digits[i] = s;
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

// upper bound of the CTAs of one sweep launch (sizes the partial-sum scratch): serial sweep K / 128, time-parallel
// sweep (K <= 12000) at most 32 lanes per buoy
int buoy_max_blocks(int K) { return std::max(3200, (K + kBuoyThreads - 1) / kBuoyThreads); }

int buoy_private_copies(int K, int nc, int nn) {
    if ((long long)K < 8LL * nc) return 1;                        // few buoys per cell: no contention to avoid
    const long long cap = (256LL << 20) / (16LL * nn);            // at most 256 MiB of private copies
    return (int)std::max(1LL, std::min(148LL, cap));
}

size_t buoy_exact_digits(int nn, int nrep) { return 8 * (size_t)nn * (size_t)std::max(nrep, 1) + 1; }

namespace {

int device_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// Persistent launch shape of the staged kernels: one CTA per SM.  Small launches (K <= #SMs x 512) give every CTA
// ONE contiguous share of ceil(K / #SMs) buoys (threads = that share rounded up to whole warps) so that all SMs hold
// work - at K = 10^4 that is 148 CTAs of 68 buoys instead of 79 CTAs of 128; large launches loop over 512-buoy blocks.
struct StagedShape {
    int grid, threads, per_block;
};

StagedShape staged_shape(int K) {
    const int sms = device_sms();
    StagedShape sh;
    if ((long long)K <= (long long)sms * kStagedMaxThreads) {
        int per = std::max(32, (K + sms - 1) / sms);
        per = (per + 3) & ~3;                                  // 64-byte aligned shares of the (nt, K, 2) arrays
        sh.per_block = per;
        sh.threads = (per + 31) & ~31;
        sh.grid = (K + per - 1) / per;
    } else {
        sh.per_block = sh.threads = kStagedMaxThreads;
        sh.grid = sms;
    }
    return sh;
}

template <class Kern>
bool allow_smem(Kern k, size_t bytes) {
    return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess;
}

}  // namespace

// Lanes per buoy of the time-parallel backward sweep: 32 for the thesis' buoy counts (<= 2000), 8 up to the 10 000-buoy
// run, 1 (serial sweep, bandwidth-bound) beyond.  OCP_BUOY_TP=0 forces the serial sweep.
int time_parallel_lanes(int K, int nt, int nrep) {
    static int enabled = -1, forced = 0;
    if (enabled < 0) {
        const char *e = getenv("OCP_BUOY_TP");
        enabled = (e && atoi(e) == 0) ? 0 : 1;
        if (const char *l = getenv("OCP_BUOY_TP_LANES")) {
            const int v = atoi(l);
            forced = (v == 32 || v == 16 || v == 8 || v == 4) ? v : 0;
        }
    }
    if (!enabled || nrep > 1 || nt < 64) return 1;
    if (K > 12000) return 1;
    if (forced) return forced;
    return K <= 2000 ? 32 : 8;
}

bool buoy_tables_fit_shared(int nc, int nn, int nv) {
    return staged_bytes(nc, 16 * (size_t)nn) <= kSmemBudget && staged_bytes(nc, 32 * (size_t)nv) <= kSmemBudget;
}

void launch_cell_records(const DeviceTables &t, const double *vel, double *cellvel, const double *g, double *cellg,
                         cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    cell_records_kernel<<<(t.nc * 6 + 255) / 256, 256, 0, s>>>(
        t.nc, t.cell_nodes, reinterpret_cast<const double2 *>(vel), reinterpret_cast<double2 *>(cellvel),
        reinterpret_cast<const double2 *>(g), reinterpret_cast<double2 *>(cellg));
}

void launch_buoy_forward(const DeviceTables &t, bool staged, const double *field, const double *x0, int K, int nt,
                         double h, double cx, double cy, double *x, double *u, int *cell, double *mask,
                         uint8_t *parked, cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    if (K <= 0) return;
    const double2 *f2 = reinterpret_cast<const double2 *>(field), *x02 = reinterpret_cast<const double2 *>(x0);
    double2 *x2 = reinterpret_cast<double2 *>(x), *u2 = reinterpret_cast<double2 *>(u);
    if (staged) {
        const size_t smem = staged_bytes(t.nc, 16 * (size_t)t.nn);
        static bool ok = allow_smem(buoy_forward_kernel<true>, kSmemBudget);
        (void)ok;
        const StagedShape sh = staged_shape(K);
        buoy_forward_kernel<true><<<sh.grid, sh.threads, smem, s>>>(t, f2, x02, K, sh.per_block, nt, h, cx, cy, x2, u2,
                                                                    cell, mask, parked);
    } else {
        // small launches: cell data in registers (0.075 -> 0.065 ms at K = 10^4); OCP_BUOY_FWD_BLOCK = 32 / 64 spreads the
        // warps over all SMs and measured no faster than 128-thread CTAs
        static const bool cached = !(getenv("OCP_BUOY_FWD_CACHE") && atoi(getenv("OCP_BUOY_FWD_CACHE")) == 0);
        static const int small_block = getenv("OCP_BUOY_FWD_BLOCK") ? std::max(32, std::min(kBuoyThreads, atoi(getenv("OCP_BUOY_FWD_BLOCK")) & ~31)) : kBuoyThreads;
        if (cached && K <= 12000)
            buoy_forward_cached_kernel<<<(K + small_block - 1) / small_block, small_block, 0, s>>>(
                t, f2, x02, K, nt, h, cx, cy, x2, u2, cell, mask, parked);
        else
            buoy_forward_global_kernel<<<(K + kBuoyThreads - 1) / kBuoyThreads, kBuoyThreads, 0, s>>>(
                t, f2, x02, K, nt, h, cx, cy, x2, u2, cell, mask, parked);
    }
}

void launch_buoy_adjoint_scatter(const DeviceTables &t, bool staged, const double *fieldv, const double *fieldg, int K,
                                 int nt, double h, double cx, double cy, const double *x, const double *u,
                                 const double *ud, const double *mask, const uint8_t *parked, double *mu, double *acc,
                                 double *scratch, unsigned *counter, double *bpriv, int nrep, long long *digits,
                                 cudaStream_t s) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    if (K <= 0) return;
    const bool exact = digits != nullptr;
    if (exact)
        cudaMemsetAsync(digits, 0, sizeof(long long) * buoy_exact_digits(t.nn, nrep), s);
    else if (nrep > 1)
        cudaMemsetAsync(bpriv, 0, sizeof(double) * 2 * (size_t)t.nn * nrep, s);
    const double2 *fv = reinterpret_cast<const double2 *>(fieldv), *fg = reinterpret_cast<const double2 *>(fieldg);
    const double2 *x2 = reinterpret_cast<const double2 *>(x), *u2 = reinterpret_cast<const double2 *>(u),
                  *d2 = reinterpret_cast<const double2 *>(ud);
    double2 *mu2 = reinterpret_cast<double2 *>(mu);
#define OCP_BWD_ARGS t, fv, fg, K, per, nt, h, cx, cy, x2, u2, d2, mask, parked, mu2, acc, scratch, counter, bpriv, nrep, digits
    if (staged) {
        const size_t smem = staged_bytes(t.nc, 32 * (size_t)t.nv);
        static bool ok = allow_smem(buoy_adjoint_scatter_kernel<true, kDepthLarge, false>, kSmemBudget) &&
                         allow_smem(buoy_adjoint_scatter_kernel<true, kDepthSmall, false>, kSmemBudget) &&
                         allow_smem(buoy_adjoint_scatter_kernel<true, kDepthLarge, true>, kSmemBudget) &&
                         allow_smem(buoy_adjoint_scatter_kernel<true, kDepthSmall, true>, kSmemBudget);
        (void)ok;
        const StagedShape sh = staged_shape(K);
        const int per = sh.per_block;
        if (sh.threads <= 128) {       // small launch: one warp per scheduler must hide the stream latency itself
            if (exact)
                buoy_adjoint_scatter_kernel<true, kDepthSmall, true><<<sh.grid, sh.threads, smem, s>>>(OCP_BWD_ARGS);
            else
                buoy_adjoint_scatter_kernel<true, kDepthSmall, false><<<sh.grid, sh.threads, smem, s>>>(OCP_BWD_ARGS);
        } else {
            if (exact)
                buoy_adjoint_scatter_kernel<true, kDepthLarge, true><<<sh.grid, sh.threads, smem, s>>>(OCP_BWD_ARGS);
            else
                buoy_adjoint_scatter_kernel<true, kDepthLarge, false><<<sh.grid, sh.threads, smem, s>>>(OCP_BWD_ARGS);
        }
    } else if (time_parallel_lanes(K, nt, nrep) > 1) {
        // small launch: T lanes per buoy (see buoy_adjoint_scatter_tp_kernel); no private copies at these sizes
        const int T = time_parallel_lanes(K, nt, nrep);
        const long long threads = (long long)((K + (32 / T) - 1) / (32 / T)) * 32;
        const int grid = (int)((threads + kBuoyThreads - 1) / kBuoyThreads);
        // (opt-in, OCP_BUOY_TP_CACHE=1: measured slower - 0.132 vs 0.100 ms at K = 10^4 - the 154 registers of the cached
        // variant leave 12 instead of 20 warps per SM)
        static const bool cached = getenv("OCP_BUOY_TP_CACHE") && atoi(getenv("OCP_BUOY_TP_CACHE")) != 0;
#define OCP_TP_ARGS t, fv, fg, K, nt, h, cx, cy, x2, u2, d2, mask, parked, mu2, acc, scratch, counter, digits
#define OCP_TP_LAUNCH(TT)                                                                                       \
    do {                                                                                                        \
        if (exact && cached) buoy_adjoint_scatter_tp_kernel<TT, true, true><<<grid, kBuoyThreads, 0, s>>>(OCP_TP_ARGS);    \
        else if (exact) buoy_adjoint_scatter_tp_kernel<TT, true, false><<<grid, kBuoyThreads, 0, s>>>(OCP_TP_ARGS);        \
        else if (cached) buoy_adjoint_scatter_tp_kernel<TT, false, true><<<grid, kBuoyThreads, 0, s>>>(OCP_TP_ARGS);       \
        else buoy_adjoint_scatter_tp_kernel<TT, false, false><<<grid, kBuoyThreads, 0, s>>>(OCP_TP_ARGS);                  \
    } while (0)
        if (T == 32) OCP_TP_LAUNCH(32);
        else if (T == 16) OCP_TP_LAUNCH(16);
        else if (T == 4) OCP_TP_LAUNCH(4);
        else OCP_TP_LAUNCH(8);
#undef OCP_TP_LAUNCH
#undef OCP_TP_ARGS
    } else {
        const int grid = (K + kBuoyThreads - 1) / kBuoyThreads;
        if (exact)
            buoy_adjoint_scatter_global_kernel<true><<<grid, kBuoyThreads, 0, s>>>(
                t, fv, fg, K, nt, h, cx, cy, x2, u2, d2, mask, parked, mu2, acc, scratch, counter, bpriv, nrep, digits);
        else
            buoy_adjoint_scatter_global_kernel<false><<<grid, kBuoyThreads, 0, s>>>(
                t, fv, fg, K, nt, h, cx, cy, x2, u2, d2, mask, parked, mu2, acc, scratch, counter, bpriv, nrep, digits);
    }
#undef OCP_BWD_ARGS
    if (exact) {
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        if (nrep > 1) {
            const size_t n = 8 * (size_t)t.nn;
            g_launch_count.fetch_add(1, std::memory_order_relaxed);
            reduce_digits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, nrep, digits);
            // the overflow counter of the copies sits behind the last copy: move it behind copy 0 for the finalize
            cudaMemcpyAsync(digits + n, digits + n * nrep, sizeof(long long), cudaMemcpyDeviceToDevice, s);
        }
        finalize_exact_kernel<<<(2 * t.nn + 255) / 256, 256, 0, s>>>(2 * t.nn, digits, acc);
    } else if (nrep > 1) {
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
