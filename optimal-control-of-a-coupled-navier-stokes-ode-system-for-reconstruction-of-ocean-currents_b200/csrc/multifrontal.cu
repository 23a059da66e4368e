// GPU numeric phase of the multifrontal LU (see multifrontal.hpp).  Fronts are dense column-major m x m blocks in one
// HBM workspace (L2 resident for the reference meshes); one launch per level of the nested-dissection tree, the whole
// factorisation / solve replayed as a CUDA graph.
//   small fronts (order <= 768): one thread-block CLUSTER per front - extend-add of the children's Schur complements,
//     then a right-looking blocked LU of the np fully-summed columns (16-column panel in shared memory / registers,
//     4 x 4 register tiles for the trailing update, one cluster barrier per panel);
//   large fronts (refined meshes): a GROUP of co-resident CTAs per front working out of L2 / HBM with 32-column
//     panels and group barriers in global memory (mf_big_factor_kernel).
// Pivoting is static (see the comment above mf_factor_kernel); the triangular solves work in 16-row blocks with the
// inverses of the diagonal blocks (mf_dinv_kernel).
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>

#include "kernels.cuh"
#include "multifrontal.cuh"

namespace cg = cooperative_groups;

namespace ocp {

namespace {

// factor-kernel variant chosen in "auto" mode (measured, profiles/README.md round 2)
#define MF_AUTO_VARIANT(max_m, legacy) (legacy)

constexpr int NB = 16;     // panel width
constexpr int CWO = 4;     // column-ownership granularity inside a cluster (= tile width)

struct MFDev {
    const int *m, *np, *first, *idx_ptr, *idx, *child_ptr, *child, *rel_ptr, *rel;
    const long long *front_ptr;
    double *F;
    int *piv;
    double *dinv;              // per 16-column panel: inverse of the unit-lower and of the upper diagonal block (2 x 256)
    const int *dinv_ptr;       // per front: index of its first panel in dinv
    double *dinv64;            // per 64-pivot block of a SMALL front: [L11^-1 | U11^-1], each 64 x 64 row-major
    const int *dinv64_ptr;     // per front: index of its first 64-pivot block in dinv64 (small fronts only)
    int ea_atomic;             // small fronts: extend-add with fp64 atomics (OCP_MF_EA_ATOMIC=1) instead of child-by-child
};

__global__ void scatter_values_kernel(int nnz, const long long *__restrict__ dest, const double *__restrict__ vals,
                                      double *__restrict__ F) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) F[dest[k]] = vals[k];
}

// One CLUSTER of C CTAs per front (C = 1 for the many small fronts at the bottom of the tree, up to 16 for the few
// large fronts near the root), one launch per tree level.
//
// Static pivoting: inside a front velocities precede pressures and every pressure dof sits above all velocities it
// couples to (mf_analyse), so the velocity block is positive definite and each pressure pivot is a negative
// definite Schur complement - LU in the given order is stable for the diffusion-dominated systems of this path
// (validated against partial pivoting on the host, tests/test_capi_cpu.py) and nothing has to be searched or swapped.
//
// Per 16-column panel every CTA of the cluster redundantly (a) loads the panel, one row per thread, in registers,
// (b) factors the 16x16 diagonal block inside warp 0 with shuffles, (c) turns its rows into L = A U11^{-1};
// then it solves U12 and updates the trailing matrix only for the columns it owns (absolute column / 16 mod C).
// One cluster barrier per panel publishes the updated columns.  Front accesses bypass L1 (ld.cg / st.cg).
// 1/x on the serial critical path of the diagonal block: hardware approximation (rcp.approx.ftz.f64, ~20 bits)
// plus two Newton steps (error < 1 ulp) - four dependent DFMAs instead of the IEEE division subroutine.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}


// Inverses of the 16 x 16 diagonal blocks of the factors (unit-lower L11 and upper U11 of every 16-pivot block), used
// by the blocked triangular solves.  One launch after the numeric factorisation, one warp per (block, factor): the
// inversions are shuffle-heavy and used to sit on the critical path of every panel.
__global__ void __launch_bounds__(256)
mf_dinv_kernel(MFDev d, const int *__restrict__ panel_node, int npanels, int *info, int check_only) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int p = w >> 1, which = w & 1;                 // which = 1: L11^{-1}, 0: U11^{-1}
    if (p >= npanels) return;
    // check_only: every front is solved with the 64-row kernels (own inverses), so only the growth check below is
    // wanted from this launch
    if (check_only && which) return;
    const int s = panel_node[p];
    const int m = d.m[s], np = d.np[s];
    const int k0 = (p - d.dinv_ptr[s]) * NB, kb = min(NB, np - k0);
    const double *F = d.F + d.front_ptr[s];
    double r[NB];                                        // row `lane` of the factored block
#pragma unroll
    for (int c = 0; c < NB; ++c) r[c] = (lane < kb && c < kb) ? __ldcg(F + (k0 + lane) + (size_t)(k0 + c) * m) : 0.0;
    // Growth check of the static pivoting, off the critical path of the factorisation: a pivot that is tiny against the
    // rest of its row of U11 (relative 1e-11) marks the factorisation as unreliable - reported like a zero pivot
    // (OCP_ERR_SOLVER) instead of relying on the residual checks downstream.
    if (!which && lane < kb) {
        double rowmax = 0.0, piv = 0.0;
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            if (c == lane) piv = fabs(r[c]);
            if (c >= lane && c < kb) rowmax = fmax(rowmax, fabs(r[c]));
        }
        if (!(piv > 1e-11 * rowmax)) atomicExch(info, s + 1);
    }
    if (check_only) return;
    double X[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) X[c] = (c == lane) ? 1.0 : 0.0;
    if (which) {
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            const double lik = (lane > k && lane < kb && k < kb) ? r[k] : 0.0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (c <= k) {
                    const double xkc = __shfl_sync(0xffffffffu, X[c], k);
                    X[c] = fma(-lik, xkc, X[c]);
                }
            }
        }
    } else {
        double diag = 1.0;
#pragma unroll
        for (int c = 0; c < NB; ++c)
            if (c == lane && lane < kb) diag = r[c];
        const double rdiag = fast_rcp(diag);             // same reciprocal as the factorisation used
#pragma unroll
        for (int k = NB - 1; k >= 0; --k) {
            const bool live = k < kb;
            if (lane == k && live) {
#pragma unroll
                for (int c = 0; c < NB; ++c) X[c] *= rdiag;
            }
            const double uik = (lane < k && live) ? r[k] : 0.0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (c >= k) {
                    const double ykc = __shfl_sync(0xffffffffu, X[c], k);
                    X[c] = fma(-uik, ykc, X[c]);
                }
            }
        }
    }
    if (lane < NB) {
        double *dst = d.dinv + (size_t)p * (2 * NB * NB) + (which ? 0 : NB * NB) + lane * NB;
#pragma unroll
        for (int c = 0; c < NB; ++c) dst[c] = X[c];
    }
}

// Split-phase cluster barrier and a barrier over the worker warps only (the look-ahead warp does not take part).
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

// LU of a 16 x 16 block held one row per lane (rows >= kb are identity padding), static pivoting.  The dependent chain
// per step is pivot -> reciprocal -> multiplier -> update of the NEXT pivot, so the next pivot column is updated first
// and its reciprocal is started before the remaining columns of the step are touched; the pivot row travels through
// shared memory (lane j stores it once, everybody reads it back as broadcasts).
__device__ __forceinline__ void diag16_factor(double (&r)[NB], int lane, int kb, double (*prow2)[NB], double *rd, int *info,
                                              int front) {
    double rinv = fast_rcp(__shfl_sync(0xffffffffu, r[0], 0));
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        double *prow = prow2[j & 1];
        if (lane == j) {
            rd[j] = rinv;
            if (j < kb && !(fabs(rinv) < 1e280)) atomicExch(info, front + 1);
#pragma unroll
            for (int jj = 0; jj < NB; jj += 2)
                if (jj + 1 > j) *reinterpret_cast<double2 *>(prow + jj) = make_double2(r[jj], r[jj + 1]);
        }
        __syncwarp();
        const bool below = lane > j;
        const double l = r[j] * rinv;
        if (below) r[j] = l;
        double rnext = 0.0;
        if (j + 1 < NB) {
            const double u1 = prow[j + 1];
            if (below) r[j + 1] = fma(-l, u1, r[j + 1]);
            rnext = fast_rcp(__shfl_sync(0xffffffffu, r[j + 1], j + 1));
        }
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) {
            if (jj > j + 1) {
                const double u = prow[jj];
                if (below) r[jj] = fma(-l, u, r[jj]);
            }
        }
        rinv = rnext;
    }
}

// Measured and NOT kept (round 2, tools/microbench/lrows.cu, OCP_MF_PROF): the rows below the block as a product with an
// explicit U11^-1 instead of the 16-step recurrence.  In isolation 4.2 k -> 1.8 k cycles per panel (256 rows, one
// CTA); in the kernel the row phase of the root front dropped from 6.8 k to 4.2 k cycles per panel, but forming the
// inverse costs 3.2 k cycles of one warp per panel: on the look-ahead warp it lengthened that (co-critical) path
// (factorisation 0.867 -> 0.899 ms), on the first worker warp behind the panel loads it was not hidden either
// (0.914 ms).  The recurrence stays.
// TF threads = TF - 32 workers + one LOOK-AHEAD warp.  While the workers run panel k (rows below the block, U12, trailing
// update), the look-ahead warp of every CTA produces the factored diagonal block of panel k+1 on its own: it reads the
// 32 x 32 corner [A11 A12; A21 A22] of the trailing matrix right after the cluster barrier, forms U12' = L11^{-1} A12,
// L21' = A21 U11^{-1} and A22 - L21' U12' with exactly the arithmetic of phases (d), (c), (e) below (so the result is
// bit-identical to what the trailing update stores) and factors it.  The serial 16-step elimination - the longest
// single-warp section of a panel - is thereby off the critical path.  All global stores of a panel are issued after
// the wait (acquire) of a split-phase cluster barrier; the look-ahead warps arrive (release) once they hold their
// corner in registers and the workers once their panel loads are issued, so no CTA can overwrite an entry another
// CTA still has to read.
template <int TF, int RMAX>
__global__ void __launch_bounds__(TF, 1)
mf_factor_kernel(MFDev d, const int *__restrict__ nodes, int max_m, int *info, long long *prof) {
    constexpr int TW = TF - 32;                   // worker threads
    extern __shared__ double sm[];
    __shared__ double s_D[2][NB][NB + 1];         // factored diagonal block of the current / next panel: L below, U on/above
    __shared__ double s_rd[2][NB];                // reciprocals of its diagonal
    __shared__ __align__(16) double s_prow[2][NB];   // pivot row of the current / next elimination step
    __shared__ double s_T[NB][NB + 1];            // U12' of the look-ahead corner
    cg::cluster_group cl = cg::this_cluster();
    const int C = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int ldu = (max_m + CWO + 3) & ~3;       // leading dimension of the U12 staging area (multiple of 4)
    double *P = sm;                               // L panel, ld = mp
    double *Uc = sm + (size_t)((max_m + 3) & ~3) * NB;   // NB x ldu, row-major, indexed by "own column" number
    const int s = nodes[blockIdx.x / C];
    const int m = d.m[s], np = d.np[s], tid = threadIdx.x;
    const bool worker = tid < TW;
    const int lane = tid & 31;
    double *F = d.F + d.front_ptr[s];
    // ---- extend-add the children's update matrices.  Two children may hit the same entry of the front, one child never
    // hits an entry twice (its index map is injective): the children go one at a time with a cluster barrier in between,
    // so every entry has a single writer and plain read-modify-writes do - eight independent ones in flight per thread,
    // the (column, row) pairs of the CTA's columns flattened over its threads.  fp64 atomics (d.ea_atomic, the round-1
    // form) top out near 50 G/s on this chip and cost the levels 1-4 of the 32 x 32 tree 15-40 us each; the plain
    // form also fixes the summation order (A + child 1 + child 2): the factors are bit-reproducible.
    for (int ci = d.child_ptr[s]; ci < d.child_ptr[s + 1]; ++ci) {
        const int c = d.child[ci], mc = d.m[c], npc = d.np[c], nu = mc - npc;
        const double *Fc = d.F + d.front_ptr[c];
        const int *rel = d.rel + d.rel_ptr[c];
        if (d.ea_atomic) {
            for (int j = rank; j < nu; j += C) {
                const double *src = Fc + npc + (size_t)(npc + j) * mc;
                double *dst = F + (size_t)__ldg(rel + j) * m;
                for (int i = tid; i < nu; i += TF) atomicAdd(dst + __ldg(rel + i), __ldcg(src + i));
            }
            continue;
        }
        const int ncol = nu > rank ? (nu - rank + C - 1) / C : 0;     // columns rank, rank + C, ... of the update matrix
        const int total = ncol * nu;
        constexpr int EU = 8;
        for (int e0 = tid; e0 < total; e0 += EU * TF) {
            double *p[EU];
            double v[EU];
#pragma unroll
            for (int u = 0; u < EU; ++u) {
                const int e = e0 + u * TF;
                p[u] = nullptr;
                v[u] = 0.0;
                if (e < total) {
                    const int q = e / nu, i = e - q * nu, j = rank + q * C;
                    p[u] = F + (size_t)__ldg(rel + j) * m + __ldg(rel + i);
                    v[u] = __ldcg(Fc + npc + i + (size_t)(npc + j) * mc);
                }
            }
#pragma unroll
            for (int u = 0; u < EU; ++u)
                if (p[u]) v[u] += __ldcg(p[u]);
#pragma unroll
            for (int u = 0; u < EU; ++u)
                if (p[u]) __stcg(p[u], v[u]);
        }
        if (ci + 1 < d.child_ptr[s + 1]) cl.sync();     // the next child may hit entries another CTA has just written
    }
    cl.sync();
    if (np == 0) return;
    long long tp0 = 0, acc_panel = 0, acc_trail = 0, acc_sync = 0, acc_load = 0, acc_diag = 0, acc_trsm = 0;
#define MF_TICK(accv) do { if (prof && tid == 0) { long long t_ = clock64(); accv += t_ - tp0; tp0 = t_; } } while (0)
    if (prof && tid == 0) tp0 = clock64();
    // ---- prologue: the look-ahead warp factors the first diagonal block
    if (!worker) {
        const int kb = min(NB, np);
        double r[NB];
#pragma unroll
        for (int jj = 0; jj < NB; ++jj)
            r[jj] = (lane < kb && jj < kb) ? __ldcg(F + lane + (size_t)jj * m) : ((jj == lane) ? 1.0 : 0.0);
        diag16_factor(r, lane, kb, s_prow, s_rd[0], info, s);
        if (lane < NB) {
#pragma unroll
            for (int jj = 0; jj < NB; ++jj) s_D[0][lane][jj] = r[jj];
        }
    }
    __syncthreads();
    MF_TICK(acc_diag);
    // own columns: chunk q covers absolute columns [(rank + q C) 4, +4); "own column number" = 4 q + (c % 4)
    const int nown = (m > rank * CWO) ? (m - rank * CWO + C * CWO - 1) / (C * CWO) : 0;
    for (int k0 = 0, cur = 0; k0 < np; k0 += NB, cur ^= 1) {
        const int kb = min(NB, np - k0), mp = m - k0, ctrail = k0 + kb;
        const int ldp = (mp + 3) & ~3;          // panel leading dimension, multiple of 4 so that tiles load as 2 x 16 B
        const double (*D)[NB + 1] = s_D[cur];
        const double *rd = s_rd[cur];
        if (!worker) {
            // ================= look-ahead warp: diagonal block of the NEXT panel =================
            const int k1 = k0 + NB;
            const bool more = k1 < np;                       // (then kb == NB)
            const int kb2 = more ? min(NB, np - k1) : 0;
            double r[NB];                                    // row `lane` of A22, then of its factors
            if (more) {
                double u[NB], a21[NB];
                // lane = column j of A12 (16 rows of the current pivot block), lane = row i of A21 and A22
#pragma unroll
                for (int t = 0; t < NB; ++t) u[t] = (lane < kb2) ? __ldcg(F + (k0 + t) + (size_t)(k1 + lane) * m) : 0.0;
#pragma unroll
                for (int jj = 0; jj < NB; ++jj) {
                    a21[jj] = (lane < kb2) ? __ldcg(F + (k1 + lane) + (size_t)(k0 + jj) * m) : 0.0;
                    r[jj] = (lane < kb2 && jj < kb2) ? __ldcg(F + (k1 + lane) + (size_t)(k1 + jj) * m) : 0.0;
                }
                // U12' = L11^{-1} A12 (phase d arithmetic), staged in shared memory
#pragma unroll
                for (int t = 0; t < NB; ++t) {
#pragma unroll
                    for (int tt = 0; tt < NB; ++tt)
                        if (tt > t) u[tt] = fma(-D[tt][t], u[t], u[tt]);
                }
                if (lane < NB) {
#pragma unroll
                    for (int t = 0; t < NB; ++t) s_T[t][lane] = u[t];
                }
                // L21' = A21 U11^{-1} (phase c arithmetic)
#pragma unroll
                for (int t = 0; t < NB; ++t) {
                    const double l = a21[t] * rd[t];
                    a21[t] = l;
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj)
                        if (jj > t) a21[jj] = fma(-l, D[t][jj], a21[jj]);
                }
                __syncwarp();
                // A22 - L21' U12' (phase e arithmetic: accumulate over t, subtract once)
#pragma unroll
                for (int jj = 0; jj < NB; ++jj) {
                    double acc = 0.0;
#pragma unroll
                    for (int t = 0; t < NB; ++t) acc = fma(a21[t], s_T[t][jj], acc);
                    r[jj] = r[jj] - acc;
                }
                // identity padding of a partial last block
#pragma unroll
                for (int jj = 0; jj < NB; ++jj)
                    if (lane >= kb2 || jj >= kb2) r[jj] = (jj == lane) ? 1.0 : 0.0;
            }
            // every value read from the front is consumed above: its loads have completed
            cluster_arrive();
            if (more) {
                diag16_factor(r, lane, kb2, s_prow, s_rd[cur ^ 1], info, s);
                if (lane < NB) {
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) s_D[cur ^ 1][lane][jj] = r[jj];
                }
            }
            cluster_wait();
        } else {
            // ================= workers: panel k =================
            // (a) panel rows into registers
            double a[RMAX][NB];
#pragma unroll
            for (int q = 0; q < RMAX; ++q) {
                const int i = tid + q * TW;
#pragma unroll
                for (int jj = 0; jj < NB; ++jj)
                    a[q][jj] = (i < mp && jj < kb) ? __ldcg(F + (k0 + i) + (size_t)(k0 + jj) * m) : 0.0;
            }
            // arrive with release semantics AFTER the panel loads: CTA 0's store of the L panel into these very entries
            // comes after its wait (acquire) below, hence after every worker's loads above
            cluster_arrive();
            if (prof) { worker_sync<TW>(); MF_TICK(acc_load); }
            // (d) U12 = L11^{-1} F12 for the own trailing columns.  Runs on the highest worker ids, whose panel rows
            // (phase c) mostly do not exist (m < TW), so this serial 16-step substitution overlaps phase (c).
            for (int idx = TW - 1 - tid; idx < nown * CWO; idx += TW) {
                const int c = (rank + (idx / CWO) * C) * CWO + (idx % CWO);
                if (c >= ctrail && c < m) {
                    const double *colp = F + (size_t)c * m + k0;
                    double u[NB];
#pragma unroll
                    for (int t = 0; t < NB; ++t) u[t] = t < kb ? __ldcg(colp + t) : 0.0;
#pragma unroll
                    for (int t = 0; t < NB; ++t) {
                        if (t < kb) {
#pragma unroll
                            for (int tt = 0; tt < NB; ++tt)
                                if (tt > t && tt < kb) u[tt] = fma(-D[tt][t], u[t], u[tt]);
                        }
                    }
#pragma unroll
                    for (int t = 0; t < NB; ++t) Uc[t * ldu + idx] = u[t];
                }
            }
            if (prof) { worker_sync<TW>(); MF_TICK(acc_trsm); }
            // (c) rows below the block: L = A U11^{-1}; every row goes to the shared panel
#pragma unroll
            for (int q = 0; q < RMAX; ++q) {
                const int i = tid + q * TW;
                if (i < mp) {
                    if (i >= kb) {
#pragma unroll
                        for (int t = 0; t < NB; ++t) {
                            if (t < kb) {
                                const double l = a[q][t] * rd[t];
                                a[q][t] = l;
#pragma unroll
                                for (int jj = 0; jj < NB; ++jj)
                                    if (jj > t && jj < kb) a[q][jj] = fma(-l, D[t][jj], a[q][jj]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int jj = 0; jj < NB; ++jj) a[q][jj] = D[i][jj];     // the factored block itself
                    }
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj)
                        if (jj < kb) P[i + jj * ldp] = a[q][jj];
                }
            }
            worker_sync<TW>();
            MF_TICK(acc_panel);
            // every look-ahead warp of the cluster holds its corner: the front may be written now
            cluster_wait();
            if (rank == 0) {
#pragma unroll
                for (int q = 0; q < RMAX; ++q) {
                    const int i = tid + q * TW;
                    if (i < mp) {
#pragma unroll
                        for (int jj = 0; jj < NB; ++jj)
                            if (jj < kb) __stcg(F + (k0 + i) + (size_t)(k0 + jj) * m, a[q][jj]);
                    }
                }
            }
            for (int idx = TW - 1 - tid; idx < nown * CWO; idx += TW) {
                const int c = (rank + (idx / CWO) * C) * CWO + (idx % CWO);
                if (c >= ctrail && c < m) {
                    double *colp = F + (size_t)c * m + k0;
#pragma unroll
                    for (int t = 0; t < NB; ++t)
                        if (t < kb) __stcg(colp + t, Uc[t * ldu + idx]);
                }
            }
            // (e) trailing update of the own columns, 4x4 register tiles
            const int nrow = m - ctrail;
            if (nrow > 0 && nown > 0) {
                // first own chunk that still has trailing columns
                int q0 = 0;
                while (q0 < nown && (rank + q0 * C) * CWO + CWO <= ctrail) ++q0;
                const int rbase = kb & ~3;                   // tile origin aligned to 4 rows (kb < 16 only in a front's last panel)
                const int tr = (mp - rbase + 3) >> 2, tc = (nown - q0) * (CWO / 4);
                for (int tile = tid; tile < tr * tc; tile += TW) {
                    const int ti = tile % tr, tj = tile / tr;
                    const int i0 = rbase + 4 * ti, idx0 = q0 * CWO + 4 * tj;
                    const int c0 = (rank + (idx0 / CWO) * C) * CWO + (idx0 % CWO);
                    double f[4][4], acc[4][4];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const double *colp = F + (size_t)(c0 + b) * m + k0;
                        const bool cv = c0 + b >= ctrail && c0 + b < m;
#pragma unroll
                        for (int aa = 0; aa < 4; ++aa) {
                            f[aa][b] = (cv && i0 + aa >= kb && i0 + aa < mp) ? __ldcg(colp + i0 + aa) : 0.0;
                            acc[aa][b] = 0.0;
                        }
                    }
                    for (int t = 0; t < kb; ++t) {
                        // i0 and ldp are multiples of 4: two conflict-free 16-byte shared loads per operand (rows
                        // beyond mp hold stale data whose results are never stored)
                        const double2 la = *reinterpret_cast<const double2 *>(P + i0 + t * ldp);
                        const double2 lb = *reinterpret_cast<const double2 *>(P + i0 + t * ldp + 2);
                        const double2 ua = *reinterpret_cast<const double2 *>(Uc + t * ldu + idx0);
                        const double2 ub = *reinterpret_cast<const double2 *>(Uc + t * ldu + idx0 + 2);
                        const double l[4] = {la.x, la.y, lb.x, lb.y}, uu[4] = {ua.x, ua.y, ub.x, ub.y};
#pragma unroll
                        for (int aa = 0; aa < 4; ++aa)
#pragma unroll
                            for (int b = 0; b < 4; ++b) acc[aa][b] = fma(l[aa], uu[b], acc[aa][b]);
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (c0 + b >= ctrail && c0 + b < m) {
                            double *colp = F + (size_t)(c0 + b) * m + k0;
#pragma unroll
                            for (int aa = 0; aa < 4; ++aa)
                                if (i0 + aa >= kb && i0 + aa < mp) __stcg(colp + i0 + aa, f[aa][b] - acc[aa][b]);
                        }
                    }
                }
            }
            MF_TICK(acc_trail);
        }
        cl.sync();      // publishes the updated trailing matrix (cluster) and the next diagonal block (this CTA)
        MF_TICK(acc_sync);
    }
    if (prof && tid == 0) {
        if (rank == 0) {
            long long v[6] = {acc_load, acc_diag, acc_trsm, acc_panel, acc_trail, acc_sync};
            for (int k = 0; k < 6; ++k) atomicAdd((unsigned long long *)prof + k, (unsigned long long)v[k]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Small fronts whose pivot part fits ONE CTA's shared memory (the leaves of the tree: 128 of the 255 fronts of the
// 32 x 32 mesh).  mf_factor_kernel walks such a front panel by panel through L2: per 16-column panel a load, a
// 4 x 4-tile pass over the WHOLE trailing matrix (16 L2 loads + 16 stores per 256 FMAs) and barriers - the leaves ran
// at ~10 % of an SM's fp64 rate.  Here the first np columns (Lp, m x np) and the first np rows (Up, np x (m - np))
// of the front live in shared memory, the right-looking LU only updates that L-shaped region, and the Schur
// complement - the bulk of the flops - is formed at the end in ONE pass, S -= L21 U12 with K = np (16 loads + 16 stores
// per 16 np FMAs, loads issued ahead of the K loop).  Static pivoting and the 16 x 16 diagonal-block routine are those
// of mf_factor_kernel; the summation order differs (one subtraction of the accumulated product), so the factors agree
// to rounding, not bitwise.  Children (levels above the leaves) are extend-added first, one child at a time.
//
// TL threads = TLW workers + one LOOK-AHEAD warp (ncu of the first version, profiles/prof_r2_leaf_v1.summary.txt: 40 % of
// all warp samples waited at the barriers behind the serial 16-step elimination of the diagonal block and behind a
// second, nearly empty round of the row / column substitutions).  Per panel:
//   phase 2  workers: rows below the block L = A U11^-1 and columns right of it U = L11^-1 A, one per thread, ONE round
//            (<= 2 (m - 16) tasks <= TLW); the rows of D come as 16-byte loads from aligned copies (s_Dr, s_Dc)
//   phase 3  workers: rank-16 update of the L-shaped region, 4 x 4 tiles, EXCEPT the next diagonal block;
//            look-ahead warp: updates that block (same formula), factors it (diag16_factor) and publishes it
// The first diagonal block is factored by the look-ahead warp straight from global memory while the workers load.
constexpr int TL = 384, TLW = TL - 32;
constexpr size_t kLeafSmemMax = 220 * 1024;   // dynamic shared memory the kernel may ask for (+ 6.6 KB static)
__host__ __device__ inline int leaf_ld(int rows) { return (rows + 3) & ~3; }
__host__ __device__ inline size_t leaf_smem_doubles(int m, int np) {
    return (size_t)leaf_ld(m) * np + (size_t)leaf_ld(m - np) * np + (size_t)NB * leaf_ld(m);
}

// look-ahead warp: factor the block held one row per lane, publish D (padded + aligned row / column copies) and 1/diag
__device__ __forceinline__ void leaf_publish_block(double (&r)[NB], int lane, int kb, double (*s_D)[NB + 1],
                                                   double (*s_Dr)[NB], double (*s_Dc)[NB], double (*s_prow)[NB],
                                                   double *s_rd, int *info, int front) {
    diag16_factor(r, lane, kb, s_prow, s_rd, info, front);
    if (lane < NB) {
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) {
            s_D[lane][jj] = r[jj];
            s_Dr[lane][jj] = r[jj];            // row `lane` of the block (U on / above the diagonal)
            s_Dc[jj][lane] = r[jj];            // s_Dc[t][tt] = D[tt][t]: column t of the block (L below the diagonal)
        }
    }
}

__global__ void __launch_bounds__(TL, 1)
mf_leaf_factor_kernel(MFDev d, const int *__restrict__ nodes, int *info) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double s_D[NB][NB + 1];            // factored diagonal block: L below, U on / above the diagonal
    __shared__ __align__(16) double s_Dr[NB][NB]; // the same, rows 16-byte aligned
    __shared__ __align__(16) double s_Dc[NB][NB]; // its transpose
    __shared__ double s_rd[NB];                   // reciprocals of its diagonal
    __shared__ __align__(16) double s_prow[2][NB];
    const int s = nodes[blockIdx.x];
    const int m = d.m[s], np = d.np[s], tid = threadIdx.x, lane = tid & 31;
    const bool worker = tid < TLW;
    double *F = d.F + d.front_ptr[s];
    // ---- extend-add, one child at a time (single writer per entry, see mf_factor_kernel)
    for (int ci = d.child_ptr[s]; ci < d.child_ptr[s + 1]; ++ci) {
        const int c = d.child[ci], mc = d.m[c], npc = d.np[c], nuc = mc - npc;
        const double *Fc = d.F + d.front_ptr[c];
        const int *rel = d.rel + d.rel_ptr[c];
        const int total = nuc * nuc;
        constexpr int EU = 8;
        for (int e0 = tid; e0 < total; e0 += EU * TL) {
            double *p[EU];
            double v[EU];
#pragma unroll
            for (int u = 0; u < EU; ++u) {
                const int e = e0 + u * TL;
                p[u] = nullptr;
                v[u] = 0.0;
                if (e < total) {
                    const int j = e / nuc, i = e - j * nuc;
                    p[u] = F + (size_t)__ldg(rel + j) * m + __ldg(rel + i);
                    v[u] = __ldcg(Fc + npc + i + (size_t)(npc + j) * mc);
                }
            }
#pragma unroll
            for (int u = 0; u < EU; ++u)
                if (p[u]) v[u] += __ldcg(p[u]);
#pragma unroll
            for (int u = 0; u < EU; ++u)
                if (p[u]) __stcg(p[u], v[u]);
        }
        __syncthreads();
    }
    if (np == 0) return;
    const int nu = m - np;
    const int ldL = leaf_ld(m), ldU = leaf_ld(nu), ldS = leaf_ld(m);
    double *Lp = sm;                              // columns 0 .. np-1:            Lp[i + c * ldL]
    double *Up = Lp + (size_t)ldL * np;           // rows 0 .. np-1 of columns np..: Up[t * ldU + (c - np)]
    double *Us = Up + (size_t)ldU * np;           // U rows of the current panel:   Us[t * ldS + (c - k1)]
    if (worker) {
        // ---- load the L-shaped region, eight independent loads in flight per thread
        constexpr int LU8 = 8;
        const int nL = m * np, nU = np * nu;
        for (int e0 = tid; e0 < nL; e0 += LU8 * TLW) {
            double v[LU8];
#pragma unroll
            for (int u = 0; u < LU8; ++u) v[u] = (e0 + u * TLW < nL) ? __ldcg(F + e0 + u * TLW) : 0.0;
#pragma unroll
            for (int u = 0; u < LU8; ++u) {
                const int e = e0 + u * TLW;
                if (e < nL) {
                    const int c = e / m, i = e - c * m;
                    Lp[i + c * ldL] = v[u];
                }
            }
        }
        for (int e0 = tid; e0 < nU; e0 += LU8 * TLW) {
            double v[LU8];
            int cc[LU8], tt[LU8];
#pragma unroll
            for (int u = 0; u < LU8; ++u) {
                const int e = e0 + u * TLW;
                cc[u] = e / np;
                tt[u] = e - cc[u] * np;
                v[u] = (e < nU) ? __ldcg(F + tt[u] + (size_t)(np + cc[u]) * m) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < LU8; ++u)
                if (e0 + u * TLW < nU) Up[tt[u] * ldU + cc[u]] = v[u];
        }
    } else {
        // ---- look-ahead warp: first diagonal block straight from the front
        const int kb = min(NB, np);
        double r[NB];
#pragma unroll
        for (int jj = 0; jj < NB; ++jj)
            r[jj] = (lane < kb && jj < kb) ? __ldcg(F + lane + (size_t)jj * m) : ((jj == lane) ? 1.0 : 0.0);
        leaf_publish_block(r, lane, kb, s_D, s_Dr, s_Dc, s_prow, s_rd, info, s);
        __syncwarp();
    }
    __syncthreads();
    if (!worker && lane < min(NB, np)) {          // (the workers have stored the unfactored block: replace it)
        const int kb = min(NB, np);
#pragma unroll
        for (int jj = 0; jj < NB; ++jj)
            if (jj < kb) Lp[lane + jj * ldL] = s_D[lane][jj];
    }
    for (int k0 = 0; k0 < np; k0 += NB) {
        const int kb = min(NB, np - k0), k1 = k0 + kb;
        // (2) rows below the block: L = A U11^-1; columns right of it: U = L11^-1 A (one row / column per thread).
        // Rows / columns >= kb of the block are identity padding, so all 16 steps run unconditionally.
        const int nrest = m - k1;
        if (worker) {
            for (int task = tid; task < 2 * nrest; task += TLW) {
                if (task < nrest) {
                    const int i = k1 + task;
                    double a[NB];
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) a[jj] = jj < kb ? Lp[i + (k0 + jj) * ldL] : 0.0;
#pragma unroll
                    for (int t = 0; t < NB; ++t) {
                        // (keeps the compiler from hoisting all 120 entries of D into registers - it spilled 48 of them)
                        if ((t & 1) == 0) asm volatile("" ::: "memory");
                        const double l = a[t] * s_rd[t];
                        a[t] = l;
                        const double2 *dr = reinterpret_cast<const double2 *>(s_Dr[t]);
#pragma unroll
                        for (int j2 = 0; j2 < NB / 2; ++j2) {
                            if (2 * j2 + 1 > t) {
                                const double2 dd = dr[j2];
                                if (2 * j2 > t) a[2 * j2] = fma(-l, dd.x, a[2 * j2]);
                                a[2 * j2 + 1] = fma(-l, dd.y, a[2 * j2 + 1]);
                            }
                        }
                    }
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj)
                        if (jj < kb) Lp[i + (k0 + jj) * ldL] = a[jj];
                } else {
                    const int c = k1 + (task - nrest);
                    double *col = c < np ? Lp + k0 + (size_t)c * ldL : Up + (size_t)k0 * ldU + (c - np);
                    const int st = c < np ? 1 : ldU;
                    double u[NB];
#pragma unroll
                    for (int t = 0; t < NB; ++t) u[t] = t < kb ? col[t * st] : 0.0;
#pragma unroll
                    for (int t = 0; t < NB; ++t) {
                        if ((t & 1) == 0) asm volatile("" ::: "memory");
                        const double ut = u[t];
                        const double2 *dc = reinterpret_cast<const double2 *>(s_Dc[t]);
#pragma unroll
                        for (int j2 = 0; j2 < NB / 2; ++j2) {
                            if (2 * j2 + 1 > t) {
                                const double2 dd = dc[j2];
                                if (2 * j2 > t) u[2 * j2] = fma(-dd.x, ut, u[2 * j2]);
                                u[2 * j2 + 1] = fma(-dd.y, ut, u[2 * j2 + 1]);
                            }
                        }
                    }
#pragma unroll
                    for (int t = 0; t < NB; ++t) {
                        if (t < kb) {
                            col[t * st] = u[t];
                            Us[t * ldS + (c - k1)] = u[t];
                        }
                    }
                }
            }
        }
        __syncthreads();
        // (3) rank-16 update of the L-shaped region only: rows [k1, m) x columns [k1, np) (in Lp) and rows [k1, np) x
        // columns [np, m) (in Up); 4 x 4 tiles, a tile that straddles column / row np is predicated per element.  The
        // next diagonal block [k1, k1 + kb2)^2 belongs to the look-ahead warp.
        if (k1 < np) {
            const int kb2 = min(NB, np - k1), kend = k1 + kb2;
            if (worker) {
                const int tr = (m - k1 + 3) >> 2, tcA = (np - k1 + 3) >> 2, trB = tcA;
                const int ntA = tr * tcA, ntB = trB * (tr - tcA);
                for (int tile = tid; tile < ntA + ntB; tile += TLW) {
                    int ti, tj;
                    if (tile < ntA) {
                        ti = tile % tr;
                        tj = tile / tr;
                    } else {
                        const int q = tile - ntA;
                        ti = q % trB;
                        tj = tcA + q / trB;
                    }
                    const int i0 = k1 + 4 * ti, c0 = k1 + 4 * tj;
                    if (i0 + 3 < kend && c0 + 3 < kend) continue;          // entirely inside the look-ahead block
                    double acc[4][4];
#pragma unroll
                    for (int aa = 0; aa < 4; ++aa)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[aa][b] = 0.0;
#pragma unroll 4
                    for (int t = 0; t < NB; ++t) {
                        const double2 la = *reinterpret_cast<const double2 *>(Lp + i0 + (k0 + t) * ldL);
                        const double2 lb = *reinterpret_cast<const double2 *>(Lp + i0 + (k0 + t) * ldL + 2);
                        const double2 ua = *reinterpret_cast<const double2 *>(Us + t * ldS + (c0 - k1));
                        const double2 ub = *reinterpret_cast<const double2 *>(Us + t * ldS + (c0 - k1) + 2);
                        const double l[4] = {la.x, la.y, lb.x, lb.y}, uu[4] = {ua.x, ua.y, ub.x, ub.y};
#pragma unroll
                        for (int aa = 0; aa < 4; ++aa)
#pragma unroll
                            for (int b = 0; b < 4; ++b) acc[aa][b] = fma(l[aa], uu[b], acc[aa][b]);
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int c = c0 + b;
#pragma unroll
                        for (int aa = 0; aa < 4; ++aa) {
                            const int i = i0 + aa;
                            if (i < m && c < m && !(i >= np && c >= np) && !(i < kend && c < kend)) {
                                double *e = c < np ? Lp + i + (size_t)c * ldL : Up + (size_t)i * ldU + (c - np);
                                *e -= acc[aa][b];
                            }
                        }
                    }
                }
            } else {
                // look-ahead warp: row k1 + lane of the next diagonal block, updated by this panel and factored
                double r[NB];
                const int i = k1 + lane;
                const bool live = lane < kb2;
#pragma unroll
                for (int jj = 0; jj < NB; ++jj) r[jj] = 0.0;
                if (live) {
                    for (int t = 0; t < NB; ++t) {
                        const double l = Lp[i + (k0 + t) * ldL];
                        const double2 *ur = reinterpret_cast<const double2 *>(Us + t * ldS);
#pragma unroll
                        for (int j2 = 0; j2 < NB / 2; ++j2) {
                            const double2 uu = ur[j2];
                            r[2 * j2] = fma(l, uu.x, r[2 * j2]);
                            r[2 * j2 + 1] = fma(l, uu.y, r[2 * j2 + 1]);
                        }
                    }
                }
#pragma unroll
                for (int jj = 0; jj < NB; ++jj)
                    r[jj] = (live && jj < kb2) ? Lp[i + (k1 + jj) * ldL] - r[jj] : ((jj == lane) ? 1.0 : 0.0);
                leaf_publish_block(r, lane, kb2, s_D, s_Dr, s_Dc, s_prow, s_rd, info, s);
                if (live) {
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj)
                        if (jj < kb2) Lp[i + (k1 + jj) * ldL] = r[jj];
                }
            }
            __syncthreads();
        }
    }
    // (4) Schur complement in one pass: F[i, c] -= sum_t L[i, t] U[t, c], i, c >= np.  A tile takes the row PAIRS ti and
    // ti + H (pairs counted from the 4-aligned row rb): consecutive lanes then read consecutive 16-byte pieces of a
    // column of L - conflict-free, where four consecutive rows per lane cost two wavefronts per quarter warp (ncu of the
    // second version: the phase is bound by shared-memory wavefronts, 40 % of them excess).
    {
        const int rb = np & ~3;
        const int npair = (m - rb + 1) >> 1, H = (npair + 1) >> 1, tcS = (nu + 3) >> 2;
        for (int tile = tid; tile < H * tcS; tile += TL) {
            const int ti = tile % H, tj = tile / H;
            const int rA = rb + 2 * ti, rB = rb + 2 * (ti + H), cc0 = 4 * tj;
            const int ia[4] = {rA, rA + 1, rB, rB + 1};
            double f[4][4], acc[4][4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double *colp = F + (size_t)(np + cc0 + b) * m;
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) {
                    const int i = ia[aa];
                    f[aa][b] = (i >= np && i < m && cc0 + b < nu) ? __ldcg(colp + i) : 0.0;
                    acc[aa][b] = 0.0;
                }
            }
            const double *La = Lp + rA, *Lb = Lp + rB, *Uq = Up + cc0;
#pragma unroll 4
            for (int t = 0; t < np; ++t) {
                const double2 la = *reinterpret_cast<const double2 *>(La);
                const double2 lb = *reinterpret_cast<const double2 *>(Lb);
                const double2 ua = *reinterpret_cast<const double2 *>(Uq);
                const double2 ub = *reinterpret_cast<const double2 *>(Uq + 2);
                La += ldL;
                Lb += ldL;
                Uq += ldU;
                const double l[4] = {la.x, la.y, lb.x, lb.y}, uu[4] = {ua.x, ua.y, ub.x, ub.y};
#pragma unroll
                for (int aa = 0; aa < 4; ++aa)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[aa][b] = fma(l[aa], uu[b], acc[aa][b]);
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                double *colp = F + (size_t)(np + cc0 + b) * m;
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) {
                    const int i = ia[aa];
                    if (i >= np && i < m && cc0 + b < nu) __stcg(colp + i, f[aa][b] - acc[aa][b]);
                }
            }
        }
    }
    // (5) factors back to the front
    for (int e = tid; e < m * np; e += TL) {
        const int c = e / m, i = e - c * m;
        __stcg(F + e, Lp[i + c * ldL]);
    }
    for (int e = tid; e < np * nu; e += TL) {
        const int c = e / np, t = e - c * np;
        __stcg(F + t + (size_t)(np + c) * m, Up[t * ldU + c]);
    }
}

// Triangular solves, one CTA per front and one launch per level.  Inside a front the pivot block is processed in
// 16-row blocks: the block itself is a 16x16 mat-vec with the inverse stored by the factorisation (warp 0), the rows
// outside the block are updated by all threads with the 16 block values; the matrix entries of the NEXT block are
// loaded into registers while the current one is being applied.
constexpr int TS = 512;    // threads per CTA in the solve kernels

// element (i, c) of the triangular factor a sweep works with: the stored factor, or its transpose (TR) for solves
// with A^T = U^T L^T (the adjoint system reuses the factors of the last Newton matrix that way)
#define MF_E(i, c) (TR ? __ldg(F + (c) + (size_t)(i) * m) : __ldg(F + (i) + (size_t)(c) * m))
#define MF_DI(row, col) (TR ? (col) * NB + (row) : (row) * NB + (col))



// The factors are read-only during a solve, so they may live in L1: the whole panel a sweep will walk in dependent
// 16-column steps is requested up front (one prefetch per 128-byte line), and the per-block loads then hit L1 instead
// of paying an L2 round trip on the serial path.  cols: the first np columns (L panel), else the first np rows (U panel).
__device__ __forceinline__ void mf_prefetch_l1(const double *F, int m, int np, bool cols, int tid, int nthreads) {
    if (cols) {
        const size_t n = (size_t)np * m;
        for (size_t off = (size_t)tid * 16; off < n; off += (size_t)nthreads * 16)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(F + off));
    } else {
        const int nch = (np + 15) >> 4;
        for (int e = tid; e < m * nch; e += nthreads) {
            const int j = e / nch, r = (e - j * nch) << 4;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(F + (size_t)j * m + r));
        }
    }
}

// forward: y_P = L11^{-1} b_P,  b_U -= L21 y_P      (RS = rows per thread: 1 for fronts <= 512, 2 up to 1024)
// TR: the same sweep with U^T in place of L:  y_P = U11^{-T} b_P,  b_U -= U12^T y_P
// NR right-hand sides at once (vector r starts at x + r * ldx): the factor entries are loaded once for all of them.
template <int RS, bool TR, int NR>
__global__ void __launch_bounds__(TS)
mf_forward_kernel(MFDev d, const int *__restrict__ nodes, double *__restrict__ x, int ldx) {
    extern __shared__ double y[];          // NR vectors of length m
    __shared__ double yb[NR][NB];
    const int s = nodes[blockIdx.x];
    const int m = d.m[s], np = d.np[s], tid = threadIdx.x;
    if (np == 0) return;
    const double *F = d.F + d.front_ptr[s];
    const int *I = d.idx + d.idx_ptr[s];
    const double *Dinv = d.dinv + (size_t)d.dinv_ptr[s] * (2 * NB * NB) + (TR ? NB * NB : 0);
    mf_prefetch_l1(F, m, np, !TR, tid, TS);
    for (int off = tid * 16; off < ((np + NB - 1) / NB) * 2 * NB * NB; off += TS * 16)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(d.dinv + (size_t)d.dinv_ptr[s] * (2 * NB * NB) + off));
    for (int r = 0; r < NR; ++r)
        for (int k = tid; k < m; k += TS) y[r * m + k] = k < np ? x[(size_t)r * ldx + I[k]] : 0.0;
    // warps whose rows all lie beyond the front skip the per-block row work (they still meet the barriers): with 512
    // threads and fronts of order ~100-380 that is most warps, and their predicated address arithmetic was a fifth
    // of the issue slots of a block step
    const bool live = (tid & ~31) < m;
    double lc[RS][NB], ln[RS][NB];
#pragma unroll
    for (int q = 0; q < RS; ++q) {
        const int i = tid + q * TS;
#pragma unroll
        for (int t = 0; t < NB; ++t) lc[q][t] = (live && i < m && i >= min(NB, np) && t < np) ? MF_E(i, t) : 0.0;
    }
    double dl[NB];
#pragma unroll
    for (int t = 0; t < NB; ++t) dl[t] = (tid < NB) ? __ldg(Dinv + MF_DI(tid, t)) : 0.0;
    __syncthreads();
    const int nblk = (np + NB - 1) / NB;
    for (int b = 0; b < nblk; ++b) {
        const int k0 = b * NB, kb = min(NB, np - k0), k1 = k0 + NB;
        // prefetch the next block's columns / inverse (rows below the next block; its last block may be partial)
        if (live && b + 1 < nblk) {
            const int rnext = k1 + min(NB, np - k1);
#pragma unroll
            for (int q = 0; q < RS; ++q) {
                const int i = tid + q * TS;
#pragma unroll
                for (int t = 0; t < NB; ++t)
                    ln[q][t] = (i < m && i >= rnext && k1 + t < np) ? MF_E(i, k1 + t) : 0.0;
            }
        }
        if (tid < NB) {
            double v[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                v[r] = 0.0;
#pragma unroll
                for (int t = 0; t < NB; ++t) v[r] = fma(dl[t], (t < kb) ? y[r * m + k0 + t] : 0.0, v[r]);
            }
            __syncwarp(0xffffu);
            if (tid < kb) {
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    yb[r][tid] = v[r];
                    y[r * m + k0 + tid] = v[r];
                }
            }
            if (b + 1 < nblk) {
#pragma unroll
                for (int t = 0; t < NB; ++t) dl[t] = __ldg(Dinv + (size_t)(b + 1) * (2 * NB * NB) + MF_DI(tid, t));
            }
        }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int q = 0; q < RS; ++q) {
                const int i = tid + q * TS;
                if (i < m && i >= k0 + kb) {
#pragma unroll
                    for (int r = 0; r < NR; ++r) {
                        double acc = 0.0;
#pragma unroll
                        for (int t = 0; t < NB; ++t) acc = fma(lc[q][t], (t < kb) ? yb[r][t] : 0.0, acc);
                        y[r * m + i] -= acc;
                    }
                }
            }
        }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int q = 0; q < RS; ++q)
#pragma unroll
                for (int t = 0; t < NB; ++t) lc[q][t] = ln[q][t];
        }
    }
    for (int r = 0; r < NR; ++r) {
        for (int k = tid; k < np; k += TS) x[(size_t)r * ldx + I[k]] = y[r * m + k];
        for (int i = np + tid; i < m; i += TS) atomicAdd(x + (size_t)r * ldx + I[i], y[r * m + i]);
    }
}

// backward: x_P = U11^{-1} (y_P - U12 x_U);   TR: x_P = L11^{-T} (y_P - L21^T x_U)
template <int RS, bool TR, int NR>
__global__ void __launch_bounds__(TS)
mf_backward_kernel(MFDev d, const int *__restrict__ nodes, double *__restrict__ x, int ldx) {
    extern __shared__ double y[];
    __shared__ double yb[NR][NB];
    const int s = nodes[blockIdx.x];
    const int m = d.m[s], np = d.np[s], tid = threadIdx.x;
    if (np == 0) return;
    const double *F = d.F + d.front_ptr[s];
    const int *I = d.idx + d.idx_ptr[s];
    const double *Dinv = d.dinv + (size_t)d.dinv_ptr[s] * (2 * NB * NB) + (TR ? 0 : NB * NB);
    mf_prefetch_l1(F, m, np, TR, tid, TS);
    for (int off = tid * 16; off < ((np + NB - 1) / NB) * 2 * NB * NB; off += TS * 16)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(d.dinv + (size_t)d.dinv_ptr[s] * (2 * NB * NB) + off));
    for (int r = 0; r < NR; ++r)
        for (int k = tid; k < m; k += TS) y[r * m + k] = x[(size_t)r * ldx + I[k]];
    const int nblk = (np + NB - 1) / NB;
    double uc[RS][NB], un[RS][NB];
    {
        const int k0 = (nblk - 1) * NB;
#pragma unroll
        for (int q = 0; q < RS; ++q) {
            const int i = tid + q * TS;
#pragma unroll
            for (int t = 0; t < NB; ++t) uc[q][t] = (i < k0 && k0 + t < np) ? MF_E(i, k0 + t) : 0.0;
        }
    }
    double du[NB];
#pragma unroll
    for (int t = 0; t < NB; ++t) du[t] = (tid < NB) ? __ldg(Dinv + (size_t)(nblk - 1) * (2 * NB * NB) + MF_DI(tid, t)) : 0.0;
    __syncthreads();
    // y_P -= U12 x_U (TR: L21^T x_U): all 16 warps share the (np x nu) mat-vec so that no thread walks a long
    // dependent chain of L2 loads.  Plain: warp w takes the columns j = np + w, np + w + 16, ... with its lanes over
    // the rows (coalesced), partial sums meet in shared memory.  Transposed: the factor is contiguous along j, so warp
    // w takes rows k = w, w + 16, ... with its lanes over j and a shuffle reduction.
    {
        const int lane = tid & 31, wid = tid >> 5, NW = TS / 32;
        if (!TR) {
            double *part = y + (size_t)NR * m;                 // NW x (NR x np) partial sums (dynamic smem sized by the host)
            for (int k0 = 0; k0 < np; k0 += 32) {
                const int k = k0 + lane;
                double a[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) a[r] = 0.0;
                if (k < np) {
                    for (int j = np + wid; j < m; j += NW) {
                        const double e = __ldg(F + k + (size_t)j * m);
#pragma unroll
                        for (int r = 0; r < NR; ++r) a[r] = fma(e, y[r * m + j], a[r]);
                    }
#pragma unroll
                    for (int r = 0; r < NR; ++r) part[((size_t)wid * NR + r) * np + k] = a[r];
                }
            }
            __syncthreads();
            for (int e = tid; e < NR * np; e += TS) {
                const int r = e / np, k = e % np;
                double sum = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) sum += part[((size_t)w * NR + r) * np + k];
                y[r * m + k] -= sum;
            }
        } else {
            for (int k = wid; k < np; k += NW) {
                double a[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) a[r] = 0.0;
                for (int j = np + lane; j < m; j += 32) {
                    const double e = __ldg(F + j + (size_t)k * m);
#pragma unroll
                    for (int r = 0; r < NR; ++r) a[r] = fma(e, y[r * m + j], a[r]);
                }
#pragma unroll
                for (int r = 0; r < NR; ++r) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) a[r] += __shfl_xor_sync(0xffffffffu, a[r], o);
                    if (lane == 0) y[r * m + k] -= a[r];
                }
            }
        }
    }
    __syncthreads();
    for (int b = nblk - 1; b >= 0; --b) {
        const int k0 = b * NB, kb = min(NB, np - k0);
        const bool live = (tid & ~31) < k0;       // rows above the block only: dead warps just meet the barriers
        if (live && b > 0) {
            const int kp = k0 - NB;
#pragma unroll
            for (int q = 0; q < RS; ++q) {
                const int i = tid + q * TS;
#pragma unroll
                for (int t = 0; t < NB; ++t) un[q][t] = (i < kp) ? MF_E(i, kp + t) : 0.0;
            }
        }
        if (tid < NB) {
            double v[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                v[r] = 0.0;
#pragma unroll
                for (int t = 0; t < NB; ++t) v[r] = fma(du[t], (t < kb) ? y[r * m + k0 + t] : 0.0, v[r]);
            }
            __syncwarp(0xffffu);
            if (tid < kb) {
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    yb[r][tid] = v[r];
                    y[r * m + k0 + tid] = v[r];
                }
            }
            if (b > 0) {
#pragma unroll
                for (int t = 0; t < NB; ++t) du[t] = __ldg(Dinv + (size_t)(b - 1) * (2 * NB * NB) + MF_DI(tid, t));
            }
        }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int q = 0; q < RS; ++q) {
                const int i = tid + q * TS;
                if (i < k0) {
#pragma unroll
                    for (int r = 0; r < NR; ++r) {
                        double acc = 0.0;
#pragma unroll
                        for (int t = 0; t < NB; ++t) acc = fma(uc[q][t], (t < kb) ? yb[r][t] : 0.0, acc);
                        y[r * m + i] -= acc;
                    }
                }
            }
        }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int q = 0; q < RS; ++q)
#pragma unroll
                for (int t = 0; t < NB; ++t) uc[q][t] = un[q][t];
        }
    }
    for (int r = 0; r < NR; ++r)
        for (int k = tid; k < np; k += TS) x[(size_t)r * ldx + I[k]] = y[r * m + k];
}


// ---------------------------------------------------------------------------------------------------------------
// Triangular solves of the small fronts in 64-ROW blocks.  The solve of a front is a dependent chain over its pivot
// blocks (block b needs the update of blocks < b); with 16-row blocks the root front of the 32 x 32 mesh walks 12 such
// steps, each costing two block barriers and an L2 round trip (~3 k cycles).  Here the explicit inverses of the
// 64 x 64 diagonal blocks of L and U are formed once per factorisation (mf_dinv64_kernel), so a front has a quarter
// of the steps, and each step is two fully parallel mat-vecs:
//   (a) y_blk = D^-1 y_blk        256 threads, four per row (interleaved columns: coalesced), partial sums via smem
//   (b) y_i  -= E(i, blk) y_blk   one thread per remaining row, its 64 factor entries requested up front - before
//                                 (a) - so that their L2 latency overlaps the pivot-block product.
constexpr int SB = 64;
// Programmatic dependent launch between the level kernels of one sweep (OCP_MF_PDL, default on): a level kernel lets
// the next level's grid start at once (launch_dependents) and the next level does everything that only touches the
// READ-ONLY data of a solve - front metadata, index list, inverse blocks, its first factor entries - before it waits
// (griddepcontrol.wait) for the previous level's right-hand-side updates.  Both instructions are no-ops in a launch
// without the attribute.  Every exit path waits, so a level never completes before its predecessor has.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
constexpr int TS6 = 384;     // threads of the 64-row-block solve kernels: 168 registers each keep the 64 entries of a row resident

// One CTA of 64 threads per (64-pivot block, L or U): thread c forms column c of the inverse by substitution with the
// block in shared memory and all of its 64 unknowns in registers (outer-product form: the 64 dependent steps each
// update independent registers).  Rows / columns beyond a partial last block are identity padding.
__global__ void __launch_bounds__(SB)
mf_dinv64_kernel(MFDev d, const int *__restrict__ block_node, int first, int nblocks) {
    __shared__ double B[SB][SB + 1];
    const int p = first + (blockIdx.x >> 1), which = blockIdx.x & 1;      // which = 0: L11^-1, 1: U11^-1
    if (p >= first + nblocks) return;
    const int s = block_node[p];
    const int m = d.m[s], np = d.np[s];
    const int k0 = (p - d.dinv64_ptr[s]) * SB, kb = min(SB, np - k0);
    const double *F = d.F + d.front_ptr[s];
    const int c = threadIdx.x;
    for (int j = 0; j < SB; ++j)                                 // column j of the block, rows over the threads
        B[c][j] = (c < kb && j < kb) ? __ldcg(F + (k0 + c) + (size_t)(k0 + j) * m) : ((c == j) ? 1.0 : 0.0);
    __syncthreads();
    double x[SB];
#pragma unroll
    for (int i = 0; i < SB; ++i) x[i] = (i == c) ? 1.0 : 0.0;
    if (which == 0) {
        // L x = e_c, L unit lower: for k ascending, x_i -= L[i][k] x_k (i > k)
#pragma unroll
        for (int k = 0; k < SB; ++k) {
            const double xk = x[k];
#pragma unroll
            for (int i = 0; i < SB; ++i)
                if (i > k) x[i] = fma(-B[i][k], xk, x[i]);
        }
    } else {
        // U x = e_c, U upper with its diagonal: for k descending, x_k /= U[k][k], x_i -= U[i][k] x_k (i < k)
#pragma unroll
        for (int k = SB - 1; k >= 0; --k) {
            const double xk = x[k] * fast_rcp(B[k][k]);
            x[k] = xk;
#pragma unroll
            for (int i = 0; i < SB; ++i)
                if (i < k) x[i] = fma(-B[i][k], xk, x[i]);
        }
    }
    double *dst = d.dinv64 + (size_t)p * (2 * SB * SB) + (size_t)which * (SB * SB);
#pragma unroll
    for (int i = 0; i < SB; ++i) dst[i * SB + c] = x[i];         // row-major X[i][c]: coalesced over the threads
}

// (a) of a block step: yb[r][.] = sum_t Dm(., t) y[r][k0 + t] with Dm = the stored inverse (DT = false) or its
// transpose (DT = true); 256 threads.  Partial sums over four interleaved column classes meet in shared memory.
template <bool DT, int NR>
__device__ __forceinline__ void block_inverse_apply(const double *__restrict__ Dblk, const double *y, int m, int k0,
                                                    int kb, int tid, double (*part)[NR][SB]) {
    if (tid < 4 * SB) {
        // DT = false: lanes run over the column class q (four lanes share a row: 32 contiguous bytes per row);
        // DT = true: lanes run over the rows (the transposed entry X[t][r] is contiguous in r)
        const int r = DT ? (tid & (SB - 1)) : (tid >> 2), q = DT ? (tid >> 6) : (tid & 3);
        double a[NR];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) a[rr] = 0.0;
#pragma unroll
        for (int s4 = 0; s4 < SB / 4; ++s4) {
            const int t = q + 4 * s4;
            const double e = __ldg(Dblk + (DT ? t * SB + r : r * SB + t));
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) a[rr] = fma(e, (t < kb) ? y[rr * m + k0 + t] : 0.0, a[rr]);
        }
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) part[q][rr][r] = a[rr];
    }
}

// forward: y_P = L11^-1 b_P, b_U -= L21 y_P;  TR: the same sweep with U^T in place of L
template <bool TR, int NR>
__global__ void __launch_bounds__(TS6, 1)
mf_forward64_kernel(MFDev d, const int *__restrict__ nodes, double *__restrict__ x, int ldx) {
    extern __shared__ double y[];          // NR vectors of length m
    __shared__ double yb[NR][SB];
    __shared__ double part[4][NR][SB];
    pdl_launch_dependents();
    const int s = nodes[blockIdx.x];
    const int m = d.m[s], np = d.np[s], tid = threadIdx.x;
    if (np == 0) {
        pdl_wait();
        return;
    }
    const double *F = d.F + d.front_ptr[s];
    const int *I = d.idx + d.idx_ptr[s];
    // plain: L11^-1; transposed: (U11^-1)^T
    const double *Dinv = d.dinv64 + (size_t)d.dinv64_ptr[s] * (2 * SB * SB) + (TR ? SB * SB : 0);
    const int nblk = (np + SB - 1) / SB;
    for (int off = tid * 16; off < nblk * 2 * SB * SB; off += TS6 * 16)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(d.dinv64 + (size_t)d.dinv64_ptr[s] * (2 * SB * SB) + off));
    // read-only prologue (ahead of the previous level's completion under PDL): the thread's entries of the index list
    // and the factor entries of its first remaining row
    const int ik0 = tid < m ? I[tid] : 0, ik1 = tid + TS6 < m ? I[tid + TS6] : 0;
    double e[SB];
    {
        const int kb = min(SB, np), i0 = kb + tid;
        if (i0 < m) {
#pragma unroll
            for (int t = 0; t < SB; ++t) e[t] = (t < kb) ? MF_E(i0, t) : 0.0;
        }
    }
    pdl_wait();
    for (int r = 0; r < NR; ++r)
        for (int k = tid, j = 0; k < m; k += TS6, ++j) {
            const int ii = j == 0 ? ik0 : (j == 1 ? ik1 : I[k]);
            y[r * m + k] = k < np ? __ldcg(x + (size_t)r * ldx + ii) : 0.0;
        }
    __syncthreads();
    for (int b = 0; b < nblk; ++b) {
        const int k0 = b * SB, kb = min(SB, np - k0);
        // factor entries of this thread's first remaining row: requested before the pivot-block product
        const int i0 = k0 + kb + tid;
        if (b > 0 && i0 < m) {
#pragma unroll
            for (int t = 0; t < SB; ++t) e[t] = (t < kb) ? MF_E(i0, k0 + t) : 0.0;
        }
        block_inverse_apply<TR, NR>(Dinv + (size_t)b * (2 * SB * SB), y, m, k0, kb, tid, part);
        __syncthreads();
        if (tid < SB * NR) {
            const int rr = tid / SB, r = tid % SB;
            const double v = (part[0][rr][r] + part[1][rr][r]) + (part[2][rr][r] + part[3][rr][r]);
            yb[rr][r] = v;
            if (r < kb) y[rr * m + k0 + r] = v;
        }
        __syncthreads();
        for (int i = i0; i < m; i += TS6) {
            if (i != i0) {
#pragma unroll
                for (int t = 0; t < SB; ++t) e[t] = (t < kb) ? MF_E(i, k0 + t) : 0.0;
            }
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;          // four chains of 16 instead of one of 64
#pragma unroll
                for (int t = 0; t < SB; t += 4) {
                    a0 = fma(e[t], yb[rr][t], a0);
                    a1 = fma(e[t + 1], yb[rr][t + 1], a1);
                    a2 = fma(e[t + 2], yb[rr][t + 2], a2);
                    a3 = fma(e[t + 3], yb[rr][t + 3], a3);
                }
                y[rr * m + i] -= (a0 + a1) + (a2 + a3);
            }
        }
        __syncthreads();
    }
    for (int r = 0; r < NR; ++r)
        for (int k = tid, j = 0; k < m; k += TS6, ++j) {
            const int ii = j == 0 ? ik0 : (j == 1 ? ik1 : I[k]);
            if (k < np) x[(size_t)r * ldx + ii] = y[r * m + k];
            else atomicAdd(x + (size_t)r * ldx + ii, y[r * m + k]);
        }
}

// backward: x_P = U11^-1 (y_P - U12 x_U);   TR: x_P = L11^-T (y_P - L21^T x_U)
template <bool TR, int NR>
__global__ void __launch_bounds__(TS6, 1)
mf_backward64_kernel(MFDev d, const int *__restrict__ nodes, double *__restrict__ x, int ldx) {
    extern __shared__ double y[];          // NR x m, then (plain) the per-warp partial sums of the U12 mat-vec
    __shared__ double yb[NR][SB];
    __shared__ double part4[4][NR][SB];
    pdl_launch_dependents();
    const int s = nodes[blockIdx.x];
    const int m = d.m[s], np = d.np[s], tid = threadIdx.x;
    if (np == 0) {
        pdl_wait();
        return;
    }
    const double *F = d.F + d.front_ptr[s];
    const int *I = d.idx + d.idx_ptr[s];
    // plain: U11^-1; transposed: (L11^-1)^T
    const double *Dinv = d.dinv64 + (size_t)d.dinv64_ptr[s] * (2 * SB * SB) + (TR ? 0 : SB * SB);
    const int nblk = (np + SB - 1) / SB;
    for (int off = tid * 16; off < nblk * 2 * SB * SB; off += TS6 * 16)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(d.dinv64 + (size_t)d.dinv64_ptr[s] * (2 * SB * SB) + off));
    // read-only prologue (ahead of the previous level's completion under PDL): the off-diagonal panel of the first
    // mat-vec into L1 when it is small (plain: columns np.. = U12 and below; transposed: columns ..np = L21 and above),
    // the thread's entries of the index list and the factor entries of its first row above the last pivot block
    if ((size_t)(m - np) * m <= 12288 && !TR) {
        for (size_t off = (size_t)tid * 16; off < (size_t)(m - np) * m; off += (size_t)TS6 * 16)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(F + (size_t)np * m + off));
    } else if ((size_t)np * m <= 12288 && TR && m > np) {
        for (size_t off = (size_t)tid * 16; off < (size_t)np * m; off += (size_t)TS6 * 16)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(F + off));
    }
    const int ik0 = tid < m ? I[tid] : 0, ik1 = tid + TS6 < m ? I[tid + TS6] : 0;
    double e[SB];
    {
        const int k0 = (nblk - 1) * SB, kb = min(SB, np - k0);
        if (tid < k0) {
#pragma unroll
            for (int t = 0; t < SB; ++t) e[t] = (t < kb) ? MF_E(tid, k0 + t) : 0.0;
        }
    }
    pdl_wait();
    for (int r = 0; r < NR; ++r)
        for (int k = tid, j = 0; k < m; k += TS6, ++j) {
            const int ii = j == 0 ? ik0 : (j == 1 ? ik1 : I[k]);
            y[r * m + k] = __ldcg(x + (size_t)r * ldx + ii);
        }
    __syncthreads();
    // y_P -= U12 x_U (TR: L21^T x_U), shared by all 16 warps (see mf_backward_kernel)
    {
        const int lane = tid & 31, wid = tid >> 5, NW = TS6 / 32;
        if (!TR) {
            double *part = y + (size_t)NR * m;                 // NW x (NR x np) partial sums
            for (int k0 = 0; k0 < np; k0 += 32) {
                const int k = k0 + lane;
                double a[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) a[r] = 0.0;
                if (k < np) {
                    for (int j = np + wid; j < m; j += NW) {
                        const double e = __ldg(F + k + (size_t)j * m);
#pragma unroll
                        for (int r = 0; r < NR; ++r) a[r] = fma(e, y[r * m + j], a[r]);
                    }
#pragma unroll
                    for (int r = 0; r < NR; ++r) part[((size_t)wid * NR + r) * np + k] = a[r];
                }
            }
            __syncthreads();
            for (int e = tid; e < NR * np; e += TS6) {
                const int r = e / np, k = e % np;
                double sum = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) sum += part[((size_t)w * NR + r) * np + k];
                y[r * m + k] -= sum;
            }
        } else {
            for (int k = wid; k < np; k += NW) {
                double a[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) a[r] = 0.0;
                for (int j = np + lane; j < m; j += 32) {
                    const double e = __ldg(F + j + (size_t)k * m);
#pragma unroll
                    for (int r = 0; r < NR; ++r) a[r] = fma(e, y[r * m + j], a[r]);
                }
#pragma unroll
                for (int r = 0; r < NR; ++r) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) a[r] += __shfl_xor_sync(0xffffffffu, a[r], o);
                    if (lane == 0) y[r * m + k] -= a[r];
                }
            }
        }
    }
    __syncthreads();
    for (int b = nblk - 1; b >= 0; --b) {
        const int k0 = b * SB, kb = min(SB, np - k0);
        // rows above the block: entries of this thread's first row, requested before the pivot-block product
        if (b < nblk - 1 && tid < k0) {
#pragma unroll
            for (int t = 0; t < SB; ++t) e[t] = (t < kb) ? MF_E(tid, k0 + t) : 0.0;
        }
        block_inverse_apply<TR, NR>(Dinv + (size_t)b * (2 * SB * SB), y, m, k0, kb, tid, part4);
        __syncthreads();
        if (tid < SB * NR) {
            const int rr = tid / SB, r = tid % SB;
            const double v = (part4[0][rr][r] + part4[1][rr][r]) + (part4[2][rr][r] + part4[3][rr][r]);
            yb[rr][r] = v;
            if (r < kb) y[rr * m + k0 + r] = v;
        }
        __syncthreads();
        for (int i = tid; i < k0; i += TS6) {
            if (i != tid) {
#pragma unroll
                for (int t = 0; t < SB; ++t) e[t] = (t < kb) ? MF_E(i, k0 + t) : 0.0;
            }
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int t = 0; t < SB; t += 4) {
                    a0 = fma(e[t], yb[rr][t], a0);
                    a1 = fma(e[t + 1], yb[rr][t + 1], a1);
                    a2 = fma(e[t + 2], yb[rr][t + 2], a2);
                    a3 = fma(e[t + 3], yb[rr][t + 3], a3);
                }
                y[rr * m + i] -= (a0 + a1) + (a2 + a3);
            }
        }
        __syncthreads();
    }
    for (int r = 0; r < NR; ++r)
        for (int k = tid, j = 0; k < np; k += TS6, ++j) x[(size_t)r * ldx + (j == 0 ? ik0 : (j == 1 ? ik1 : I[k]))] = y[r * m + k];
}


// ---------------------------------------------------------------------------------------------------------------
// Large fronts (m > kBigM: the top of the tree on refined meshes, cfg5).  A front no longer fits a cluster's shared
// memory, so a GROUP of G co-resident CTAs (cooperative launch; the SMs are shared out among the level's large fronts in
// proportion to their flops) works on it
// out of L2 / HBM with a right-looking blocked LU, panel width 32, static pivoting as above:
//   phase A  every CTA factors the 32 x 32 diagonal block redundantly (warp 0, pivot row via shared memory) and turns its share of the
//            rows below it into L = A U11^{-1} (one row per thread, in registers)
//   -- group barrier --
//   phase B  the CTA owning a column (4-column blocks, cyclic over the group) solves U12 for it and applies the
//            rank-32 update to it: U12 of the own columns sits in shared memory, every thread keeps the L rows of
//            two matrix rows in registers (2 x 32) and streams along them through the own columns, eight at a time -
//            C in, 512 DFMAs in 16 independent chains, C out - with the next eight columns' C values in flight.  Lanes are
//            consecutive rows, so every global access of a warp is 256 contiguous bytes and U is a broadcast.
//   -- group barrier --
// Column ownership is static, so a column is only ever read and written by its owner between barriers.
constexpr int BB = 32;        // panel width
constexpr int BIG_T = 256;    // threads per CTA (up to 255 registers each: two L rows of 32 live in registers)
constexpr int BIG_UC = 640;   // own columns whose U12 column blocks are resident in shared memory at a time
constexpr int kBigM = 768;    // fronts above this order take the group path
constexpr size_t kBigSmem = sizeof(double) * (BB * BIG_UC);

__device__ __forceinline__ void group_barrier(unsigned *cnt, unsigned target, int G) {
    __syncthreads();
    if (G > 1) {
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(cnt, 1u);
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
            } while (v < target);
            __threadfence();
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BIG_T, 1)
mf_big_factor_kernel(MFDev d, const int *__restrict__ nodes, const int4 *__restrict__ cta_map, unsigned *bar, int *info,
                     long long *prof) {
    extern __shared__ __align__(16) double sm[];
    double *Us = sm;                      // BIG_UC x BB: Us[c * BB + t], U12 of own column slot c
    __shared__ double s_D[BB][BB + 1];
    __shared__ double s_rd[BB];
    __shared__ __align__(16) double s_prow[2][BB];
    const int4 me = cta_map[blockIdx.x];  // (group = front of this launch, rank in the group, group size)
    const int grp = me.x, g = me.y, G = me.z, tid = threadIdx.x;
    const int s = nodes[grp];
    const int m = d.m[s], np = d.np[s];
    double *F = d.F + d.front_ptr[s];
    unsigned *cnt = bar + grp;
    unsigned target = 0;
    long long tq0 = 0, pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // optional phase timing (OCP_MF_PROF): CTA 0 of group 0
    const bool timing = prof != nullptr && blockIdx.x == 0 && tid == 0;
#define BIG_TICK(k) do { if (timing) { long long t_ = clock64(); pacc[k] += t_ - tq0; tq0 = t_; } } while (0)
    if (timing) tq0 = clock64();
    // ---- extend-add of the children's update matrices: one child at a time, so that every entry of the front has a
    // single writer (plain read-modify-write streams at L2 bandwidth; fp64 atomics are several times slower)
    for (int ci = d.child_ptr[s]; ci < d.child_ptr[s + 1]; ++ci) {
        const int c = d.child[ci], mc = d.m[c], npc = d.np[c], nu = mc - npc;
        const double *Fc = d.F + d.front_ptr[c];
        const int *rel = d.rel + d.rel_ptr[c];
        for (int j = g; j < nu; j += G) {
            const double *src = Fc + npc + (size_t)(npc + j) * mc;
            double *dst = F + (size_t)__ldg(rel + j) * m;
            for (int i = tid; i < nu; i += 4 * BIG_T) {      // four independent read-modify-writes in flight per thread
                double *p[4];
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int iu = i + u * BIG_T;
                    p[u] = dst + (iu < nu ? __ldg(rel + iu) : 0);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int iu = i + u * BIG_T;
                    v[u] = iu < nu ? __ldcg(p[u]) + __ldcg(src + iu) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i + u * BIG_T < nu) __stcg(p[u], v[u]);
            }
        }
        target += G;
        group_barrier(cnt, target, G);
    }
    if (d.child_ptr[s] == d.child_ptr[s + 1]) {
        target += G;
        group_barrier(cnt, target, G);
    }
    BIG_TICK(0);
    const int nblk_total = (m + 3) >> 2;
    for (int k0 = 0; k0 < np; k0 += BB) {
        const int kb = min(BB, np - k0), ctrail = k0 + kb, nrows = m - ctrail;
        // ---- A1: diagonal block, lane = row (padded with the identity)
        if (tid < 32) {
            double r[BB];
#pragma unroll
            for (int jj = 0; jj < BB; ++jj)
                r[jj] = (tid < kb && jj < kb) ? __ldcg(F + (k0 + tid) + (size_t)(k0 + jj) * m) : ((jj == tid) ? 1.0 : 0.0);
            // dependent chain per step: pivot -> reciprocal -> multiplier -> next pivot; the next pivot column is updated
            // first and its reciprocal started before the rest of the row is touched (as in the small-front kernel)
            double rinv = fast_rcp(__shfl_sync(0xffffffffu, r[0], 0));
#pragma unroll
            for (int j = 0; j < BB; ++j) {
                double *prow = s_prow[j & 1];                 // the pivot row goes through shared memory (see mf_factor_kernel)
                if (tid == j) {
                    s_rd[j] = rinv;
                    if (j < kb && !(fabs(rinv) < 1e280)) atomicExch(info, s + 1);
#pragma unroll
                    for (int jj = 0; jj < BB; jj += 2)
                        if (jj + 1 > j) *reinterpret_cast<double2 *>(prow + jj) = make_double2(r[jj], r[jj + 1]);
                }
                __syncwarp();
                const bool below = tid > j;
                const double l = r[j] * rinv;
                if (below) r[j] = l;
                double rnext = 0.0;
                if (j + 1 < BB) {
                    const double u1 = prow[j + 1];
                    if (below) r[j + 1] = fma(-l, u1, r[j + 1]);
                    rnext = fast_rcp(__shfl_sync(0xffffffffu, r[j + 1], j + 1));
                }
#pragma unroll
                for (int jj = 0; jj < BB; ++jj) {
                    if (jj > j + 1) {
                        const double u = prow[jj];
                        if (below) r[jj] = fma(-l, u, r[jj]);
                    }
                }
                rinv = rnext;
            }
#pragma unroll
            for (int jj = 0; jj < BB; ++jj) s_D[tid][jj] = r[jj];
        }
        __syncthreads();
        BIG_TICK(1);
        // ---- A2: this CTA's share of the rows below the block: L = A U11^{-1}
        {
            const int chunk = (((nrows + G - 1) / G) + 31) & ~31;
            const int rbeg = g * chunk, rend = min(nrows, rbeg + chunk);
            for (int i = rbeg + tid; i < rend; i += BIG_T) {
                double *rowp = F + ctrail + i + (size_t)k0 * m;
                double a[BB];
#pragma unroll
                for (int jj = 0; jj < BB; ++jj) a[jj] = jj < kb ? __ldcg(rowp + (size_t)jj * m) : 0.0;
#pragma unroll
                for (int t = 0; t < BB; ++t) {
                    if (t < kb) {
                        const double l = a[t] * s_rd[t];
                        a[t] = l;
#pragma unroll
                        for (int jj = 0; jj < BB; ++jj)
                            if (jj > t) a[jj] = fma(-l, s_D[t][jj], a[jj]);
                    }
                }
#pragma unroll
                for (int jj = 0; jj < BB; ++jj)
                    if (jj < kb) __stcg(rowp + (size_t)jj * m, a[jj]);
            }
        }
        if (prof != nullptr && blockIdx.x == 0) __syncthreads();
        BIG_TICK(2);
        target += G;
        group_barrier(cnt, target, G);
        BIG_TICK(3);
        // ---- B0: one CTA stores the factored block
        if (g == (k0 / BB) % G) {
            for (int e = tid; e < kb * kb; e += BIG_T) {
                const int i = e % kb, j = e / kb;
                __stcg(F + (k0 + i) + (size_t)(k0 + j) * m, s_D[i][j]);
            }
        }
        if (nrows > 0) {
            // ---- B1: U12 = L11^{-1} F12 for the own trailing columns (one column per thread)
            const int cb0 = ctrail >> 2;
            const int cbf = cb0 + ((g - cb0 % G) + G) % G;                 // first own 4-column block
            const int nown = cbf < nblk_total ? (nblk_total - 1 - cbf) / G + 1 : 0;
            for (int idx = tid; idx < nown * 4; idx += BIG_T) {
                const int c = (cbf + (idx >> 2) * G) * 4 + (idx & 3);
                if (c >= ctrail && c < m) {
                    double *colp = F + (size_t)c * m + k0;
                    double u[BB];
#pragma unroll
                    for (int t = 0; t < BB; ++t) u[t] = t < kb ? __ldcg(colp + t) : 0.0;
#pragma unroll
                    for (int t = 0; t < BB; ++t) {
                        if (t < kb) {
#pragma unroll
                            for (int tt = 0; tt < BB; ++tt)
                                if (tt > t) u[tt] = fma(-s_D[tt][t], u[t], u[tt]);
                        }
                    }
#pragma unroll
                    for (int t = 0; t < BB; ++t)
                        if (t < kb) __stcg(colp + t, u[t]);
                }
            }
            if (prof != nullptr && blockIdx.x == 0) __syncthreads();
            BIG_TICK(4);
            // ---- B2: rank-kb update of the own columns
            const int ncols_all = nown * 4;
            const int npass = (nrows + 2 * BIG_T - 1) / (2 * BIG_T);
            const int rpp = ((((nrows + npass - 1) / npass) + 63) & ~63) >> 1;    // rows per pass / 2 (multiple of 32)
            for (int sp0 = 0; sp0 < ncols_all; sp0 += BIG_UC) {
                const int ncols = min(BIG_UC, ncols_all - sp0), qbase = sp0 >> 2, nq = ncols >> 2;
                __syncthreads();   // U12 written above / previous super-pass done with Us
                for (int e = tid; e < BB * ((ncols + 7) & ~7); e += BIG_T) {     // (an odd block count is padded with zeros)
                    const int t = e % BB, ci = e / BB;
                    const int c = (cbf + (qbase + (ci >> 2)) * G) * 4 + (ci & 3);
                    Us[e] = (ci < ncols && c >= ctrail && c < m && t < kb) ? __ldcg(F + (size_t)c * m + k0 + t) : 0.0;
                }
                __syncthreads();
                for (int rp = 0; rp < npass; ++rp) {
                    const int ia = rp * 2 * rpp + tid, ib2 = ia + rpp;             // the two rows of this thread
                    if (tid >= rpp || ia >= nrows) continue;                       // (no barrier inside this loop)
                    const bool vb = ib2 < nrows;
                    double la[BB], lb[BB];                                        // negated L rows
                    {
                        const double *lp = F + ctrail + (size_t)k0 * m;
#pragma unroll
                        for (int t = 0; t < BB; ++t) {
                            la[t] = t < kb ? -__ldcg(lp + ia + (size_t)t * m) : 0.0;
                            lb[t] = (t < kb && vb) ? -__ldcg(lp + ib2 + (size_t)t * m) : 0.0;
                        }
                    }
                    double *rowa = F + ctrail + ia, *rowb = F + ctrail + ib2;
                    double fa[8], fb[8];
                    auto fetch = [&](int q2, double *xa, double *xb) {           // C values of the block pair q2
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int q = 2 * q2 + h;
                            const int cb = (cbf + (qbase + q) * G) * 4;
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                const bool cv = q < nq && cb + b >= ctrail && cb + b < m;
                                xa[4 * h + b] = cv ? __ldcg(rowa + (size_t)(cb + b) * m) : 0.0;
                                xb[4 * h + b] = (cv && vb) ? __ldcg(rowb + (size_t)(cb + b) * m) : 0.0;
                            }
                        }
                    };
                    fetch(0, fa, fb);
                    const int nq2 = (nq + 1) >> 1;
                    for (int q2 = 0; q2 < nq2; ++q2) {
                        double xa[8], xb[8], na[8], nb[8];
#pragma unroll
                        for (int b = 0; b < 8; ++b) {
                            xa[b] = fa[b];
                            xb[b] = fb[b];
                        }
                        fetch(q2 + 1, na, nb);                                     // next pair's C values in flight
                        const double *up = Us + (size_t)q2 * 8 * BB;              // (an odd block count is zero padded)
#pragma unroll
                        for (int t = 0; t < BB; t += 2) {
#pragma unroll
                            for (int b = 0; b < 8; ++b) {
                                const double2 u = *reinterpret_cast<const double2 *>(up + b * BB + t);
                                xa[b] = fma(la[t], u.x, xa[b]);
                                xb[b] = fma(lb[t], u.x, xb[b]);
                                xa[b] = fma(la[t + 1], u.y, xa[b]);
                                xb[b] = fma(lb[t + 1], u.y, xb[b]);
                            }
                        }
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int q = 2 * q2 + h;
                            const int cb = (cbf + (qbase + q) * G) * 4;
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                if (q < nq && cb + b >= ctrail && cb + b < m) {
                                    __stcg(rowa + (size_t)(cb + b) * m, xa[4 * h + b]);
                                    if (vb) __stcg(rowb + (size_t)(cb + b) * m, xb[4 * h + b]);
                                }
                            }
                        }
#pragma unroll
                        for (int b = 0; b < 8; ++b) {
                            fa[b] = na[b];
                            fb[b] = nb[b];
                        }
                    }
                }
            }
        }
        if (prof != nullptr && blockIdx.x == 0) __syncthreads();
        BIG_TICK(5);
        target += G;
        group_barrier(cnt, target, G);
        BIG_TICK(6);
    }
    if (timing)
        for (int k = 0; k < 7; ++k) prof[k] = pacc[k];
#undef BIG_TICK
}

// Triangular solves of a large front: one CTA of 1024 threads, right-hand side(s) in shared memory, 16-row blocks
// with the stored inverses of the diagonal blocks as in the small-front kernels, factor entries straight from L2.
constexpr int TSB = 1024;

template <bool TR, int NR>
__global__ void __launch_bounds__(TSB)
mf_forward_big_kernel(MFDev d, const int *__restrict__ nodes, double *__restrict__ x, int ldx) {
    extern __shared__ double y[];          // NR vectors of length m
    __shared__ double yb[NR][NB];
    const int s = nodes[blockIdx.x];
    const int m = d.m[s], np = d.np[s], tid = threadIdx.x;
    if (np == 0) return;
    const double *F = d.F + d.front_ptr[s];
    const int *I = d.idx + d.idx_ptr[s];
    const double *Dinv = d.dinv + (size_t)d.dinv_ptr[s] * (2 * NB * NB) + (TR ? NB * NB : 0);
    for (int r = 0; r < NR; ++r)
        for (int k = tid; k < m; k += TSB) y[r * m + k] = k < np ? x[(size_t)r * ldx + I[k]] : 0.0;
    __syncthreads();
    const int nblk = (np + NB - 1) / NB;
    for (int b = 0; b < nblk; ++b) {
        const int k0 = b * NB, kb = min(NB, np - k0);
        if (tid < NB * NR) {
            const int r = tid / NB, row = tid % NB;
            const double *Db = Dinv + (size_t)b * (2 * NB * NB);
            double v = 0.0;
#pragma unroll
            for (int t = 0; t < NB; ++t) v = fma(__ldcg(Db + MF_DI(row, t)), (t < kb) ? y[r * m + k0 + t] : 0.0, v);
            yb[r][row] = v;
        }
        __syncthreads();
        if (tid < NB * NR && (tid % NB) < kb) y[(tid / NB) * m + k0 + (tid % NB)] = yb[tid / NB][tid % NB];
        for (int i = k0 + kb + tid; i < m; i += TSB) {
            double e[NB];
#pragma unroll
            for (int t = 0; t < NB; ++t) e[t] = (t < kb) ? MF_E(i, k0 + t) : 0.0;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                double acc = 0.0;
#pragma unroll
                for (int t = 0; t < NB; ++t) acc = fma(e[t], yb[r][t], acc);
                y[r * m + i] -= acc;
            }
        }
        __syncthreads();
    }
    for (int r = 0; r < NR; ++r) {
        for (int k = tid; k < np; k += TSB) x[(size_t)r * ldx + I[k]] = y[r * m + k];
        for (int i = np + tid; i < m; i += TSB) atomicAdd(x + (size_t)r * ldx + I[i], y[r * m + i]);
    }
}

template <bool TR, int NR>
__global__ void __launch_bounds__(TSB)
mf_backward_big_kernel(MFDev d, const int *__restrict__ nodes, double *__restrict__ x, int ldx) {
    extern __shared__ double y[];
    __shared__ double yb[NR][NB];
    const int s = nodes[blockIdx.x];
    const int m = d.m[s], np = d.np[s], tid = threadIdx.x;
    if (np == 0) return;
    const double *F = d.F + d.front_ptr[s];
    const int *I = d.idx + d.idx_ptr[s];
    const double *Dinv = d.dinv + (size_t)d.dinv_ptr[s] * (2 * NB * NB) + (TR ? 0 : NB * NB);
    for (int r = 0; r < NR; ++r)
        for (int k = tid; k < m; k += TSB) y[r * m + k] = x[(size_t)r * ldx + I[k]];
    __syncthreads();
    // y_P -= U12 x_U (TR: L21^T x_U)
    if (!TR) {
        for (int k = tid; k < np; k += TSB) {       // lanes over the rows: coalesced column reads
            double a[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) a[r] = 0.0;
            int j = np;
            for (; j + 8 <= m; j += 8) {
                double e[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) e[u] = __ldcg(F + k + (size_t)(j + u) * m);
#pragma unroll
                for (int u = 0; u < 8; ++u)
#pragma unroll
                    for (int r = 0; r < NR; ++r) a[r] = fma(e[u], y[r * m + j + u], a[r]);
            }
            for (; j < m; ++j) {
                const double e = __ldcg(F + k + (size_t)j * m);
#pragma unroll
                for (int r = 0; r < NR; ++r) a[r] = fma(e, y[r * m + j], a[r]);
            }
#pragma unroll
            for (int r = 0; r < NR; ++r) y[r * m + k] -= a[r];
        }
    } else {
        const int lane = tid & 31, wid = tid >> 5, NW = TSB / 32;
        for (int k = wid; k < np; k += NW) {        // the transposed factor is contiguous along j
            double a[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) a[r] = 0.0;
            for (int j = np + lane; j < m; j += 32) {
                const double e = __ldcg(F + j + (size_t)k * m);
#pragma unroll
                for (int r = 0; r < NR; ++r) a[r] = fma(e, y[r * m + j], a[r]);
            }
#pragma unroll
            for (int r = 0; r < NR; ++r) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a[r] += __shfl_xor_sync(0xffffffffu, a[r], o);
                if (lane == 0) y[r * m + k] -= a[r];
            }
        }
    }
    __syncthreads();
    const int nblk = (np + NB - 1) / NB;
    for (int b = nblk - 1; b >= 0; --b) {
        const int k0 = b * NB, kb = min(NB, np - k0);
        if (tid < NB * NR) {
            const int r = tid / NB, row = tid % NB;
            const double *Db = Dinv + (size_t)b * (2 * NB * NB);
            double v = 0.0;
#pragma unroll
            for (int t = 0; t < NB; ++t) v = fma(__ldcg(Db + MF_DI(row, t)), (t < kb) ? y[r * m + k0 + t] : 0.0, v);
            yb[r][row] = v;
        }
        __syncthreads();
        if (tid < NB * NR && (tid % NB) < kb) y[(tid / NB) * m + k0 + (tid % NB)] = yb[tid / NB][tid % NB];
        for (int i = tid; i < k0; i += TSB) {
            double e[NB];
#pragma unroll
            for (int t = 0; t < NB; ++t) e[t] = (t < kb) ? MF_E(i, k0 + t) : 0.0;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                double acc = 0.0;
#pragma unroll
                for (int t = 0; t < NB; ++t) acc = fma(e[t], yb[r][t], acc);
                y[r * m + i] -= acc;
            }
        }
        __syncthreads();
    }
    for (int r = 0; r < NR; ++r)
        for (int k = tid; k < np; k += TSB) x[(size_t)r * ldx + I[k]] = y[r * m + k];
}

// kernel variants: <threads, panel rows per thread>
// kernel variants: <threads, panel rows per thread>; one warp of every CTA is the look-ahead warp, the others take one
// or two panel rows each.  The thread counts keep the register budget per thread above what the kernel needs
// (65536 / 288 = 227, / 416 = 157) - at 512 threads (128 registers) it spills.
enum { kVar288 = 0, kVar416 = 1, kVar512 = 2, kVar512x2 = 3, kVar256x2 = 4, kVar384x2 = 5, kNumVariants = 6 };
// Register budget: a warp scheduler owns 16384 registers and the warps of a CTA are dealt round-robin, so a CTA of
// 9 warps (288 threads) may use 168 registers per thread, 12 warps (384) 168, but 13 warps (416) or 16 (512) only 128 -
// `ptxas -v` (profiles/ptxas_r2.txt) shows the 416- and 512-thread variants spilling 300-500 bytes per thread.  The
// x2 variants give every worker two panel rows instead: 256 threads (7 worker warps, 255 registers) cover fronts of
// order <= 448, 384 threads (11 worker warps, 168 registers) <= 704.  OCP_MF_THREADS selects: "auto" (default, the
// measured best), "legacy" (288 / 416 / 512), "256", "384".
inline int factor_variant(int max_m) {
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("OCP_MF_THREADS");
        const std::string v = e ? e : "auto";
        mode = v == "legacy" ? 0 : (v == "256" ? 1 : (v == "384" ? 2 : 3));
    }
    const int legacy = max_m <= 256 ? kVar288 : (max_m <= 384 ? kVar416 : (max_m <= 480 ? kVar512 : kVar512x2));
    if (mode == 1) return max_m <= 448 ? kVar256x2 : (max_m <= 704 ? kVar384x2 : kVar512x2);
    if (mode == 2) return max_m <= 704 ? kVar384x2 : kVar512x2;
    if (mode == 3) return MF_AUTO_VARIANT(max_m, legacy);
    return legacy;
}
inline int variant_threads(int v) {
    return v == kVar288 ? 288 : (v == kVar416 ? 416 : (v == kVar256x2 ? 256 : (v == kVar384x2 ? 384 : 512)));
}
typedef void (*FactorKernel)(MFDev, const int *, int, int *, long long *);
inline FactorKernel factor_kernel(int v) {
    switch (v) {
        case kVar288: return mf_factor_kernel<288, 1>;
        case kVar416: return mf_factor_kernel<416, 1>;
        case kVar512: return mf_factor_kernel<512, 1>;
        case kVar256x2: return mf_factor_kernel<256, 2>;
        case kVar384x2: return mf_factor_kernel<384, 2>;
        default: return mf_factor_kernel<512, 2>;
    }
}

template <class T>
bool up(T **dst, const std::vector<T> &src, std::string &err) {
    cudaError_t e = cudaMalloc((void **)dst, sizeof(T) * std::max<size_t>(src.size(), 1));
    if (e == cudaSuccess && !src.empty())
        e = cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        err = std::string("multifrontal upload: ") + cudaGetErrorString(e);
        return false;
    }
    return true;
}

}  // namespace

struct MultifrontalLU::Impl {
    MFSymbolic S;
    int *m = nullptr, *np = nullptr, *first = nullptr, *idx_ptr = nullptr, *idx = nullptr, *child_ptr = nullptr,
        *child = nullptr, *rel_ptr = nullptr, *rel = nullptr, *level_nodes = nullptr, *piv = nullptr, *info = nullptr;
    long long *front_ptr = nullptr, *a_dest = nullptr;
    double *F = nullptr, *dinv = nullptr;
    int *dinv_ptr = nullptr;
    long long *prof = nullptr;   // optional per-level phase cycle counters (OCP_MF_PROF=1)
    std::vector<int> level_max_m, level_max_np, level_cluster;   // over the SMALL fronts of a level
    // per level the node list is [small fronts | large fronts]; large = order above kBigM (group kernels)
    std::vector<int> level_off, level_nsmall, level_nbig, level_big_max_m;
    unsigned *bar = nullptr;     // group-barrier counters of the large-front factor kernel
    int big_ctas = 0;            // co-resident CTAs available to that kernel
    struct BigLaunch {           // one cooperative launch: fronts [first, first + nb) of the level's large list
        int level, first, nb, grid, map_off;
    };
    std::vector<BigLaunch> big_launches;
    int *panel_node = nullptr;   // front of every 16-pivot block (mf_dinv_kernel)
    int npanels = 0;
    double *dinv64 = nullptr;    // inverses of the 64 x 64 diagonal blocks of the small fronts (mf_dinv64_kernel)
    int *dinv64_ptr = nullptr, *block_node = nullptr;
    int nblocks64 = 0;
    bool solve64 = true;         // 64-row-block solve kernels for the small fronts (OCP_MF_SOLVE16=1: the 16-row ones)
    bool pdl = true;             // programmatic dependent launches between the levels of a sweep (OCP_MF_PDL=0: plain)
    // bottom levels of the tree whose fronts take the shared-memory-resident single-CTA kernel (mf_leaf_factor_kernel)
    std::vector<char> level_leaf;
    std::vector<size_t> level_leaf_smem;
    // diagonal-block inverses of a level formed on a side stream while the next levels are factored (OCP_MF_OVERLAP=1).
    // Measured on B200 and NOT the default: the side kernels take SM time from the level kernels, 2 factorisations
    // 1.73 -> 1.88 ms per GD iteration; the default forms all inverses in one launch behind the factorisation.
    bool overlap_dinv = false;
    // ONE fork instead: the inverses of all blocks below the top `overlap_top` levels are formed on the side stream
    // while those levels - which leave most SMs idle (4, 2, 1 fronts) - are factored (OCP_MF_OVERLAP_FROM = levels from
    // the top that run beside the inverse kernel; 0 = off).  Measured on B200 and NOT the default either: 2
    // factorisations 1.504 -> 1.538 / 1.549 / 1.544 ms for 1 / 2 / 3 levels (profiles/README.md).
    int overlap_top = 0;
    bool dinv_pair = false;      // the 16- and 64-block inverse kernels behind the factorisation side by side (OCP_MF_DINV_PAIR=1; measured: no gain)
    std::vector<int> level_blk_off;
    cudaStream_t side = nullptr;
    std::vector<cudaEvent_t> ev_level;
    cudaEvent_t ev_join = nullptr;
    bool ensure_side(int nlevels) {
        if (side && (int)ev_level.size() >= nlevels) return true;
        if (!side && cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            side = nullptr;
            overlap_dinv = false;
            return false;
        }
        if (!ev_join && cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            overlap_dinv = false;
            return false;
        }
        while ((int)ev_level.size() < nlevels) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                overlap_dinv = false;
                return false;
            }
            ev_level.push_back(e);
        }
        return true;
    }
    int4 *cta_map = nullptr;     // per CTA of every large-front launch: (front of the launch, rank in its group, group size)
    // CUDA graphs of the factor / solve launch sequences, keyed by the (fixed) device pointer they operate on
    std::map<const void *, cudaGraphExec_t> factor_graphs, solve_graphs, solve_t_graphs, solve4_graphs;
    cudaStream_t cap_stream = nullptr;
    bool use_graphs = true;
    bool direct = false;         // enqueue directly (an outer graph is being captured)
    int *h_info = nullptr;       // pinned: zero-pivot flag of the most recent factorisations (checked lazily)
    MFDev dev{};
    ~Impl() {
        void *p[] = {m, np, first, idx_ptr, idx, child_ptr, child, rel_ptr, rel, level_nodes, piv, info, front_ptr,
                     a_dest, F, prof, dinv, dinv_ptr, bar, cta_map, panel_node, dinv64, dinv64_ptr, block_node};
        for (void *q : p) cudaFree(q);
        for (auto &kv : factor_graphs) cudaGraphExecDestroy(kv.second);
        for (auto &kv : solve_graphs) cudaGraphExecDestroy(kv.second);
        for (auto &kv : solve_t_graphs) cudaGraphExecDestroy(kv.second);
        for (auto &kv : solve4_graphs) cudaGraphExecDestroy(kv.second);
        if (cap_stream) cudaStreamDestroy(cap_stream);
        for (cudaEvent_t e : ev_level) cudaEventDestroy(e);
        if (ev_join) cudaEventDestroy(ev_join);
        if (side) cudaStreamDestroy(side);
        if (h_info) cudaFreeHost(h_info);
    }
    bool enqueue_factor(const double *d_vals, int nnz, cudaStream_t s, std::string &err);
    bool enqueue_solve(double *d_x, int variant, cudaStream_t s, std::string &err);
    template <class Fn>
    bool run(std::map<const void *, cudaGraphExec_t> &cache, const void *key, cudaStream_t s, std::string &err, Fn enqueue) {
        if (!use_graphs || prof || direct) return enqueue(s);
        auto it = cache.find(key);
        if (it == cache.end()) {
            cudaGraph_t g = nullptr;
            cudaGraphExec_t ge = nullptr;
            if (cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
                cudaGetLastError();
                use_graphs = false;
                return enqueue(s);
            }
            const bool ok = enqueue(cap_stream);
            cudaError_t e = cudaStreamEndCapture(cap_stream, &g);
            if (ok && e == cudaSuccess) e = cudaGraphInstantiate(&ge, g, 0);
            if (g) cudaGraphDestroy(g);
            if (!ok || e != cudaSuccess) {
                cudaGetLastError();
                use_graphs = false;      // capture not possible here: plain launches from now on
                return enqueue(s);
            }
            it = cache.emplace(key, ge).first;
        }
        cudaError_t e = cudaGraphLaunch(it->second, s);
        if (e != cudaSuccess) {
            err = std::string("multifrontal graph launch: ") + cudaGetErrorString(e);
            return false;
        }
        return true;
    }
};

MultifrontalLU::MultifrontalLU() = default;
void MultifrontalLU::set_direct_enqueue(bool on) {
    if (impl_) impl_->direct = on;
}
bool MultifrontalLU::needs_cooperative_launch() const { return impl_ && !impl_->big_launches.empty(); }
MultifrontalLU::~MultifrontalLU() { delete impl_; }

bool MultifrontalLU::configure(int n, int nnz, const int *h_rowptr, const int *h_col, const double *xy,
                               const unsigned char *kind, std::string &err, const MultifrontalLU *share) {
    auto t0 = std::chrono::steady_clock::now();
    delete impl_;
    impl_ = new Impl();
    Impl &I = *impl_;
    n_ = n;
    nnz_ = nnz;
    if (share && share->impl_ && share->n_ == n && share->nnz_ == nnz)
        I.S = share->impl_->S;
    else
        mf_analyse(n, h_rowptr, h_col, xy, kind, mf_leaf_size(), I.S);
    const MFSymbolic &S = I.S;
    for (long long dd : S.a_dest)
        if (dd < 0) {
            err = "multifrontal analysis: matrix entry outside its front";
            return false;
        }
    // split every level into small fronts (cluster kernel, panel in shared memory) and large fronts (group kernel)
    I.level_max_m.assign(S.nlevels, 0);
    I.level_max_np.assign(S.nlevels, 0);
    I.level_off.assign(S.nlevels, 0);
    I.level_nsmall.assign(S.nlevels, 0);
    I.level_nbig.assign(S.nlevels, 0);
    I.level_big_max_m.assign(S.nlevels, 0);
    std::vector<int> ordered(S.nnodes);
    int big_limit = kBigM;
    if (const char *eb = getenv("OCP_MF_BIG")) big_limit = std::max(32, atoi(eb));   // testing: force the group path
    for (int l = 0; l < S.nlevels; ++l) {
        int w = S.level_ptr[l];
        I.level_off[l] = w;
        for (int pass = 0; pass < 2; ++pass)
            for (int k = S.level_ptr[l]; k < S.level_ptr[l + 1]; ++k) {
                const int nd = S.level_nodes[k];
                const bool big = S.m[nd] > big_limit;
                if (big != (pass == 1)) continue;
                ordered[w++] = nd;
                if (big) {
                    I.level_nbig[l]++;
                    I.level_big_max_m[l] = std::max(I.level_big_max_m[l], S.m[nd]);
                } else {
                    I.level_nsmall[l]++;
                    I.level_max_m[l] = std::max(I.level_max_m[l], S.m[nd]);
                    I.level_max_np[l] = std::max(I.level_max_np[l], S.np[nd]);
                }
            }
    }
    {
        // OCP_MF_LEAF_LEVELS: how many bottom levels may take the shared-memory-resident kernel [1]; a level qualifies
        // when it has no large front and every front's pivot columns / rows fit one CTA's shared memory
        int leaf_levels = 1;
        if (const char *el = getenv("OCP_MF_LEAF_LEVELS")) leaf_levels = std::max(0, atoi(el));
        I.level_leaf.assign(S.nlevels, 0);
        I.level_leaf_smem.assign(S.nlevels, 0);
        for (int l = 0; l < S.nlevels && l < leaf_levels; ++l) {
            if (I.level_nbig[l] > 0 || I.level_nsmall[l] == 0) break;
            size_t need = 0;
            for (int k = 0; k < I.level_nsmall[l]; ++k) {
                const int nd = ordered[I.level_off[l] + k];
                need = std::max(need, leaf_smem_doubles(S.m[nd], S.np[nd]) * sizeof(double));
            }
            if (need > kLeafSmemMax) break;
            I.level_leaf[l] = 1;
            I.level_leaf_smem[l] = need;
        }
    }
    if (S.max_front > 24000) {   // right-hand side of a large front must fit the solve kernels' shared memory
        err = "multifrontal: largest front (" + std::to_string(S.max_front) + ") exceeds the solve kernels' shared memory";
        return false;
    }
    if (!up(&I.m, S.m, err) || !up(&I.np, S.np, err) || !up(&I.first, S.first, err) ||
        !up(&I.idx_ptr, S.idx_ptr, err) || !up(&I.idx, S.idx, err) || !up(&I.child_ptr, S.child_ptr, err) ||
        !up(&I.child, S.child, err) || !up(&I.rel_ptr, S.rel_ptr, err) || !up(&I.rel, S.rel, err) ||
        !up(&I.level_nodes, ordered, err) || !up(&I.front_ptr, S.front_ptr, err) ||
        !up(&I.a_dest, S.a_dest, err))
        return false;
    // the front workspace (the bulk of the memory: 5 GB at 256 x 256) is allocated by the first factorisation
    cudaError_t e = cudaMalloc((void **)&I.piv, sizeof(int) * std::max(n, 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&I.info, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(I.info, 0, sizeof(int));
    // the attributes are per kernel, not per solver instance: always allow the full opt-in budget
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_backward_kernel<1, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_backward_kernel<2, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_backward_kernel<1, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_backward_kernel<2, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_backward64_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_backward64_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int v = 0; v < kNumVariants && e == cudaSuccess; ++v) {
        e = cudaFuncSetAttribute(factor_kernel(v), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(factor_kernel(v), cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    }
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_leaf_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLeafSmemMax);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_big_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBigSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_forward_big_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_forward_big_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_forward_big_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_backward_big_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_backward_big_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_backward_big_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) {
        // the group kernel spins on barriers between its CTAs: never launch more of them than can be co-resident
        int dev_id = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev_id);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mf_big_factor_kernel, BIG_T, kBigSmem);
        I.big_ctas = sms * std::min(per_sm, 1);
        if (const char *eg = getenv("OCP_MF_BIG_CTAS")) I.big_ctas = std::max(1, std::min(I.big_ctas, atoi(eg)));
        if (e == cudaSuccess && I.big_ctas < 1) {
            err = "multifrontal setup: the large-front kernel does not fit an SM";
            return false;
        }
        if (e == cudaSuccess) e = cudaMalloc((void **)&I.bar, sizeof(unsigned) * 1024);
        // CTAs are shared out among the large fronts of a launch in proportion to their flops (the orders inside one
        // level differ by more than 2x), at least one each
        std::vector<int4> map;
        for (int l = 0; l < S.nlevels && e == cudaSuccess; ++l) {
            const int *lst = ordered.data() + I.level_off[l] + I.level_nsmall[l];
            for (int done = 0; done < I.level_nbig[l];) {
                const int nb = std::min(I.level_nbig[l] - done, std::min(I.big_ctas, 1024));
                std::vector<double> w(nb);
                double W = 0.0;
                for (int k = 0; k < nb; ++k) {
                    const double pp = S.np[lst[done + k]], mm = S.m[lst[done + k]];
                    w[k] = 2.0 * pp * mm * mm - 2.0 * pp * pp * mm + 2.0 / 3.0 * pp * pp * pp + 1.0;
                    W += w[k];
                }
                std::vector<int> Gk(nb);
                int used = 0;
                for (int k = 0; k < nb; ++k) {
                    Gk[k] = std::max(1, (int)std::floor((I.big_ctas - nb) * w[k] / W) + 1);
                    used += Gk[k];
                }
                while (used > I.big_ctas) {      // rounding overshoot: take from the best-served front
                    int best = -1;
                    for (int k = 0; k < nb; ++k)
                        if (Gk[k] > 1 && (best < 0 || w[k] / Gk[k] < w[best] / Gk[best])) best = k;
                    if (best < 0) break;
                    Gk[best]--;
                    used--;
                }
                while (used < I.big_ctas) {      // leftovers to the front with the most work per CTA
                    int best = 0;
                    for (int k = 1; k < nb; ++k)
                        if (w[k] / Gk[k] > w[best] / Gk[best]) best = k;
                    Gk[best]++;
                    used++;
                }
                Impl::BigLaunch bl{l, done, nb, used, (int)map.size()};
                for (int k = 0; k < nb; ++k)
                    for (int r = 0; r < Gk[k]; ++r) map.push_back(make_int4(k, r, Gk[k], 0));
                I.big_launches.push_back(bl);
                done += nb;
            }
        }
        if (e == cudaSuccess && !up(&I.cta_map, map, err)) return false;
    }
    // cluster size per level: as many CTAs per front as the chip has room for (powers of two, <= 16)
    int max_cluster = 16;
    if (const char *envc = getenv("OCP_MF_MAX_CLUSTER")) max_cluster = std::max(1, atoi(envc));
    I.level_cluster.assign(S.nlevels, 1);
    for (int l = 0; l < S.nlevels && e == cudaSuccess; ++l) {
        const int nf = I.level_nsmall[l];
        if (nf == 0) continue;
        int c = 1;
        while (c * 2 <= max_cluster && nf * c * 2 <= 148 && c * 2 * 8 <= I.level_max_m[l]) c *= 2;
        while (c > 1) {   // make sure the cluster shape is launchable with this kernel's resources
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(nf * c);
            cfg.blockDim = dim3(variant_threads(factor_variant(I.level_max_m[l])));
            cfg.dynamicSmemBytes = (size_t)NB * (2 * (size_t)((I.level_max_m[l] + 3) & ~3) + CWO + 4) * sizeof(double);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = c;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int ncl = 0;
            cudaError_t oe = cudaOccupancyMaxActiveClusters(&ncl, factor_kernel(factor_variant(I.level_max_m[l])), &cfg);
            if (oe == cudaSuccess && ncl >= 1) break;
            cudaGetLastError();
            c /= 2;
        }
        I.level_cluster[l] = c;
    }
    if (e != cudaSuccess) {
        err = std::string("multifrontal setup: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return false;
    }
    if (const char *eg = getenv("OCP_MF_GRAPHS")) I.use_graphs = atoi(eg) != 0;
    if (cudaStreamCreateWithFlags(&I.cap_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost((void **)&I.h_info, sizeof(int)) != cudaSuccess) {
        err = "multifrontal setup: stream / pinned allocation failed";
        return false;
    }
    *I.h_info = 0;
    if (getenv("OCP_MF_PROF")) {
        cudaMalloc((void **)&I.prof, sizeof(long long) * 8 * S.nlevels);
        cudaMemset(I.prof, 0, sizeof(long long) * 8 * S.nlevels);
    }
    {
        std::vector<int> dp(S.nnodes + 1, 0);
        for (int k = 0; k < S.nnodes; ++k) dp[k + 1] = dp[k] + (S.np[k] + NB - 1) / NB;
        if (!up(&I.dinv_ptr, dp, err)) return false;
        std::vector<int> pn(dp[S.nnodes]);
        for (int k = 0; k < S.nnodes; ++k)
            for (int q = dp[k]; q < dp[k + 1]; ++q) pn[q] = k;
        I.npanels = dp[S.nnodes];
        if (!up(&I.panel_node, pn, err)) return false;
        if (cudaMalloc((void **)&I.dinv, sizeof(double) * 2 * NB * NB * std::max(dp[S.nnodes], 1)) != cudaSuccess) {
            err = "multifrontal setup: out of memory";
            return false;
        }
    }
    {
        // 64-pivot blocks of the SMALL fronts (the large fronts keep the 16-row solve kernels)
        if (const char *e16 = getenv("OCP_MF_SOLVE16")) I.solve64 = atoi(e16) == 0;
        if (const char *ep = getenv("OCP_MF_PDL")) I.pdl = atoi(ep) != 0;
        if (const char *eo = getenv("OCP_MF_OVERLAP")) I.overlap_dinv = atoi(eo) != 0;
        if (const char *eo = getenv("OCP_MF_DINV_PAIR")) I.dinv_pair = atoi(eo) != 0;
        if (const char *eo = getenv("OCP_MF_OVERLAP_FROM")) I.overlap_top = std::max(0, atoi(eo));
        // blocks are numbered level by level (the order of the launch lists), so that the inverses of a level's blocks
        // can be formed on the side stream while the next level is being factored (enqueue_factor)
        std::vector<int> dp(S.nnodes + 1, 0), bn;
        I.level_blk_off.assign(S.nlevels + 1, 0);
        for (int l = 0; l < S.nlevels; ++l) {
            I.level_blk_off[l] = (int)bn.size();
            for (int k = S.level_ptr[l]; k < S.level_ptr[l + 1]; ++k) {
                const int nd = ordered[k];
                dp[nd] = (int)bn.size();
                const int nb = S.m[nd] > big_limit ? 0 : (S.np[nd] + SB - 1) / SB;
                for (int q = 0; q < nb; ++q) bn.push_back(nd);
            }
        }
        I.level_blk_off[S.nlevels] = dp[S.nnodes] = (int)bn.size();
        I.nblocks64 = (int)bn.size();
        if (!up(&I.dinv64_ptr, dp, err) || !up(&I.block_node, bn, err)) return false;
        if (I.solve64 &&
            cudaMalloc((void **)&I.dinv64, sizeof(double) * 2 * SB * SB * std::max(I.nblocks64, 1)) != cudaSuccess) {
            err = "multifrontal setup: out of memory";
            return false;
        }
    }
    I.dev = MFDev{I.m, I.np, I.first, I.idx_ptr, I.idx, I.child_ptr, I.child, I.rel_ptr, I.rel, I.front_ptr, I.F, I.piv,
                  I.dinv, I.dinv_ptr, I.dinv64, I.dinv64_ptr, 0};
    if (const char *ea = getenv("OCP_MF_EA_ATOMIC")) I.dev.ea_atomic = atoi(ea) != 0;
    factor_nnz_ = 0;
    for (int s = 0; s < S.nnodes; ++s)
        factor_nnz_ += (long long)S.m[s] * S.m[s] - (long long)(S.m[s] - S.np[s]) * (S.m[s] - S.np[s]);
    flops_ = S.flops;
    fsize_ = S.fsize;
    nfronts_ = S.nnodes;
    nlevels_ = S.nlevels;
    max_front_ = S.max_front;
    analyse_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return true;
}

bool MultifrontalLU::Impl::enqueue_factor(const double *d_vals, int nnz, cudaStream_t s, std::string &err) {
    const MFSymbolic &S = this->S;
    const bool side_on = overlap_dinv && !prof && ensure_side(S.nlevels);
    // single fork in front of level fork_level (none: -1); only when every block below belongs to a small front
    int fork_level = -1;
    if (!side_on && overlap_top > 0 && !prof && solve64 && nblocks64 > 0 && big_launches.empty() && S.nlevels > overlap_top &&
        level_blk_off[S.nlevels - overlap_top] > 0 && ensure_side(S.nlevels))
        fork_level = S.nlevels - overlap_top;
    cudaMemsetAsync(F, 0, sizeof(double) * S.fsize, s);
    scatter_values_kernel<<<(nnz + 255) / 256, 256, 0, s>>>(nnz, a_dest, d_vals, F);
    for (int l = 0; l < S.nlevels; ++l) {
        const int nf = level_nsmall[l];
        if (l == fork_level) {
            cudaEventRecord(ev_level[0], s);
            cudaStreamWaitEvent(side, ev_level[0], 0);
            mf_dinv64_kernel<<<2 * level_blk_off[l], SB, 0, side>>>(dev, block_node, 0, level_blk_off[l]);
            g_launch_count.fetch_add(1, std::memory_order_relaxed);      // (one inverse launch more than factor() counts)
        }
        if (nf > 0 && level_leaf[l] && !prof) {
            mf_leaf_factor_kernel<<<nf, TL, level_leaf_smem[l], s>>>(dev, level_nodes + level_off[l], info);
        } else if (nf > 0) {
            const int c = level_cluster[l];
            cudaLaunchConfig_t cfg = {};
            const int var = factor_variant(level_max_m[l]);
            cfg.gridDim = dim3(nf * c);
            cfg.blockDim = dim3(variant_threads(var));
            cfg.dynamicSmemBytes = (size_t)NB * (2 * (size_t)((level_max_m[l] + 3) & ~3) + CWO + 4) * sizeof(double);
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = c;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            long long *lprof = this->prof ? this->prof + 8 * l : nullptr;
            const int *lvl = level_nodes + level_off[l];
            cudaError_t le = cudaLaunchKernelEx(&cfg, factor_kernel(var), dev, lvl, level_max_m[l], info, lprof);
            if (le != cudaSuccess) {
                err = std::string("multifrontal factor launch (level ") + std::to_string(l) + ", cluster " +
                      std::to_string(c) + "): " + cudaGetErrorString(le);
                return false;
            }
        }
        // large fronts of the level: groups of co-resident CTAs (cooperative launch)
        for (const BigLaunch &bl : big_launches) {
            if (bl.level != l) continue;
            cudaMemsetAsync(bar, 0, sizeof(unsigned) * bl.nb, s);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(bl.grid);
            cfg.blockDim = dim3(BIG_T);
            cfg.dynamicSmemBytes = kBigSmem;
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeCooperative;
            at[0].val.cooperative = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            const int *lvl = level_nodes + level_off[l] + level_nsmall[l] + bl.first;
            long long *bprof = this->prof ? this->prof + 8 * l : nullptr;
            cudaError_t le = cudaLaunchKernelEx(&cfg, mf_big_factor_kernel, dev, lvl, (const int4 *)(cta_map + bl.map_off), bar, info, bprof);
            if (le != cudaSuccess) {
                err = std::string("multifrontal large-front launch (level ") + std::to_string(l) + ", " + std::to_string(bl.nb) +
                      " fronts on " + std::to_string(bl.grid) + " CTAs): " + cudaGetErrorString(le);
                return false;
            }
        }
        if (side_on) {
            // the level's factors are final: their 64 x 64 diagonal inverses go to the side stream (a parallel branch of
            // the captured graph) and run on the SMs the next levels leave idle
            cudaEventRecord(ev_level[l], s);
            cudaStreamWaitEvent(side, ev_level[l], 0);
            const int b0 = level_blk_off[l], nb = level_blk_off[l + 1] - b0;
            if (solve64 && nb > 0) mf_dinv64_kernel<<<2 * nb, SB, 0, side>>>(dev, block_node, b0, nb);
        }
    }
    // the 16 x 16 inverses are only read by the 16-row solve kernels (large fronts, OCP_MF_SOLVE16=1)
    const int check_only = (solve64 && big_launches.empty()) ? 1 : 0;
    // opt-in: the two inverse kernels side by side (a two-branch fork of the captured graph)
    const bool pair_on = !side_on && fork_level < 0 && dinv_pair && !prof && npanels > 0 && solve64 && nblocks64 > 0 && ensure_side(S.nlevels);
    if (pair_on) {
        cudaEventRecord(ev_level[0], s);
        cudaStreamWaitEvent(side, ev_level[0], 0);
    }
    if (npanels > 0)
        mf_dinv_kernel<<<(2 * npanels * 32 + 255) / 256, 256, 0, (side_on || pair_on) ? side : s>>>(dev, panel_node, npanels, info, check_only);
    if (fork_level >= 0) {
        const int b0 = level_blk_off[fork_level];
        if (nblocks64 > b0) mf_dinv64_kernel<<<2 * (nblocks64 - b0), SB, 0, s>>>(dev, block_node, b0, nblocks64 - b0);
    } else if (!side_on && solve64 && nblocks64 > 0) {
        mf_dinv64_kernel<<<2 * nblocks64, SB, 0, s>>>(dev, block_node, 0, nblocks64);
    }
    if (side_on || pair_on || fork_level >= 0) {
        cudaEventRecord(ev_join, side);
        cudaStreamWaitEvent(s, ev_join, 0);
    }
    cudaMemcpyAsync(h_info, info, sizeof(int), cudaMemcpyDeviceToHost, s);
    return true;
}

bool MultifrontalLU::factor(const double *d_vals, cudaStream_t s, std::string &err) {
    if (!impl_) {
        err = "MultifrontalLU::factor before configure";
        return false;
    }
    Impl &I = *impl_;
    if (!check(err)) return false;
    if (!I.F) {
        if (cudaMalloc((void **)&I.F, sizeof(double) * I.S.fsize) != cudaSuccess) {
            cudaGetLastError();
            err = "multifrontal factor: out of memory for the front workspace (" + std::to_string(I.S.fsize * 8 >> 20) + " MiB)";
            return false;
        }
        I.dev.F = I.F;
    }
    g_launch_count.fetch_add(3 + I.S.nlevels + (int)I.big_launches.size(), std::memory_order_relaxed);
    const int nnz = nnz_;
    if (!I.run(I.factor_graphs, d_vals, s, err, [&](cudaStream_t q) { return I.enqueue_factor(d_vals, nnz, q, err); }))
        return false;
    if (I.prof) {
        const MFSymbolic &S = I.S;
        cudaStreamSynchronize(s);
        std::vector<long long> h(8 * S.nlevels);
        cudaMemcpy(h.data(), I.prof, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
        cudaMemset(I.prof, 0, sizeof(long long) * h.size());
        for (int l = 0; l < S.nlevels; ++l)
            if (I.level_nbig[l] > 0 && I.level_nsmall[l] == 0)
                fprintf(stderr, "[mf prof] level %d large fronts %d | CTA0 cycles: extend-add %lld diag %lld Lrows %lld bar1 %lld u12 %lld update %lld bar2 %lld\n",
                        l, I.level_nbig[l], h[8 * l], h[8 * l + 1], h[8 * l + 2], h[8 * l + 3], h[8 * l + 4], h[8 * l + 5], h[8 * l + 6]);
        for (int l = 0; l < S.nlevels; ++l)
            fprintf(stderr, "[mf prof] level %d fronts %d cluster %d | CTA0 cycles: load %lld diag %lld u12 %lld Lrows %lld trail %lld sync %lld\n",
                    l, I.level_nsmall[l], I.level_cluster[l], h[8 * l], h[8 * l + 1], h[8 * l + 2],
                    h[8 * l + 3], h[8 * l + 4], h[8 * l + 5]);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("multifrontal factor: ") + cudaGetErrorString(e);
        return false;
    }
    return true;
}

// Zero-pivot flag of factorisations that have already completed (never blocks: the flag is copied to pinned host
// memory at the end of every factorisation and inspected by the next factor / solve call or by the caller after
// it has synchronised the stream).
bool MultifrontalLU::check(std::string &err) {
    if (impl_ && impl_->h_info && *impl_->h_info != 0) {
        err = "multifrontal factor: zero or tiny pivot in front " + std::to_string(*impl_->h_info - 1) +
              " (static pivoting broke down; run with OCP_SOLVER=rf)";
        *impl_->h_info = 0;
        cudaMemsetAsync(impl_->info, 0, sizeof(int), nullptr);
        return false;
    }
    return true;
}

// A level of the 64-row-block solve kernels; pdl: the previous launch on the stream is such a level of the same sweep,
// so this one may start early (programmatic stream serialisation) and overlap its read-only prologue with it.
template <class K>
static void launch_solve64(K kernel, int nf, size_t smem, cudaStream_t s, bool pdl, const MFDev &dev, const int *nodes,
                           double *d_x, int ldx) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nf);
    cfg.blockDim = dim3(TS6);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, dev, nodes, d_x, ldx);
}

template <bool TR, int NR>
static void launch_level_fwd(const MFDev &dev, const int *nodes, int nf, int max_m, double *d_x, int ldx, cudaStream_t s,
                             bool pdl = false) {
    if (nf <= 0) return;
    if (dev.dinv64) {
        launch_solve64(mf_forward64_kernel<TR, NR>, nf, sizeof(double) * max_m * NR, s, pdl, dev, nodes, d_x, ldx);
        return;
    }
    if (max_m <= TS)
        mf_forward_kernel<1, TR, NR><<<nf, TS, sizeof(double) * max_m * NR, s>>>(dev, nodes, d_x, ldx);
    else
        mf_forward_kernel<2, TR, NR><<<nf, TS, sizeof(double) * max_m * NR, s>>>(dev, nodes, d_x, ldx);
}

template <bool TR, int NR>
static void launch_level_bwd(const MFDev &dev, const int *nodes, int nf, int max_m, int max_np, double *d_x, int ldx,
                             cudaStream_t s, bool pdl = false) {
    if (nf <= 0) return;
    // y (NR x m) plus, for the plain sweep, the per-warp partial sums of the U12 mat-vec (16 x NR x np)
    const size_t smem = sizeof(double) * NR * ((size_t)max_m + (TR ? 0 : (size_t)(TS / 32) * max_np));
    if (dev.dinv64) {
        const size_t smem6 = sizeof(double) * NR * ((size_t)max_m + (TR ? 0 : (size_t)(TS6 / 32) * max_np));
        launch_solve64(mf_backward64_kernel<TR, NR>, nf, smem6, s, pdl, dev, nodes, d_x, ldx);
        return;
    }
    if (max_m <= TS)
        mf_backward_kernel<1, TR, NR><<<nf, TS, smem, s>>>(dev, nodes, d_x, ldx);
    else
        mf_backward_kernel<2, TR, NR><<<nf, TS, smem, s>>>(dev, nodes, d_x, ldx);
}

// large fronts: one CTA each; four right-hand sides that do not fit shared memory together go one at a time
template <bool TR>
static void launch_big_fwd(const MFDev &dev, const int *nodes, int nf, int max_m, int nr, double *d_x, int ldx, cudaStream_t s) {
    if (nf <= 0) return;
    if (nr == 4 && sizeof(double) * 4 * (size_t)max_m <= 196 * 1024) {
        mf_forward_big_kernel<false, 4><<<nf, TSB, sizeof(double) * 4 * max_m, s>>>(dev, nodes, d_x, ldx);
        return;
    }
    for (int r = 0; r < nr; ++r)
        mf_forward_big_kernel<TR, 1><<<nf, TSB, sizeof(double) * max_m, s>>>(dev, nodes, d_x + (size_t)r * ldx, ldx);
}

template <bool TR>
static void launch_big_bwd(const MFDev &dev, const int *nodes, int nf, int max_m, int nr, double *d_x, int ldx, cudaStream_t s) {
    if (nf <= 0) return;
    if (nr == 4 && sizeof(double) * 4 * (size_t)max_m <= 196 * 1024) {
        mf_backward_big_kernel<false, 4><<<nf, TSB, sizeof(double) * 4 * max_m, s>>>(dev, nodes, d_x, ldx);
        return;
    }
    for (int r = 0; r < nr; ++r)
        mf_backward_big_kernel<TR, 1><<<nf, TSB, sizeof(double) * max_m, s>>>(dev, nodes, d_x + (size_t)r * ldx, ldx);
}

// variant: 0 plain, 1 transposed, 2 plain with four right-hand sides
bool MultifrontalLU::Impl::enqueue_solve(double *d_x, int variant, cudaStream_t s, std::string &err) {
    const MFSymbolic &S = this->S;
    const int ldx = S.n;
    // chain: the previous launch on the stream was a 64-row-block level of THIS pass (factors and inverses are read-only
    // from there on), so the next such level may be a programmatic dependent launch
    bool chain = false;
    auto link = [&](int nf, int nbg) {
        const bool p = pdl && chain && dev.dinv64 && nf > 0;
        if (nf > 0) chain = dev.dinv64 != nullptr;
        if (nbg > 0) chain = false;
        return p;
    };
    for (int l = 0; l < S.nlevels; ++l) {
        const int nf = level_nsmall[l], nbg = level_nbig[l];
        const int *nodes = level_nodes + level_off[l], *big = nodes + nf;
        const bool p = link(nf, nbg);
        if (variant == 1) {
            launch_level_fwd<true, 1>(dev, nodes, nf, level_max_m[l], d_x, ldx, s, p);
            launch_big_fwd<true>(dev, big, nbg, level_big_max_m[l], 1, d_x, ldx, s);
        } else if (variant == 2) {
            launch_level_fwd<false, 4>(dev, nodes, nf, level_max_m[l], d_x, ldx, s, p);
            launch_big_fwd<false>(dev, big, nbg, level_big_max_m[l], 4, d_x, ldx, s);
        } else {
            launch_level_fwd<false, 1>(dev, nodes, nf, level_max_m[l], d_x, ldx, s, p);
            launch_big_fwd<false>(dev, big, nbg, level_big_max_m[l], 1, d_x, ldx, s);
        }
    }
    for (int l = S.nlevels - 1; l >= 0; --l) {
        const int nf = level_nsmall[l], nbg = level_nbig[l];
        const int *nodes = level_nodes + level_off[l], *big = nodes + nf;
        if (variant == 1) {
            launch_level_bwd<true, 1>(dev, nodes, nf, level_max_m[l], level_max_np[l], d_x, ldx, s, link(nf, nbg));
            launch_big_bwd<true>(dev, big, nbg, level_big_max_m[l], 1, d_x, ldx, s);
        } else if (variant == 2) {
            // four right-hand sides at once unless their shared-memory footprint (y + per-warp partial sums) is too large
            if (sizeof(double) * 4 * ((size_t)level_max_m[l] + (size_t)(TS / 32) * level_max_np[l]) <= 200 * 1024)
                launch_level_bwd<false, 4>(dev, nodes, nf, level_max_m[l], level_max_np[l], d_x, ldx, s, link(nf, nbg));
            else {
                for (int r = 0; r < 4; ++r)
                    launch_level_bwd<false, 1>(dev, nodes, nf, level_max_m[l], level_max_np[l], d_x + (size_t)r * ldx, ldx, s,
                                               link(nf, 0));
                link(0, nbg);
            }
            launch_big_bwd<false>(dev, big, nbg, level_big_max_m[l], 4, d_x, ldx, s);
        } else {
            launch_level_bwd<false, 1>(dev, nodes, nf, level_max_m[l], level_max_np[l], d_x, ldx, s, link(nf, nbg));
            launch_big_bwd<false>(dev, big, nbg, level_big_max_m[l], 1, d_x, ldx, s);
        }
    }
    return true;
}

bool MultifrontalLU::solve4(double *d_x, cudaStream_t s, std::string &err) {
    if (!impl_) {
        err = "MultifrontalLU::solve4 before configure";
        return false;
    }
    Impl &I = *impl_;
    if (!check(err)) return false;
    if (!I.F) {
        err = "MultifrontalLU::solve4 before factor";
        return false;
    }
    g_launch_count.fetch_add(2 * I.S.nlevels, std::memory_order_relaxed);
    if (!I.run(I.solve4_graphs, d_x, s, err, [&](cudaStream_t q) { return I.enqueue_solve(d_x, 2, q, err); })) return false;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("multifrontal solve4: ") + cudaGetErrorString(e);
        return false;
    }
    return true;
}

bool MultifrontalLU::solve(double *d_x, cudaStream_t s, std::string &err, bool transposed) {
    if (!impl_) {
        err = "MultifrontalLU::solve before configure";
        return false;
    }
    Impl &I = *impl_;
    if (!check(err)) return false;
    if (!I.F) {
        err = "MultifrontalLU::solve before factor";
        return false;
    }
    g_launch_count.fetch_add(2 * I.S.nlevels, std::memory_order_relaxed);
    auto &cache = transposed ? I.solve_t_graphs : I.solve_graphs;
    if (!I.run(cache, d_x, s, err, [&](cudaStream_t q) { return I.enqueue_solve(d_x, transposed ? 1 : 0, q, err); })) return false;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("multifrontal solve: ") + cudaGetErrorString(e);
        return false;
    }
    return true;
}

}  // namespace ocp
