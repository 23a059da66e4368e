/*
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's three per-buoy loops.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library; the product path (libocp_b200.so) never does.
 *
 *   oracle_buoy_forward   <- solve_primal_ode          OCP_dolfin.py:201-230
 *   oracle_buoy_adjoint   <- solve_adjoint_ode         OCP_dolfin.py:234-252
 *   oracle_point_sources  <- the PointSource loop      OCP_dolfin.py:353-366
 *   oracle_misfit         <- partA of J                OCP_dolfin.py:259
 *
 * dolfin's `Function.__call__` (BB-tree point location + P2 `evaluate_basis`) and
 * `PointSource.apply` are not in /root/reference (legacy FEniCS, un-vendored,
 * un-pinned).  Their published semantics are restated here:
 *   - first colliding cell, closed containment  -> lowest cell index whose
 *     barycentric coordinates are all >= -1e-14 (tie rule of SURVEY App. C / K1);
 *   - no colliding cell -> RuntimeError -> the `except` branches of the reference;
 *   - UFC P2 basis  l_i(2 l_i - 1), 4 l_1 l_2, 4 l_0 l_2, 4 l_0 l_1.
 * Pinned by tests/test_oracle_golden.py against reference_runs/{2,4,6,10,100,400}_buoys.
 *
 * Arithmetic contract (shared with the CUDA kernels so that cell indices and
 * trajectories are bit-identical; compile with -ffp-contract=off):
 *   dx = x - x0; dy = y - y0
 *   l1 = fma(a1, dx, b1*dy);  l2 = fma(a2, dx, b2*dy);  l0 = (1 - l1) - l2
 *   phi_i = l_i * (2 l_i - 1), phi_3 = (4 l1) l2, phi_4 = (4 l0) l2, phi_5 = (4 l0) l1
 *   u = phi_0 c_0, then u = fma(phi_i, c_i, u) for i = 1..5
 *   x_{k+1} = x_k + (h * u)      two roundings, as numpy does at OCP_dolfin.py:212
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* The reference loops are serial (no MPI/threads anywhere in the pipelines).  For the timed CPU baseline the
 * outer per-buoy loops may additionally be spread over host threads (buoys are independent); nthreads = 1
 * keeps the reference's loop structure. */
static int g_threads = 1;
void oracle_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#define LOCATE_TOL 1.0e-14
#define LOCATE_MARGIN 1.0e-9

typedef struct {
    int32_t nc, nn, nv;
    const double *geom;        /* (nc,6) x0 y0 a1 b1 a2 b2 */
    const int32_t *cell_nodes; /* (nc,6) */
    double ox, oy, ihx, ihy;   /* bins */
    int32_t nbx, nby;
    const int32_t *bin_ptr, *bin_cells;
    int32_t brute;             /* 1: scan all cells in ascending order (definition), 0: bins + previous-cell fast path */
} Tables;

static inline void bary(const Tables *t, int c, double x, double y, double *l0, double *l1, double *l2) {
    const double *g = t->geom + 6 * (size_t)c;
    double dx = x - g[0], dy = y - g[1];
    *l1 = fma(g[2], dx, g[3] * dy);
    *l2 = fma(g[4], dx, g[5] * dy);
    *l0 = (1.0 - *l1) - *l2;
}

/* returns cell index or -1; hint = previous cell (or -1) */
static int locate(const Tables *t, double x, double y, int hint, double *l0, double *l1, double *l2) {
    if (!(x == x) || !(y == y)) return -1;
    if (t->brute) {
        for (int c = 0; c < t->nc; ++c) {
            bary(t, c, x, y, l0, l1, l2);
            if (*l0 >= -LOCATE_TOL && *l1 >= -LOCATE_TOL && *l2 >= -LOCATE_TOL) return c;
        }
        return -1;
    }
    if (hint >= 0) {
        bary(t, hint, x, y, l0, l1, l2);
        if (*l0 > LOCATE_MARGIN && *l1 > LOCATE_MARGIN && *l2 > LOCATE_MARGIN) return hint;
    }
    double fx = floor((x - t->ox) * t->ihx), fy = floor((y - t->oy) * t->ihy);
    int ix = fx < 0 ? 0 : (fx > t->nbx - 1 ? t->nbx - 1 : (int)fx);
    int iy = fy < 0 ? 0 : (fy > t->nby - 1 ? t->nby - 1 : (int)fy);
    int b = iy * t->nbx + ix;
    for (int j = t->bin_ptr[b]; j < t->bin_ptr[b + 1]; ++j) {
        int c = t->bin_cells[j];
        bary(t, c, x, y, l0, l1, l2);
        if (*l0 >= -LOCATE_TOL && *l1 >= -LOCATE_TOL && *l2 >= -LOCATE_TOL) return c;
    }
    return -1;
}

static inline void p2_basis(double l0, double l1, double l2, double *phi) {
    phi[0] = l0 * (2.0 * l0 - 1.0);
    phi[1] = l1 * (2.0 * l1 - 1.0);
    phi[2] = l2 * (2.0 * l2 - 1.0);
    phi[3] = (4.0 * l1) * l2;
    phi[4] = (4.0 * l0) * l2;
    phi[5] = (4.0 * l0) * l1;
}

/* wSol.sub(0)(point): returns cell (>=0) or -1 (= the exception of the reference) */
static int eval_velocity(const Tables *t, const double *vel, double x, double y, int hint, double *ux, double *uy) {
    double l0, l1, l2, phi[6];
    int c = locate(t, x, y, hint, &l0, &l1, &l2);
    if (c < 0) return -1;
    p2_basis(l0, l1, l2, phi);
    const int32_t *n = t->cell_nodes + 6 * (size_t)c;
    double sx = phi[0] * vel[2 * n[0]], sy = phi[0] * vel[2 * n[0] + 1];
    for (int i = 1; i < 6; ++i) {
        sx = fma(phi[i], vel[2 * n[i]], sx);
        sy = fma(phi[i], vel[2 * n[i] + 1], sy);
    }
    *ux = sx;
    *uy = sy;
    return c;
}

/* solve_primal_ode, OCP_dolfin.py:201-230.  x,u: (K,nt,2) C-order, zero-initialised here.
 * cell: (K,nt) cell used for each stored velocity sample (-1 where none).
 * mask: (K) doubles, set to 1.0 like `buoy_mask[b_iter] = True` (never cleared here).
 * parked: (K) uint8, 1 when only the last point left the domain (OCP_dolfin.py:226-229). */
void oracle_buoy_forward(const Tables *t, const double *vel, int K, int nt, double h, const double *x0,
                         const double *center, double *x, double *u, int32_t *cell, double *mask, uint8_t *parked) {
#pragma omp parallel for schedule(dynamic, 64) num_threads(g_threads)
    for (int b = 0; b < K; ++b) {
        double *xb = x + (size_t)b * nt * 2, *ub = u + (size_t)b * nt * 2;
        int32_t *cb = cell + (size_t)b * nt;
        memset(xb, 0, sizeof(double) * nt * 2);
        memset(ub, 0, sizeof(double) * nt * 2);
        for (int k = 0; k < nt; ++k) cb[k] = -1;
        parked[b] = 0;
        xb[0] = x0[2 * b];
        xb[1] = x0[2 * b + 1];
        int hint = -1, k, failed = 0;
        for (k = 0; k < nt - 1; ++k) {
            double ux, uy;
            int c = eval_velocity(t, vel, xb[2 * k], xb[2 * k + 1], hint, &ux, &uy);
            if (c < 0) { /* except: park the whole trajectory, mask, break */
                for (int j = 0; j < nt; ++j) {
                    xb[2 * j] = center[0];
                    xb[2 * j + 1] = center[1];
                }
                mask[b] = 1.0;
                failed = 1;
                break;
            }
            xb[2 * k + 2] = xb[2 * k] + h * ux;
            xb[2 * k + 3] = xb[2 * k + 1] + h * uy;
            ub[2 * k] = ux;
            ub[2 * k + 1] = uy;
            cb[k] = c;
            hint = c;
        }
        /* python leaves k at the failing index after `break`, at nt-2 otherwise */
        if (!failed) k = nt - 2;
        {
            double ux, uy;
            int c = eval_velocity(t, vel, xb[2 * k + 2], xb[2 * k + 3], failed ? -1 : hint, &ux, &uy);
            if (c >= 0) {
                ub[2 * k + 2] = ux;
                ub[2 * k + 3] = uy;
                cb[k + 1] = c;
            } else {
                ub[2 * k + 2] = 0.0;
                ub[2 * k + 3] = 0.0;
                xb[2 * k + 2] = center[0];
                xb[2 * k + 3] = center[1];
                parked[b] = 1;
            }
        }
    }
}

/* grad_u(point) on the continuous P1 tensor function: g (nv,4) row-major [g00 g01 g10 g11] */
static int eval_grad(const Tables *t, const double *g, double x, double y, int hint, double *G) {
    double l[3];
    int c = locate(t, x, y, hint, &l[0], &l[1], &l[2]);
    if (c < 0) return -1;
    const int32_t *n = t->cell_nodes + 6 * (size_t)c;
    for (int j = 0; j < 4; ++j) G[j] = l[0] * g[4 * n[0] + j] + l[1] * g[4 * n[1] + j] + l[2] * g[4 * n[2] + j];
    return c;
}

/* solve_adjoint_ode, OCP_dolfin.py:234-252.  mu zero-initialised here. */
void oracle_buoy_adjoint(const Tables *t, const double *g, int K, int nt, double h, const double *x, const double *u,
                         const double *ud, const double *mask, double *mu) {
    memset(mu, 0, sizeof(double) * (size_t)K * nt * 2);
#pragma omp parallel for schedule(dynamic, 64) num_threads(g_threads)
    for (int b = 0; b < K; ++b) {
        if (mask[b] != 0.0) continue;
        const double *xb = x + (size_t)b * nt * 2, *ub = u + (size_t)b * nt * 2, *db = ud + (size_t)b * nt * 2;
        double *mb = mu + (size_t)b * nt * 2;
        double G[4] = {0, 0, 0, 0}; /* stale value survives a failed evaluation, OCP_dolfin.py:242-251 */
        int hint = -1;
        for (int k = nt - 2; k >= 0; --k) {
            int c = eval_grad(t, g, xb[2 * k + 2], xb[2 * k + 3], hint, G);
            if (c >= 0) hint = c;
            double rx = (ub[2 * k + 2] - db[2 * k + 2]) - mb[2 * k + 2];
            double ry = (ub[2 * k + 3] - db[2 * k + 3]) - mb[2 * k + 3];
            /* mu_k = mu_{k+1} - h * G^T r */
            mb[2 * k] = mb[2 * k + 2] - h * (G[0] * rx + G[2] * ry);
            mb[2 * k + 1] = mb[2 * k + 3] - h * (G[1] * rx + G[3] * ry);
        }
    }
}

/* PointSource loop, OCP_dolfin.py:353-366: bnode (nn,2) += gamma_c * phi_i(point).  bnode is NOT cleared. */
static void point_sources_range(const Tables *t, const double *vel, int b0, int b1, int nt, double h,
                                const double *x, const double *ud, const double *mu, const double *mask,
                                const double *center, double *bnode) {
    for (int b = b0; b < b1; ++b) {
        if (mask[b] != 0.0) continue;
        int hint = -1;
        for (int k = 0; k < nt; ++k) {
            size_t o = ((size_t)b * nt + k) * 2;
            double px = x[o], py = x[o + 1], ux, uy;
            int c = eval_velocity(t, vel, px, py, hint, &ux, &uy);
            if (c < 0) { /* except: u_x = 0, point = centre */
                ux = uy = 0.0;
                px = center[0];
                py = center[1];
            }
            double gx = h * ((ud[o] - ux) + mu[o]);
            double gy = h * ((ud[o + 1] - uy) + mu[o + 1]);
            double l0, l1, l2, phi[6];
            c = locate(t, px, py, c, &l0, &l1, &l2);
            if (c < 0) continue; /* PointSource outside the mesh: dolfin would raise; centre is always inside */
            hint = c;
            p2_basis(l0, l1, l2, phi);
            const int32_t *n = t->cell_nodes + 6 * (size_t)c;
            for (int i = 0; i < 6; ++i) {
                bnode[2 * n[i]] += gx * phi[i];
                bnode[2 * n[i] + 1] += gy * phi[i];
            }
        }
    }
}

void oracle_point_sources(const Tables *t, const double *vel, int K, int nt, double h, const double *x,
                          const double *ud, const double *mu, const double *mask, const double *center,
                          double *bnode) {
    if (g_threads <= 1) {
        point_sources_range(t, vel, 0, K, nt, h, x, ud, mu, mask, center, bnode);
        return;
    }
    /* threaded baseline: private accumulators, summed in thread order */
    const int nth = g_threads;
    const size_t len = 2 * (size_t)t->nn;
    double *priv = (double *)calloc(len * nth, sizeof(double));
#pragma omp parallel for schedule(static) num_threads(nth)
    for (int th = 0; th < nth; ++th) {
        int b0 = (int)((long long)K * th / nth), b1 = (int)((long long)K * (th + 1) / nth);
        point_sources_range(t, vel, b0, b1, nt, h, x, ud, mu, mask, center, priv + len * th);
    }
    for (int th = 0; th < nth; ++th)
        for (size_t i = 0; i < len; ++i) bnode[i] += priv[len * th + i];
    free(priv);
}

/* partA of J, OCP_dolfin.py:259 (all buoys, masked ones included) */
double oracle_misfit(int K, int nt, double h, const double *u, const double *ud) {
    double s = 0.0;
    for (int b = 0; b < K; ++b) {
        double sb = 0.0;
        for (int k = 0; k < nt; ++k) {
            size_t o = ((size_t)b * nt + k) * 2;
            double ex = u[o] - ud[o], ey = u[o + 1] - ud[o + 1];
            sb += h * (ex * ex + ey * ey);
        }
        s += sb;
    }
    return 0.5 * s;
}
