// Element-level arithmetic shared by the kernels (device) and the host-side self-test entry points.
// Pure functions, no memory traffic, no atomics.
//
//   cell_row     one row of the 15x15 Taylor-Hood element matrix dF/dw and of the residual F
//                (UFL forms at OCP_dolfin.py:321-323; adjoint form OCP_dolfin.py:344-347 is its transpose at nu=1)
//   facet_row    one row of the 6x6 Gamma_1 facet block  -1/2 (u.n)(u.v) ds(1) - f.v ds(1)
//   bary / p2    point location predicate and P2 basis with the fixed operation order that the CPU oracle
//                (oracle/buoy_oracle.c) restates, so that cell indices and trajectories are bit-identical
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDA_ARCH__)
#define OCP_FMA(a, b, c) __fma_rn((a), (b), (c))
#define OCP_MUL(a, b) __dmul_rn((a), (b))
#define OCP_ADD(a, b) __dadd_rn((a), (b))
#define OCP_SUB(a, b) __dsub_rn((a), (b))
#else
#define OCP_FMA(a, b, c) fma((a), (b), (c))
#define OCP_MUL(a, b) ((a) * (b))
#define OCP_ADD(a, b) ((a) + (b))
#define OCP_SUB(a, b) ((a) - (b))
#endif

#define OCP_HD __host__ __device__ __forceinline__

namespace ocp {

constexpr double kLocateTol = 1.0e-14;
constexpr double kLocateMargin = 1.0e-9;

// ---- point location / basis (bit-exact contract) -----------------------------------------------------------
OCP_HD void bary(const double g[6], double x, double y, double &l0, double &l1, double &l2) {
    const double dx = OCP_SUB(x, g[0]), dy = OCP_SUB(y, g[1]);
    l1 = OCP_FMA(g[2], dx, OCP_MUL(g[3], dy));
    l2 = OCP_FMA(g[4], dx, OCP_MUL(g[5], dy));
    l0 = OCP_SUB(OCP_SUB(1.0, l1), l2);
}

OCP_HD void p2_basis(double l0, double l1, double l2, double phi[6]) {
    phi[0] = OCP_MUL(l0, OCP_SUB(OCP_MUL(2.0, l0), 1.0));
    phi[1] = OCP_MUL(l1, OCP_SUB(OCP_MUL(2.0, l1), 1.0));
    phi[2] = OCP_MUL(l2, OCP_SUB(OCP_MUL(2.0, l2), 1.0));
    phi[3] = OCP_MUL(OCP_MUL(4.0, l1), l2);
    phi[4] = OCP_MUL(OCP_MUL(4.0, l0), l2);
    phi[5] = OCP_MUL(OCP_MUL(4.0, l0), l1);
}

// ---- 7-point degree-5 triangle rule (Radon) -----------------------------------------------------------------
// centroid 9/40; a = (6 -+ sqrt(15))/21 with weights (155 -+ sqrt(15))/1200
#define OCP_Q_A1 0.10128650732345633880
#define OCP_Q_B1 0.79742698535308732240
#define OCP_Q_W1 0.12593918054482715260
#define OCP_Q_A2 0.47014206410511508977
#define OCP_Q_B2 0.05971587178976982045
#define OCP_Q_W2 0.13239415278850618074

OCP_HD void tri_qpoint(int q, double &l0, double &l1, double &l2, double &w) {
    switch (q) {
        case 0: l0 = l1 = l2 = 1.0 / 3.0; w = 0.225; break;
        case 1: l0 = OCP_Q_B1; l1 = OCP_Q_A1; l2 = OCP_Q_A1; w = OCP_Q_W1; break;
        case 2: l0 = OCP_Q_A1; l1 = OCP_Q_B1; l2 = OCP_Q_A1; w = OCP_Q_W1; break;
        case 3: l0 = OCP_Q_A1; l1 = OCP_Q_A1; l2 = OCP_Q_B1; w = OCP_Q_W1; break;
        case 4: l0 = OCP_Q_B2; l1 = OCP_Q_A2; l2 = OCP_Q_A2; w = OCP_Q_W2; break;
        case 5: l0 = OCP_Q_A2; l1 = OCP_Q_B2; l2 = OCP_Q_A2; w = OCP_Q_W2; break;
        default: l0 = OCP_Q_A2; l1 = OCP_Q_A2; l2 = OCP_Q_B2; w = OCP_Q_W2; break;
    }
}

// Row `row` (0..5 u_x, 6..11 u_y, 12..14 p) of the element Jacobian (A[15]) and residual (R) of the cell with
// geometry g = [x0 y0 a1 b1 a2 b2] at the state (U,V,P).
OCP_HD void cell_row(const double g[6], const double U[6], const double V[6], const double P[3], double nu,
                     int row, double A[15], double &R) {
    const double g1x = g[2], g1y = g[3], g2x = g[4], g2y = g[5];
    const double g0x = -(g1x + g2x), g0y = -(g1y + g2y);
    const double area = 0.5 / fabs(g1x * g2y - g2x * g1y);
#pragma unroll
    for (int j = 0; j < 15; ++j) A[j] = 0.0;
    R = 0.0;
    const int a = row < 6 ? row : (row < 12 ? row - 6 : row - 12);
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        double l0, l1, l2, wq;
        tri_qpoint(q, l0, l1, l2, wq);
        const double w = wq * area;
        double phi[6], gx[6], gy[6];
        phi[0] = l0 * (2.0 * l0 - 1.0);
        phi[1] = l1 * (2.0 * l1 - 1.0);
        phi[2] = l2 * (2.0 * l2 - 1.0);
        phi[3] = 4.0 * l1 * l2;
        phi[4] = 4.0 * l0 * l2;
        phi[5] = 4.0 * l0 * l1;
        const double d0 = 4.0 * l0 - 1.0, d1 = 4.0 * l1 - 1.0, d2 = 4.0 * l2 - 1.0;
        gx[0] = d0 * g0x; gy[0] = d0 * g0y;
        gx[1] = d1 * g1x; gy[1] = d1 * g1y;
        gx[2] = d2 * g2x; gy[2] = d2 * g2y;
        gx[3] = 4.0 * (l2 * g1x + l1 * g2x); gy[3] = 4.0 * (l2 * g1y + l1 * g2y);
        gx[4] = 4.0 * (l2 * g0x + l0 * g2x); gy[4] = 4.0 * (l2 * g0y + l0 * g2y);
        gx[5] = 4.0 * (l1 * g0x + l0 * g1x); gy[5] = 4.0 * (l1 * g0y + l0 * g1y);
        double ux = 0, uy = 0, uxx = 0, uxy = 0, uyx = 0, uyy = 0;
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            ux += U[b] * phi[b];
            uy += V[b] * phi[b];
            uxx += U[b] * gx[b];
            uxy += U[b] * gy[b];
            uyx += V[b] * gx[b];
            uyy += V[b] * gy[b];
        }
        const double lam[3] = {l0, l1, l2};
        if (row < 12) {
            const double pa = phi[a], gxa = gx[a], gya = gy[a];
            const bool isx = row < 6;
            const double dself = isx ? uxx : uyy;      // d u_c / d x_c
            const double dother = isx ? uxy : uyx;     // coupling to the other component
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                const double adv = ux * gx[b] + uy * gy[b];
                const double same = w * (nu * (gxa * gx[b] + gya * gy[b]) + (adv + dself * phi[b]) * pa);
                const double cross = w * (dother * phi[b] * pa);
                if (isx) {
                    A[b] += same;
                    A[6 + b] += cross;
                } else {
                    A[b] += cross;
                    A[6 + b] += same;
                }
            }
            const double gca = isx ? gxa : gya;
#pragma unroll
            for (int j = 0; j < 3; ++j) A[12 + j] += w * lam[j] * gca;
            const double p = P[0] * l0 + P[1] * l1 + P[2] * l2;
            if (isx)
                R += w * (nu * (uxx * gxa + uxy * gya) + (ux * uxx + uy * uxy) * pa + p * gxa);
            else
                R += w * (nu * (uyx * gxa + uyy * gya) + (ux * uyx + uy * uyy) * pa + p * gya);
        } else {
            const double li = lam[a];
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                A[b] += w * gx[b] * li;
                A[6 + b] += w * gy[b] * li;
            }
            R += w * (uxx + uyy) * li;
        }
    }
}

// ---- Gamma_1 facet block, 4-point Gauss-Legendre on [0,1] -----------------------------------------------------
#define OCP_GL4_X0 0.06943184420297371239
#define OCP_GL4_X1 0.33000947820757186760
#define OCP_GL4_W0 0.17392742256872692869
#define OCP_GL4_W1 0.32607257743127307131

// Row r = 3*c + a of the facet block; trial columns [u_x(va,vb,mid), u_y(va,vb,mid)].
// U,V: velocity at the three facet nodes, F: control (x,y) at the three nodes.
OCP_HD void facet_row(double len, double nx, double ny, const double U[3], const double V[3], const double Fx[3],
                      const double Fy[3], int r, double A[6], double &R) {
    const int c = r / 3, a = r % 3;
#pragma unroll
    for (int j = 0; j < 6; ++j) A[j] = 0.0;
    R = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double s = (q == 0) ? OCP_GL4_X0 : (q == 1) ? OCP_GL4_X1 : (q == 2) ? 1.0 - OCP_GL4_X1 : 1.0 - OCP_GL4_X0;
        const double w = len * ((q == 0 || q == 3) ? OCP_GL4_W0 : OCP_GL4_W1);
        const double N[3] = {(1.0 - s) * (1.0 - 2.0 * s), s * (2.0 * s - 1.0), 4.0 * s * (1.0 - s)};
        const double ux = U[0] * N[0] + U[1] * N[1] + U[2] * N[2];
        const double uy = V[0] * N[0] + V[1] * N[1] + V[2] * N[2];
        const double fx = Fx[0] * N[0] + Fx[1] * N[1] + Fx[2] * N[2];
        const double fy = Fy[0] * N[0] + Fy[1] * N[1] + Fy[2] * N[2];
        const double un = ux * nx + uy * ny;
        const double Na = N[a];
        if (c == 0) {
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                A[b] += -0.5 * w * (nx * ux + un) * N[b] * Na;
                A[3 + b] += -0.5 * w * (ny * ux) * N[b] * Na;
            }
            R += w * (-0.5 * un * ux - fx) * Na;
        } else {
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                A[b] += -0.5 * w * (nx * uy) * N[b] * Na;
                A[3 + b] += -0.5 * w * (ny * uy + un) * N[b] * Na;
            }
            R += w * (-0.5 * un * uy - fy) * Na;
        }
    }
}

}  // namespace ocp
