// One-time HOST analysis for the sparse direct solver: geometric nested-dissection column
// order + left-looking LU with threshold partial pivoting.  Its output (pattern of L and U,
// row pivot sequence, column order) is what the GPU refactorisation reuses for every Newton
// step and adjoint solve; the numeric factors it computes seed the first device solve.
//
// Replaces the analysis phase of dolfin's default LU (`solve(F == 0, w, bcs)` OCP_dolfin.py:325,
// `solve(A, zrSol.vector(), b)` OCP_dolfin.py:371 -> PETSc/UMFPACK, not in /root/reference).
#pragma once
#include <cstdint>
#include <vector>

namespace ocp {

struct HostLU {
    int n = 0;
    std::vector<int> P, Q;            // A[P[i], Q[j]] = (L U)[i, j]
    std::vector<int> Lp, Li;          // CSR, unit diagonal stored, columns ascending
    std::vector<double> Lx;
    std::vector<int> Up, Ui;          // CSR, diagonal first is NOT assumed; columns ascending
    std::vector<double> Ux;
    double min_pivot = 0.0, max_pivot = 0.0;
};

// q: fill-reducing column order from recursive coordinate bisection of the dof coordinates.
// `kind[i]` (0 velocity, 1 pressure) orders velocities before pressures inside leaves/separators
// so that pressure pivots are Schur complements rather than structural zeros.
void nested_dissection_order(int n, const int *rowptr, const int *col, const double *xy, const uint8_t *kind,
                             std::vector<int> &q);

// Returns false on a numerically singular column.
bool sparse_lu(int n, const int *rowptr, const int *col, const double *val, const std::vector<int> &q,
               double pivot_threshold, HostLU &out);

// x <- A^{-1} x using the host factors (tests, and the first solve's cross-check)
void host_lu_solve(const HostLU &lu, double *x);

}  // namespace ocp
