// Multifrontal sparse LU for the Taylor-Hood saddle-point systems of the path (Newton matrices, adjoint
// matrix, P1 mass matrix).  The SYMBOLIC analysis (geometric nested-dissection tree, front index sets,
// assembly and extend-add maps, level schedule) is done once on the host from the pattern and the dof
// coordinates alone - no numeric values are needed - and is reused by every numeric factorisation, which
// runs entirely on the GPU (multifrontal.cu): one CTA per front, one launch per tree level.
//
// Pivoting is restricted to the fully-summed rows of each front, so the structure never changes; velocities
// are ordered before pressures inside every front so that pressure pivots are Schur complements.
//
// Replaces dolfin's LU (PETSc/UMFPACK) at OCP_dolfin.py:325, 329, 371.
#pragma once
#include <cstdint>
#include <vector>

namespace ocp {

struct MFSymbolic {
    int n = 0, nnodes = 0, nlevels = 0, max_front = 0, max_np = 0;
    long long fsize = 0;                 // doubles in the front workspace
    std::vector<int> perm;               // elimination position -> dof
    std::vector<int> first, np, m;       // per node: first pivot position, #pivots, front order
    std::vector<int> parent;
    std::vector<int> idx_ptr, idx;       // per node: dof ids of the front (pivots, then update set)
    std::vector<long long> front_ptr;    // per node: offset of its column-major m x m front
    std::vector<int> child_ptr, child;   // children lists
    std::vector<int> rel_ptr, rel;       // per node: position of each update row in the PARENT's front
    std::vector<int> level_ptr, level_nodes;
    std::vector<long long> a_dest;       // per CSR non-zero of A: destination offset in the front workspace
    double flops = 0.0;
};

// largest dof set that is not dissected further (a leaf front): 80, or environment OCP_MF_LEAF
int mf_leaf_size();

// kind: 0 velocity-like, 1 pressure-like (ordered last inside a front); leaf: max dofs of a leaf front
void mf_analyse(int n, const int *rowptr, const int *col, const double *xy, const uint8_t *kind, int leaf,
                MFSymbolic &S);

// Host restatement of the numeric phase on the same data structures (unit tests of the symbolic analysis; the
// product factor/solve is the CUDA implementation).
struct MFHostNumeric {
    std::vector<double> F;
    std::vector<int> piv;                // per elimination position: local pivot row chosen inside its front
    double min_pivot = 0.0;
};
extern int g_mf_host_window;   // pivot-search window of mf_factor_host (1 = static pivoting like the device)
bool mf_factor_host(const MFSymbolic &S, const double *vals, MFHostNumeric &N);
void mf_solve_host(const MFSymbolic &S, const MFHostNumeric &N, double *x);

}  // namespace ocp
