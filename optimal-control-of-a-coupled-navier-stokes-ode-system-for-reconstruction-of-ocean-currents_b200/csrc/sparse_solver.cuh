// GPU sparse direct solver with a one-time host analysis (host_lu.hpp) whose pattern, pivot sequence and
// column order are reused by every later numeric refactorisation on the device (cusolverRf).
// Replaces dolfin's LU solves at OCP_dolfin.py:325 (inside Newton), 329 (projection) and 371 (adjoint).
// There is no CPU solve path: after the first analysis every factor/solve runs on the GPU.
#pragma once
#include <cuda_runtime.h>
#include <cusolverRf.h>

#include <string>
#include <vector>

#include "host_lu.hpp"

namespace ocp {

class SparseLU {
  public:
    SparseLU() = default;
    ~SparseLU();
    SparseLU(const SparseLU &) = delete;
    SparseLU &operator=(const SparseLU &) = delete;

    // pattern (host + device copies), dof coordinates and kinds for the ordering
    void configure(int n, int nnz, const int *h_rowptr, const int *h_col, const int *d_rowptr, const int *d_col,
                   const double *xy, const unsigned char *kind);
    // numeric (re)factorisation of the matrix whose CSR values are d_vals; first call runs the host analysis
    bool factor(const double *d_vals, cudaStream_t s, std::string &err);
    // d_x: in = right-hand side, out = solution
    bool solve(double *d_x, cudaStream_t s, std::string &err);
    bool analysed() const { return rf_ != nullptr; }
    long long factor_nnz() const { return nnz_lu_; }
    double analyse_ms = 0.0;

  private:
    int n_ = 0, nnz_ = 0;
    std::vector<int> h_rowptr_, h_col_;
    std::vector<double> xy_;
    std::vector<unsigned char> kind_;
    const int *d_rowptr_ = nullptr, *d_col_ = nullptr;
    cusolverRfHandle_t rf_ = nullptr;
    int *d_P_ = nullptr, *d_Q_ = nullptr;
    double *d_T_ = nullptr;
    long long nnz_lu_ = 0;
};

}  // namespace ocp
