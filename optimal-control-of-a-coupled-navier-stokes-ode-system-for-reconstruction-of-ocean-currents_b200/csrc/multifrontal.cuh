// Device-side multifrontal LU (numeric factorisation + triangular solves on the GPU); see multifrontal.hpp for the
// host symbolic analysis it consumes.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "multifrontal.hpp"

namespace ocp {

class MultifrontalLU {
  public:
    MultifrontalLU();
    ~MultifrontalLU();
    MultifrontalLU(const MultifrontalLU &) = delete;
    MultifrontalLU &operator=(const MultifrontalLU &) = delete;

    // symbolic analysis (host, pattern + coordinates only) and upload of the schedule
    // `share`: another solver already configured for the SAME pattern, coordinates and kinds - its symbolic analysis
    // is copied instead of being recomputed (the forward, adjoint and Stokes matrices of a context share one pattern)
    bool configure(int n, int nnz, const int *h_rowptr, const int *h_col, const double *xy, const unsigned char *kind,
                   std::string &err, const MultifrontalLU *share = nullptr);
    bool factor(const double *d_vals, cudaStream_t s, std::string &err);   // numeric factorisation on the GPU
    // in place; transposed: solves A^T x = b with the same factors
    bool solve(double *d_x, cudaStream_t s, std::string &err, bool transposed = false);
    // four right-hand sides at once: vector r occupies d_x[r * n .. r * n + n)
    bool solve4(double *d_x, cudaStream_t s, std::string &err);
    bool check(std::string &err);   // zero-pivot flag of completed factorisations (non-blocking)
    // while set, factor / solve enqueue their kernels directly on the given stream instead of launching their own
    // CUDA graphs: the caller is capturing a larger graph around them
    void set_direct_enqueue(bool on);
    bool needs_cooperative_launch() const;   // large fronts present (group kernels): not captured into outer graphs
    long long factor_nnz() const { return factor_nnz_; }
    double flops() const { return flops_; }
    int levels() const { return nlevels_; }
    int max_front() const { return max_front_; }
    int fronts() const { return nfronts_; }
    long long workspace_doubles() const { return fsize_; }
    double analyse_ms = 0.0;

  private:
    struct Impl;
    Impl *impl_ = nullptr;
    int n_ = 0, nnz_ = 0, nlevels_ = 0, max_front_ = 0;
    long long factor_nnz_ = 0, fsize_ = 0;
    int nfronts_ = 0;
    double flops_ = 0.0;
};

}  // namespace ocp
