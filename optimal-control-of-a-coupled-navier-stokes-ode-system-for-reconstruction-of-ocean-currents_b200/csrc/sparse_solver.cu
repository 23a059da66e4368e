#include "sparse_solver.cuh"

#include <chrono>
#include <cstdio>

namespace ocp {

namespace {
const char *rf_status(cusolverStatus_t st) {
    switch (st) {
        case CUSOLVER_STATUS_SUCCESS: return "success";
        case CUSOLVER_STATUS_NOT_INITIALIZED: return "not initialized";
        case CUSOLVER_STATUS_ALLOC_FAILED: return "alloc failed";
        case CUSOLVER_STATUS_INVALID_VALUE: return "invalid value";
        case CUSOLVER_STATUS_ARCH_MISMATCH: return "arch mismatch";
        case CUSOLVER_STATUS_EXECUTION_FAILED: return "execution failed";
        case CUSOLVER_STATUS_INTERNAL_ERROR: return "internal error";
        case CUSOLVER_STATUS_ZERO_PIVOT: return "zero pivot";
        default: return "other";
    }
}
}  // namespace

#define RF_CHECK(call)                                                                      \
    do {                                                                                    \
        cusolverStatus_t st_ = (call);                                                      \
        if (st_ != CUSOLVER_STATUS_SUCCESS) {                                               \
            err = std::string(#call) + ": " + rf_status(st_) + " (" + std::to_string((int)st_) + ")"; \
            return false;                                                                   \
        }                                                                                   \
    } while (0)

#define CU_CHECK(call)                                                     \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) {                                           \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);      \
            return false;                                                  \
        }                                                                  \
    } while (0)

SparseLU::~SparseLU() {
    if (rf_) cusolverRfDestroy(rf_);
    cudaFree(d_P_);
    cudaFree(d_Q_);
    cudaFree(d_T_);
}

void SparseLU::configure(int n, int nnz, const int *h_rowptr, const int *h_col, const int *d_rowptr,
                         const int *d_col, const double *xy, const unsigned char *kind) {
    n_ = n;
    nnz_ = nnz;
    h_rowptr_.assign(h_rowptr, h_rowptr + n + 1);
    h_col_.assign(h_col, h_col + nnz);
    xy_.assign(xy, xy + 2 * (size_t)n);
    kind_.assign(kind, kind + n);
    d_rowptr_ = d_rowptr;
    d_col_ = d_col;
}

bool SparseLU::factor(const double *d_vals, cudaStream_t s, std::string &err) {
    if (!rf_) {
        // ---- one-time analysis on the host with the first matrix' values
        auto t0 = std::chrono::steady_clock::now();
        std::vector<double> h_vals(nnz_);
        CU_CHECK(cudaMemcpyAsync(h_vals.data(), d_vals, sizeof(double) * nnz_, cudaMemcpyDeviceToHost, s));
        CU_CHECK(cudaStreamSynchronize(s));
        std::vector<int> q;
        nested_dissection_order(n_, h_rowptr_.data(), h_col_.data(), xy_.data(), kind_.data(), q);
        HostLU lu;
        if (!sparse_lu(n_, h_rowptr_.data(), h_col_.data(), h_vals.data(), q, 1.0e-3, lu)) {
            err = "host LU analysis: numerically singular matrix";
            return false;
        }
        nnz_lu_ = (long long)lu.Lx.size() + (long long)lu.Ux.size();
        RF_CHECK(cusolverRfCreate(&rf_));
        RF_CHECK(cusolverRfSetNumericProperties(rf_, 0.0, 0.0));
        RF_CHECK(cusolverRfSetMatrixFormat(rf_, CUSOLVERRF_MATRIX_FORMAT_CSR, CUSOLVERRF_UNIT_DIAGONAL_STORED_L));
        RF_CHECK(cusolverRfSetResetValuesFastMode(rf_, CUSOLVERRF_RESET_VALUES_FAST_MODE_ON));
        RF_CHECK(cusolverRfSetAlgs(rf_, CUSOLVERRF_FACTORIZATION_ALG0, CUSOLVERRF_TRIANGULAR_SOLVE_ALG1));
        RF_CHECK(cusolverRfSetupHost(n_, nnz_, h_rowptr_.data(), h_col_.data(), h_vals.data(), (int)lu.Lx.size(),
                                     lu.Lp.data(), lu.Li.data(), lu.Lx.data(), (int)lu.Ux.size(), lu.Up.data(),
                                     lu.Ui.data(), lu.Ux.data(), lu.P.data(), lu.Q.data(), rf_));
        CU_CHECK(cudaDeviceSynchronize());
        RF_CHECK(cusolverRfAnalyze(rf_));
        CU_CHECK(cudaMalloc(&d_P_, sizeof(int) * n_));
        CU_CHECK(cudaMalloc(&d_Q_, sizeof(int) * n_));
        CU_CHECK(cudaMalloc(&d_T_, sizeof(double) * n_));
        CU_CHECK(cudaMemcpy(d_P_, lu.P.data(), sizeof(int) * n_, cudaMemcpyHostToDevice));
        CU_CHECK(cudaMemcpy(d_Q_, lu.Q.data(), sizeof(int) * n_, cudaMemcpyHostToDevice));
        analyse_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    // cusolverRf has no stream setter: it runs on the legacy default stream, which orders itself against
    // any blocking stream; a non-blocking caller stream is synchronised explicitly.
    if (s != nullptr) CU_CHECK(cudaStreamSynchronize(s));
    RF_CHECK(cusolverRfResetValues(n_, nnz_, const_cast<int *>(d_rowptr_), const_cast<int *>(d_col_),
                                   const_cast<double *>(d_vals), d_P_, d_Q_, rf_));
    RF_CHECK(cusolverRfRefactor(rf_));
    if (s != nullptr) CU_CHECK(cudaStreamSynchronize(nullptr));
    return true;
}

bool SparseLU::solve(double *d_x, cudaStream_t s, std::string &err) {
    if (!rf_) {
        err = "SparseLU::solve before factor";
        return false;
    }
    if (s != nullptr) CU_CHECK(cudaStreamSynchronize(s));
    RF_CHECK(cusolverRfSolve(rf_, d_P_, d_Q_, 1, d_T_, n_, d_x, n_));
    if (s != nullptr) CU_CHECK(cudaStreamSynchronize(nullptr));
    return true;
}

}  // namespace ocp
