// Microbenchmark (sm_100a): cycles of the 16 x 16 diagonal-block LU (one warp, lane = row) in the variants tried for
// the multifrontal panel kernel, and of the "rows below the block" substitution L = A U11^{-1} (one row per thread).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o diag16 diag16.cu && ./diag16
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NB = 16;
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

template <int V>
__global__ void diag(const double *A, double *out, long long *cyc) {
    __shared__ double s_D[NB][NB + 1];
    __shared__ double s_rd[NB];
    __shared__ __align__(16) double s_prow[2][NB];
    const int tid = threadIdx.x;
    double r[NB];
#pragma unroll
    for (int jj = 0; jj < NB; ++jj) r[jj] = tid < NB ? A[tid * NB + jj] : (jj == tid ? 1.0 : 0.0);
    __syncthreads();
    long long t0 = clock64();
    if (tid < 32) {
        if (V == 0) {          // shuffles
            double rinv = fast_rcp(__shfl_sync(0xffffffffu, r[0], 0));
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                if (tid == j) s_rd[j] = rinv;
                const bool below = tid > j;
                const double l = r[j] * rinv;
                if (below) r[j] = l;
                double rnext = 0.0;
                if (j + 1 < NB) {
                    const double u1 = __shfl_sync(0xffffffffu, r[j + 1], j);
                    if (below) r[j + 1] = fma(-l, u1, r[j + 1]);
                    rnext = fast_rcp(__shfl_sync(0xffffffffu, r[j + 1], j + 1));
                }
#pragma unroll
                for (int jj = 0; jj < NB; ++jj)
                    if (jj > j + 1) {
                        const double u = __shfl_sync(0xffffffffu, r[jj], j);
                        if (below) r[jj] = fma(-l, u, r[jj]);
                    }
                rinv = rnext;
            }
        } else if (V == 1) {   // pivot row through shared memory
            double rinv = fast_rcp(__shfl_sync(0xffffffffu, r[0], 0));
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                double *prow = s_prow[j & 1];
                if (tid == j) {
                    s_rd[j] = rinv;
#pragma unroll
                    for (int jj = 0; jj < NB; jj += 2)
                        if (jj + 1 > j) *reinterpret_cast<double2 *>(prow + jj) = make_double2(r[jj], r[jj + 1]);
                }
                __syncwarp();
                const bool below = tid > j;
                const double l = r[j] * rinv;
                if (below) r[j] = l;
                double rnext = 0.0;
                if (j + 1 < NB) {
                    const double u1 = prow[j + 1];
                    if (below) r[j + 1] = fma(-l, u1, r[j + 1]);
                    rnext = fast_rcp(__shfl_sync(0xffffffffu, r[j + 1], j + 1));
                }
#pragma unroll
                for (int jj = 0; jj < NB; ++jj)
                    if (jj > j + 1) {
                        const double u = prow[jj];
                        if (below) r[jj] = fma(-l, u, r[jj]);
                    }
                rinv = rnext;
            }
        } else {               // all of the block in shared memory, one thread per ROW, no register block
            for (int jj = 0; jj < NB; ++jj)
                if (tid < NB) s_D[tid][jj] = r[jj];
            __syncwarp();
            for (int j = 0; j < NB; ++j) {
                const double rinv = fast_rcp(s_D[j][j]);
                if (tid == j) s_rd[j] = rinv;
                if (tid > j && tid < NB) {
                    const double l = s_D[tid][j] * rinv;
                    s_D[tid][j] = l;
                    for (int jj = j + 1; jj < NB; ++jj) s_D[tid][jj] = fma(-l, s_D[j][jj], s_D[tid][jj]);
                }
                __syncwarp();
            }
#pragma unroll
            for (int jj = 0; jj < NB; ++jj) r[jj] = tid < NB ? s_D[tid][jj] : 0.0;
        }
    }
    long long t1 = clock64();
    if (tid < NB) {
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) s_D[tid][jj] = r[jj];
    }
    __syncthreads();
    long long t2 = clock64();
    // rows below: L = A U11^{-1}, one row per thread
    double a[NB];
#pragma unroll
    for (int jj = 0; jj < NB; ++jj) a[jj] = A[(tid % NB) * NB + jj] + 1e-3 * tid;
    long long t3 = clock64();
#pragma unroll
    for (int t = 0; t < NB; ++t) {
        const double l = a[t] * s_rd[t];
        a[t] = l;
#pragma unroll
        for (int jj = 0; jj < NB; ++jj)
            if (jj > t) a[jj] = fma(-l, s_D[t][jj], a[jj]);
    }
    long long t4 = clock64();
    double sum = 0.0;
#pragma unroll
    for (int jj = 0; jj < NB; ++jj) sum += a[jj] + r[jj];
    out[tid] = sum;
    if (tid == 0) {
        cyc[0] = t1 - t0;
        cyc[1] = t4 - t3;
        cyc[2] = t2 - t1;
    }
}

int main() {
    double hA[NB * NB];
    for (int i = 0; i < NB; ++i)
        for (int j = 0; j < NB; ++j) hA[i * NB + j] = (i == j ? 8.0 : 0.0) + 1.0 / (1 + i + 2 * j);
    double *A, *out;
    long long *cyc, h[3];
    cudaMalloc(&A, sizeof(hA));
    cudaMalloc(&out, 8 * 1024);
    cudaMalloc(&cyc, 24);
    cudaMemcpy(A, hA, sizeof(hA), cudaMemcpyHostToDevice);
    const char *names[] = {"shuffle pivot row", "pivot row via shared memory", "block in shared memory, runtime loops"};
#define RUN(V, T)                                                                                     \
    for (int rep = 0; rep < 3; ++rep) diag<V><<<1, T>>>(A, out, cyc);                                 \
    cudaDeviceSynchronize();                                                                          \
    cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);                                                   \
    printf("%-40s %4d threads: diag %lld cycles, rows-below %lld cycles, store+sync %lld\n", names[V], T, h[0], h[1], h[2]);
    RUN(0, 32) RUN(0, 512) RUN(1, 32) RUN(1, 512) RUN(2, 32) RUN(2, 512)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
