// Microbenchmark of the "row substitution" phase of mf_factor_kernel (L = A21 U11^-1 for one 16-column panel):
//   A  the recurrence the kernel uses today (16 dependent steps per row, U11 read from shared memory),
//   B  the same rows as a product with the explicit inverse U11^-1 (16 independent dot products per row).
// One CTA of 256 threads, one row per thread, result stored to a shared panel like the kernel does.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lrows lrows.cu && ./lrows
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

template <int MODE>
__global__ void __launch_bounds__(256) lrows(const double *A, const double *Dg, double *out, long long *cycles, int kb, int reps) {
    __shared__ double D[NB][NB + 1], Ui[NB][NB + 1], rd[NB], P[256 * NB];
    const int tid = threadIdx.x;
    if (tid < NB * NB) {
        D[tid / NB][tid % NB] = Dg[tid];
        Ui[tid / NB][tid % NB] = Dg[256 + tid];
    }
    if (tid < NB) rd[tid] = fast_rcp(Dg[tid * NB + tid]);
    __syncthreads();
    double a[NB];
    long long t0 = 0, acc = 0;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) a[jj] = __ldcg(A + tid + (size_t)jj * 256 + (size_t)r * 4096);
        __syncthreads();
        if (tid == 0) t0 = clock64();
        if (MODE == 0) {
#pragma unroll
            for (int t = 0; t < NB; ++t) {
                if (t < kb) {
                    const double l = a[t] * rd[t];
                    a[t] = l;
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj)
                        if (jj > t && jj < kb) a[jj] = fma(-l, D[t][jj], a[jj]);
                }
            }
        } else {
            double l[NB];
#pragma unroll
            for (int jj = 0; jj < NB; ++jj) {
                double s = 0.0;
#pragma unroll
                for (int t = 0; t < NB; ++t)
                    if (t <= jj) s = fma(a[t], Ui[t][jj], s);
                l[jj] = s;
            }
#pragma unroll
            for (int jj = 0; jj < NB; ++jj) a[jj] = l[jj];
        }
#pragma unroll
        for (int jj = 0; jj < NB; ++jj)
            if (jj < kb) P[tid + jj * 256] = a[jj];
        __syncthreads();
        if (tid == 0) acc += clock64() - t0;
    }
    out[tid] = P[tid] + P[tid + 256 * 15];
    if (tid == 0) cycles[MODE] = acc / reps;
}

int main() {
    double *A, *D, *out;
    long long *cyc;
    cudaMalloc(&A, sizeof(double) * 4096 * 64);
    cudaMalloc(&D, sizeof(double) * 512);
    cudaMalloc(&out, sizeof(double) * 256);
    cudaMallocManaged(&cyc, sizeof(long long) * 2);
    double h[512];
    for (int i = 0; i < 512; ++i) h[i] = (i % 17 == 0) ? 4.0 : 0.01 * (i % 7);
    cudaMemcpy(D, h, sizeof h, cudaMemcpyHostToDevice);
    cudaMemset(A, 0, sizeof(double) * 4096 * 64);
    for (int it = 0; it < 2; ++it) {
        lrows<0><<<1, 256>>>(A, D, out, cyc, 16, 64);
        lrows<1><<<1, 256>>>(A, D, out, cyc, 16, 64);
        cudaDeviceSynchronize();
    }
    printf("row substitution, 256 rows, one CTA: recurrence %lld cycles, explicit inverse %lld cycles per panel\n", cyc[0], cyc[1]);
    return 0;
}
