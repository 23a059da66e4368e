// Microbenchmark (sm_100a): dependent-issue latencies of the FP64 operations on the critical path of the
// diagonal-block factorisation (DFMA, DMUL, rcp.approx.ftz.f64 + 2 Newton steps, 64-bit warp shuffle, LDS).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

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
__global__ void chain(double *out, long long *cyc, double a, double b, int n) {
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = a;
    __syncthreads();
    double x = a + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        if (MODE == 0) x = fma(x, b, a);
        if (MODE == 1) x = x * b;
        if (MODE == 2) x = fast_rcp(x) + a;
        if (MODE == 3) x = __shfl_sync(0xffffffffu, x, (i & 31));
        if (MODE == 4) x = sm[(int)(x) & 63] + 0.0 * x;
        if (MODE == 5) x = 1.0 / x + a;
        if (MODE == 6) { float f = __fmaf_rn((float)x, 1.0f, 0.5f); x = (double)f; }
        if (MODE == 7) x = __shfl_sync(0xffffffffu, x, (i & 31)) * b;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x;
}

int main() {
    double *out;
    long long *cyc, h;
    cudaMalloc(&out, 1024 * 8);
    cudaMalloc(&cyc, 8);
    const int n = 4096;
    const char *names[] = {"DFMA dependent", "DMUL dependent", "fast_rcp (+1 DADD)", "SHFL.f64 dependent", "LDS + cvt dependent",
                           "IEEE 1/x (+1 DADD)", "cvt f64->f32, FFMA, cvt back", "SHFL.f64 + DMUL"};
#define RUN(M, threads)                                                        \
    chain<M><<<1, threads>>>(out, cyc, 1.000001, 0.999999, n);                 \
    chain<M><<<1, threads>>>(out, cyc, 1.000001, 0.999999, n);                 \
    cudaDeviceSynchronize();                                                   \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                            \
    printf("%-32s %4d threads: %.1f cycles / iteration\n", names[M], threads, (double)h / n);
    RUN(0, 32) RUN(0, 128) RUN(0, 512) RUN(1, 32) RUN(2, 32) RUN(3, 32) RUN(4, 32) RUN(5, 32) RUN(6, 32) RUN(7, 32)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
