// GPU numeric phase of the multifrontal LU (see multifrontal.hpp): one CTA per front, one launch per level of
// the nested-dissection tree.  Fronts are dense column-major m x m blocks in one HBM workspace (L2 resident for
// the reference meshes); each CTA extend-adds its children's Schur complements, then runs a right-looking
// blocked LU of the np fully-summed columns (panel of 16 columns in shared memory, 4x4 register tiles for the
// trailing update) with partial pivoting restricted to the fully-summed rows.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"
#include "multifrontal.cuh"

namespace cg = cooperative_groups;

namespace ocp {

namespace {

constexpr int NB = 16;     // panel width
constexpr int CW = 128;    // trailing-update column chunk
constexpr int TF = 512;    // threads per front CTA
constexpr int SB = 32;     // triangular-solve block (one warp)

struct MFDev {
    const int *m, *np, *first, *idx_ptr, *idx, *child_ptr, *child, *rel_ptr, *rel;
    const long long *front_ptr;
    double *F;
    int *piv;
};

__global__ void scatter_values_kernel(int nnz, const long long *__restrict__ dest, const double *__restrict__ vals,
                                      double *__restrict__ F) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) F[dest[k]] = vals[k];
}

// One CLUSTER of C CTAs per front (C = 1 for the many small fronts at the bottom of the tree, 8-16 for the few
// large fronts near the root).  Per 16-column panel: CTA 0 factors the panel in its shared memory, the cluster
// synchronises, then every CTA applies the row interchanges, solves U12 and updates the trailing matrix for the
// column chunks it owns (chunk = absolute column index / cw, owner = chunk % C), and the cluster synchronises again.
// All front accesses bypass L1 (ld.cg / st.cg): columns migrate between the panel owner (CTA 0) and chunk owners.
__global__ void __launch_bounds__(TF)
mf_factor_kernel(MFDev d, const int *__restrict__ nodes, int max_m, int cw, int *info) {
    extern __shared__ double sm[];
    __shared__ int s_piv[NB];
    cg::cluster_group cl = cg::this_cluster();
    const int C = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    double *P = sm;                              // panel, ld = mp
    double *Uc = sm + (size_t)max_m * NB;        // NB x CW, row-major
    const int s = nodes[blockIdx.x / C];
    const int m = d.m[s], np = d.np[s], tid = threadIdx.x;
    double *F = d.F + d.front_ptr[s];
    // ---- extend-add the children's update matrices (atomic: two children may hit the same entry)
    for (int ci = d.child_ptr[s]; ci < d.child_ptr[s + 1]; ++ci) {
        const int c = d.child[ci], mc = d.m[c], npc = d.np[c], nu = mc - npc;
        const double *Fc = d.F + d.front_ptr[c];
        const int *rel = d.rel + d.rel_ptr[c];
        for (int j = rank; j < nu; j += C) {
            const double *src = Fc + npc + (size_t)(npc + j) * mc;
            double *dst = F + (size_t)__ldg(rel + j) * m;
            for (int i = tid; i < nu; i += TF) atomicAdd(dst + __ldg(rel + i), __ldcg(src + i));
        }
    }
    cl.sync();
    if (np == 0) return;
    int *gpiv = d.piv + d.first[s];
    for (int k0 = 0; k0 < np; k0 += NB) {
        const int kb = min(NB, np - k0), mp = m - k0;
        for (int j = 0; j < kb; ++j)
            for (int i = tid; i < mp; i += TF) P[i + j * mp] = __ldcg(F + (k0 + i) + (size_t)(k0 + j) * m);
        if (rank == 0) {
            __syncthreads();
            // ---- panel factorisation, pivot rows restricted to the fully-summed block
            for (int j = 0; j < kb; ++j) {
                if (tid < 32) {
                    double best = -1.0;
                    int r = j;
                    for (int i = j + tid; i < np - k0; i += 32) {
                        const double a = fabs(P[i + j * mp]);
                        if (a > best) {
                            best = a;
                            r = i;
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                        const int orr = __shfl_xor_sync(0xffffffffu, r, o);
                        if (ob > best || (ob == best && orr < r)) {
                            best = ob;
                            r = orr;
                        }
                    }
                    if (tid == 0) {
                        s_piv[j] = r;
                        gpiv[k0 + j] = k0 + r;
                        if (!(best > 0.0)) atomicExch(info, s + 1);
                    }
                }
                __syncthreads();
                const int r = s_piv[j];
                if (r != j && tid < kb) {
                    const double t = P[j + tid * mp];
                    P[j + tid * mp] = P[r + tid * mp];
                    P[r + tid * mp] = t;
                }
                __syncthreads();
                const double inv = 1.0 / P[j + j * mp];
                for (int i = j + 1 + tid; i < mp; i += TF) {
                    const double l = P[i + j * mp] * inv;
                    P[i + j * mp] = l;
                    for (int jj = j + 1; jj < kb; ++jj) P[i + jj * mp] -= l * P[j + jj * mp];
                }
                __syncthreads();
            }
            for (int j = 0; j < kb; ++j)
                for (int i = tid; i < mp; i += TF) __stcg(F + (k0 + i) + (size_t)(k0 + j) * m, P[i + j * mp]);
        }
        cl.sync();
        if (rank != 0) {
            // the factored panel and its pivots, written by CTA 0
            for (int j = 0; j < kb; ++j)
                for (int i = tid; i < mp; i += TF) P[i + j * mp] = __ldcg(F + (k0 + i) + (size_t)(k0 + j) * m);
            if (tid < kb) s_piv[tid] = __ldcg(gpiv + k0 + tid) - k0;
        }
        __syncthreads();
        // ---- own column chunks: row interchanges, U12 = L11^{-1} F12, trailing update
        const int nrow = m - k0 - kb;
        for (int cbeg = rank * cw; cbeg < m; cbeg += C * cw) {
            const int cend = min(m, cbeg + cw);
            const int lo = max(cbeg, k0 + kb);        // first trailing column of this chunk
            const int c = cbeg + tid;
            if (c < cend && !(c >= k0 && c < k0 + kb)) {
                double *colp = F + (size_t)c * m + k0;
                for (int j = 0; j < kb; ++j) {
                    const int r = s_piv[j];
                    if (r != j) {
                        const double t = __ldcg(colp + j);
                        __stcg(colp + j, __ldcg(colp + r));
                        __stcg(colp + r, t);
                    }
                }
                if (c >= k0 + kb) {
                    double u[NB];
#pragma unroll
                    for (int t = 0; t < NB; ++t) u[t] = t < kb ? __ldcg(colp + t) : 0.0;
#pragma unroll
                    for (int t = 1; t < NB; ++t) {
                        if (t < kb) {
                            double a = u[t];
#pragma unroll
                            for (int tt = 0; tt < NB; ++tt)
                                if (tt < t) a -= P[t + tt * mp] * u[tt];
                            u[t] = a;
                        }
                    }
#pragma unroll
                    for (int t = 0; t < NB; ++t) {
                        if (t < kb) __stcg(colp + t, u[t]);
                        Uc[t * CW + (c - lo)] = u[t];
                    }
                }
            }
            __syncthreads();
            const int ncol = cend - lo;
            if (nrow > 0 && ncol > 0) {
                const int tr = (nrow + 3) >> 2, tc = (ncol + 3) >> 2;
                for (int tile = tid; tile < tr * tc; tile += TF) {
                    const int ti = tile % tr, tj = tile / tr;
                    const int i0 = kb + 4 * ti, cc0 = 4 * tj;
                    double f[4][4], acc[4][4];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const double *colp = F + (size_t)(lo + cc0 + b) * m + k0;
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            f[a][b] = (cc0 + b < ncol && i0 + a < mp) ? __ldcg(colp + i0 + a) : 0.0;
                            acc[a][b] = 0.0;
                        }
                    }
                    for (int t = 0; t < kb; ++t) {
                        double l[4], uu[4];
#pragma unroll
                        for (int a = 0; a < 4; ++a) l[a] = (i0 + a < mp) ? P[i0 + a + t * mp] : 0.0;
#pragma unroll
                        for (int b = 0; b < 4; ++b) uu[b] = (cc0 + b < ncol) ? Uc[t * CW + cc0 + b] : 0.0;
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int b = 0; b < 4; ++b) acc[a][b] = fma(l[a], uu[b], acc[a][b]);
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (cc0 + b < ncol) {
                            double *colp = F + (size_t)(lo + cc0 + b) * m + k0;
#pragma unroll
                            for (int a = 0; a < 4; ++a)
                                if (i0 + a < mp) __stcg(colp + i0 + a, f[a][b] - acc[a][b]);
                        }
                    }
                }
            }
            __syncthreads();
        }
        cl.sync();
    }
}

// forward substitution of one level: y_P = L11^{-1} Pi b_P,  b_U -= L21 y_P
__global__ void __launch_bounds__(TF)
mf_forward_kernel(MFDev d, const int *__restrict__ nodes, double *__restrict__ x) {
    extern __shared__ double y[];
    const int s = nodes[blockIdx.x];
    const int m = d.m[s], np = d.np[s], tid = threadIdx.x;
    if (np == 0) return;
    const double *F = d.F + d.front_ptr[s];
    const int *I = d.idx + d.idx_ptr[s];
    const int *gpiv = d.piv + d.first[s];
    for (int k = tid; k < m; k += TF) y[k] = k < np ? x[I[k]] : 0.0;
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < np; ++k) {
            const int r = gpiv[k];
            if (r != k) {
                const double t = y[k];
                y[k] = y[r];
                y[r] = t;
            }
        }
    }
    __syncthreads();
    for (int k0 = 0; k0 < np; k0 += SB) {
        const int kb = min(SB, np - k0);
        if (tid < 32) {
            double Lr[SB];
#pragma unroll
            for (int t = 0; t < SB; ++t) Lr[t] = (tid < kb && t < tid) ? F[(k0 + tid) + (size_t)(k0 + t) * m] : 0.0;
            double v = tid < kb ? y[k0 + tid] : 0.0;
#pragma unroll
            for (int t = 0; t < SB; ++t) {
                const double vt = __shfl_sync(0xffffffffu, v, t);
                if (t < kb && tid > t) v -= Lr[t] * vt;
            }
            if (tid < kb) y[k0 + tid] = v;
        }
        __syncthreads();
        for (int i = k0 + kb + tid; i < m; i += TF) {
            double acc = 0.0;
            for (int t = 0; t < kb; ++t) acc = fma(F[i + (size_t)(k0 + t) * m], y[k0 + t], acc);
            y[i] -= acc;
        }
        __syncthreads();
    }
    for (int k = tid; k < np; k += TF) x[I[k]] = y[k];
    for (int i = np + tid; i < m; i += TF) atomicAdd(x + I[i], y[i]);
}

// backward substitution of one level: x_P = U11^{-1} (y_P - U12 x_U)
__global__ void __launch_bounds__(TF)
mf_backward_kernel(MFDev d, const int *__restrict__ nodes, double *__restrict__ x) {
    extern __shared__ double y[];
    const int s = nodes[blockIdx.x];
    const int m = d.m[s], np = d.np[s], tid = threadIdx.x;
    if (np == 0) return;
    const double *F = d.F + d.front_ptr[s];
    const int *I = d.idx + d.idx_ptr[s];
    for (int k = tid; k < m; k += TF) y[k] = x[I[k]];
    __syncthreads();
    for (int k = tid; k < np; k += TF) {
        double acc = 0.0;
        for (int j = np; j < m; ++j) acc = fma(F[k + (size_t)j * m], y[j], acc);
        y[k] -= acc;
    }
    __syncthreads();
    const int nblk = (np + SB - 1) / SB;
    for (int b = nblk - 1; b >= 0; --b) {
        const int k0 = b * SB, kb = min(SB, np - k0);
        if (tid < 32) {
            double Ur[SB];
#pragma unroll
            for (int t = 0; t < SB; ++t) Ur[t] = (tid < kb && t < kb && t >= tid) ? F[(k0 + tid) + (size_t)(k0 + t) * m] : 1.0;
            double v = tid < kb ? y[k0 + tid] : 0.0;
#pragma unroll
            for (int t = SB - 1; t >= 0; --t) {
                if (t < kb) {
                    if (tid == t) v = v / Ur[t];
                    const double vt = __shfl_sync(0xffffffffu, v, t);
                    if (tid < t) v -= Ur[t] * vt;
                }
            }
            if (tid < kb) y[k0 + tid] = v;
        }
        __syncthreads();
        for (int i = tid; i < k0; i += TF) {
            double acc = 0.0;
            for (int t = 0; t < kb; ++t) acc = fma(F[i + (size_t)(k0 + t) * m], y[k0 + t], acc);
            y[i] -= acc;
        }
        __syncthreads();
    }
    for (int k = tid; k < np; k += TF) x[I[k]] = y[k];
}

template <class T>
bool up(T **dst, const std::vector<T> &src, std::string &err) {
    cudaError_t e = cudaMalloc((void **)dst, sizeof(T) * std::max<size_t>(src.size(), 1));
    if (e == cudaSuccess && !src.empty())
        e = cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        err = std::string("multifrontal upload: ") + cudaGetErrorString(e);
        return false;
    }
    return true;
}

}  // namespace

struct MultifrontalLU::Impl {
    MFSymbolic S;
    int *m = nullptr, *np = nullptr, *first = nullptr, *idx_ptr = nullptr, *idx = nullptr, *child_ptr = nullptr,
        *child = nullptr, *rel_ptr = nullptr, *rel = nullptr, *level_nodes = nullptr, *piv = nullptr, *info = nullptr;
    long long *front_ptr = nullptr, *a_dest = nullptr;
    double *F = nullptr;
    std::vector<int> level_max_m, level_cluster, level_cw;
    MFDev dev{};
    ~Impl() {
        void *p[] = {m, np, first, idx_ptr, idx, child_ptr, child, rel_ptr, rel, level_nodes, piv, info, front_ptr,
                     a_dest, F};
        for (void *q : p) cudaFree(q);
    }
};

MultifrontalLU::MultifrontalLU() = default;
MultifrontalLU::~MultifrontalLU() { delete impl_; }

bool MultifrontalLU::configure(int n, int nnz, const int *h_rowptr, const int *h_col, const double *xy,
                               const unsigned char *kind, std::string &err) {
    auto t0 = std::chrono::steady_clock::now();
    delete impl_;
    impl_ = new Impl();
    Impl &I = *impl_;
    n_ = n;
    nnz_ = nnz;
    mf_analyse(n, h_rowptr, h_col, xy, kind, 48, I.S);
    const MFSymbolic &S = I.S;
    for (long long dd : S.a_dest)
        if (dd < 0) {
            err = "multifrontal analysis: matrix entry outside its front";
            return false;
        }
    I.level_max_m.assign(S.nlevels, 0);
    for (int l = 0; l < S.nlevels; ++l)
        for (int k = S.level_ptr[l]; k < S.level_ptr[l + 1]; ++k)
            I.level_max_m[l] = std::max(I.level_max_m[l], S.m[S.level_nodes[k]]);
    const size_t need = ((size_t)S.max_front * NB + (size_t)NB * CW) * sizeof(double);
    if (need > 220 * 1024) {
        err = "multifrontal: largest front (" + std::to_string(S.max_front) + ") exceeds the shared-memory panel";
        return false;
    }
    if (!up(&I.m, S.m, err) || !up(&I.np, S.np, err) || !up(&I.first, S.first, err) ||
        !up(&I.idx_ptr, S.idx_ptr, err) || !up(&I.idx, S.idx, err) || !up(&I.child_ptr, S.child_ptr, err) ||
        !up(&I.child, S.child, err) || !up(&I.rel_ptr, S.rel_ptr, err) || !up(&I.rel, S.rel, err) ||
        !up(&I.level_nodes, S.level_nodes, err) || !up(&I.front_ptr, S.front_ptr, err) ||
        !up(&I.a_dest, S.a_dest, err))
        return false;
    cudaError_t e = cudaMalloc((void **)&I.F, sizeof(double) * S.fsize);
    if (e == cudaSuccess) e = cudaMalloc((void **)&I.piv, sizeof(int) * std::max(n, 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&I.info, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(I.info, 0, sizeof(int));
    if (e == cudaSuccess)
        // the attribute is per kernel, not per solver instance: always allow the full opt-in budget
        e = cudaFuncSetAttribute(mf_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mf_factor_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    // cluster size per level: as many CTAs per front as the chip has room for (powers of two, <= 16)
    int max_cluster = 16;
    if (const char *envc = getenv("OCP_MF_MAX_CLUSTER")) max_cluster = std::max(1, atoi(envc));
    I.level_cluster.assign(S.nlevels, 1);
    I.level_cw.assign(S.nlevels, 32);
    for (int l = 0; l < S.nlevels && e == cudaSuccess; ++l) {
        const int nf = S.level_ptr[l + 1] - S.level_ptr[l];
        int c = 1;
        while (c * 2 <= max_cluster && nf * c * 2 <= 148 && c * 2 * 16 <= I.level_max_m[l]) c *= 2;
        while (c > 1) {   // make sure the cluster shape is launchable with this kernel's resources
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(nf * c);
            cfg.blockDim = dim3(TF);
            cfg.dynamicSmemBytes = ((size_t)I.level_max_m[l] * NB + (size_t)NB * CW) * sizeof(double);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = c;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, mf_factor_kernel, &cfg) == cudaSuccess && ncl >= 1) break;
            cudaGetLastError();
            c /= 2;
        }
        I.level_cluster[l] = c;
        I.level_cw[l] = (c >= 16) ? 16 : 32;
    }
    if (e != cudaSuccess) {
        err = std::string("multifrontal setup: ") + cudaGetErrorString(e);
        return false;
    }
    I.dev = MFDev{I.m, I.np, I.first, I.idx_ptr, I.idx, I.child_ptr, I.child, I.rel_ptr, I.rel, I.front_ptr, I.F, I.piv};
    factor_nnz_ = 0;
    for (int s = 0; s < S.nnodes; ++s)
        factor_nnz_ += (long long)S.m[s] * S.m[s] - (long long)(S.m[s] - S.np[s]) * (S.m[s] - S.np[s]);
    flops_ = S.flops;
    nlevels_ = S.nlevels;
    max_front_ = S.max_front;
    analyse_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return true;
}

bool MultifrontalLU::factor(const double *d_vals, cudaStream_t s, std::string &err) {
    if (!impl_) {
        err = "MultifrontalLU::factor before configure";
        return false;
    }
    Impl &I = *impl_;
    const MFSymbolic &S = I.S;
    cudaMemsetAsync(I.F, 0, sizeof(double) * S.fsize, s);
    g_launch_count.fetch_add(1 + S.nlevels, std::memory_order_relaxed);
    scatter_values_kernel<<<(nnz_ + 255) / 256, 256, 0, s>>>(nnz_, I.a_dest, d_vals, I.F);
    for (int l = 0; l < S.nlevels; ++l) {
        const int nf = S.level_ptr[l + 1] - S.level_ptr[l];
        const int c = I.level_cluster[l];
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(nf * c);
        cfg.blockDim = dim3(TF);
        cfg.dynamicSmemBytes = ((size_t)I.level_max_m[l] * NB + (size_t)NB * CW) * sizeof(double);
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = c;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaError_t le = cudaLaunchKernelEx(&cfg, mf_factor_kernel, I.dev, (const int *)(I.level_nodes + S.level_ptr[l]),
                                            I.level_max_m[l], I.level_cw[l], I.info);
        if (le != cudaSuccess) {
            err = std::string("multifrontal factor launch (level ") + std::to_string(l) + ", cluster " +
                  std::to_string(c) + "): " + cudaGetErrorString(le);
            return false;
        }
    }
    int info = 0;
    cudaError_t e = cudaMemcpyAsync(&info, I.info, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("multifrontal factor: ") + cudaGetErrorString(e);
        return false;
    }
    if (info != 0) {
        err = "multifrontal factor: zero pivot in front " + std::to_string(info - 1);
        cudaMemsetAsync(I.info, 0, sizeof(int), s);
        return false;
    }
    return true;
}

bool MultifrontalLU::solve(double *d_x, cudaStream_t s, std::string &err) {
    if (!impl_) {
        err = "MultifrontalLU::solve before configure";
        return false;
    }
    Impl &I = *impl_;
    const MFSymbolic &S = I.S;
    g_launch_count.fetch_add(2 * S.nlevels, std::memory_order_relaxed);
    for (int l = 0; l < S.nlevels; ++l) {
        const int nf = S.level_ptr[l + 1] - S.level_ptr[l];
        mf_forward_kernel<<<nf, TF, sizeof(double) * I.level_max_m[l], s>>>(I.dev, I.level_nodes + S.level_ptr[l], d_x);
    }
    for (int l = S.nlevels - 1; l >= 0; --l) {
        const int nf = S.level_ptr[l + 1] - S.level_ptr[l];
        mf_backward_kernel<<<nf, TF, sizeof(double) * I.level_max_m[l], s>>>(I.dev, I.level_nodes + S.level_ptr[l], d_x);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("multifrontal solve: ") + cudaGetErrorString(e);
        return false;
    }
    return true;
}

}  // namespace ocp
