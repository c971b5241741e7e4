// Covariance::ComputeKtrain / ComputeKtest as fused sm_100a kernels: pairwise distance +
// covariance evaluation + noise diagonal in one pass, Gram matrix written once to HBM.
//
// Replaces erl_covariance v0.2.0 ComputeKtrain / ComputeKtest (call sites
// src/vanilla_gp.cpp:486-487, :537; src/sparse_pseudo_input_gp.cpp:340, 761-762).
//
// Roofline: HBM-write bound.  Algorithmic bytes = (n1 + n2) * x_dim * s + n1 * n2 * s
// (SURVEY.md 8d).  A CTA owns a 128 x 32 tile of K: the 128 row points and 32 column points
// are staged once in shared memory (coalesced loads of the point-contiguous inputs), each
// thread produces 4 consecutive rows x 4 columns and stores them as 128-bit (float) /
// 2 x 128-bit (double) vectors, so every warp store instruction writes full 128 B lines of a
// column of the column-major output.
#include "erl_gp_internal.cuh"

namespace erl_gp {

    constexpr int kGramTileRows = 128;
    constexpr int kGramTileCols = 32;
    constexpr int kGramThreads = 256;

    template<typename T, int XDIM, bool TRAIN>
    __global__ void __launch_bounds__(kGramThreads)
    GramKernel(
        const Covariance<T> cov,
        const T *__restrict__ x1,
        const long ld_x1,
        const long n1,
        const T *__restrict__ x2,
        const long ld_x2,
        const long n2,
        const T *__restrict__ var,
        T *__restrict__ k,
        const long ld_k,
        const bool vec_ok) {
        __shared__ T s_x1[kGramTileRows * XDIM];
        __shared__ T s_x2[kGramTileCols * XDIM];

        const long row0 = static_cast<long>(blockIdx.x) * kGramTileRows;
        const long col0 = static_cast<long>(blockIdx.y) * kGramTileCols;
        const int tid = threadIdx.x;

        for (int e = tid; e < kGramTileRows * XDIM; e += kGramThreads) {
            const long i = row0 + e / XDIM;
            s_x1[e] = i < n1 ? x1[i * ld_x1 + e % XDIM] : T(0);
        }
        for (int e = tid; e < kGramTileCols * XDIM; e += kGramThreads) {
            const long j = col0 + e / XDIM;
            s_x2[e] = j < n2 ? x2[j * ld_x2 + e % XDIM] : T(0);
        }
        __syncthreads();

        // thread -> rows 4*(tid % 32) .. +3, columns (tid / 32) + 8*c, c = 0..3
        const int r4 = (tid & 31) * 4;
        const int cbase = tid >> 5;
        T xi[4][XDIM];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int d = 0; d < XDIM; ++d) { xi[a][d] = s_x1[(r4 + a) * XDIM + d]; }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int cj = cbase + 8 * c;
            const long j = col0 + cj;
            if (j >= n2) { continue; }
            T xj[XDIM];
#pragma unroll
            for (int d = 0; d < XDIM; ++d) { xj[d] = s_x2[cj * XDIM + d]; }
            T v[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const long i = row0 + r4 + a;
                v[a] = cov(SquaredDistance<T, XDIM>(xi[a], xj));
                if (TRAIN && i == j && i < n1) { v[a] = T(1) + var[i]; }
            }
            T *dst = k + (row0 + r4) + j * ld_k;
            if (vec_ok && row0 + r4 + 3 < n1) {
                if constexpr (sizeof(T) == 4) {
                    *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
                    *reinterpret_cast<double2 *>(dst) = make_double2(v[0], v[1]);
                    *reinterpret_cast<double2 *>(dst + 2) = make_double2(v[2], v[3]);
                }
            } else {
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    if (row0 + r4 + a < n1) { dst[a] = v[a]; }
                }
            }
        }
    }

    template<typename T, bool TRAIN>
    static int
    LaunchGram(
        Context *ctx,
        int kernel,
        T scale,
        long x_dim,
        const T *x1,
        long ld_x1,
        long n1,
        const T *x2,
        long ld_x2,
        long n2,
        const T *var,
        T *k,
        long ld_k) {
        if (n1 <= 0 || n2 <= 0) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "gram: empty input (n1=%ld, n2=%ld)", n1, n2); }
        if (x_dim < 1 || x_dim > 3) { return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "gram: x_dim=%ld (supported: 1, 2, 3)", x_dim); }
        if (kernel < ERL_GP_KERNEL_OU || kernel > ERL_GP_KERNEL_RBF) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "gram: unknown kernel %d", kernel); }
        if (ld_x1 < x_dim || ld_x2 < x_dim || ld_k < n1) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "gram: leading dimension too small"); }
        const Covariance<T> cov = Covariance<T>::Make(kernel, scale);
        const dim3 grid(static_cast<unsigned>(CeilDiv(n1, kGramTileRows)), static_cast<unsigned>(CeilDiv(n2, kGramTileCols)));
        const bool vec_ok = (ld_k % 4 == 0) && (reinterpret_cast<uintptr_t>(k) % 16 == 0);
#define ERL_GP_GRAM_CASE(D)                                                                                                       \
    case D:                                                                                                                       \
        GramKernel<T, D, TRAIN><<<grid, kGramThreads, 0, ctx->stream>>>(cov, x1, ld_x1, n1, x2, ld_x2, n2, var, k, ld_k, vec_ok); \
        break;
        switch (x_dim) {
            ERL_GP_GRAM_CASE(1)
            ERL_GP_GRAM_CASE(2)
            ERL_GP_GRAM_CASE(3)
        }
#undef ERL_GP_GRAM_CASE
        ctx->launches += 1;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    int
    LaunchKtrain(Context *ctx, int kernel, T scale, long x_dim, const T *x, long ld_x, const T *var, long n, T *k, long ld_k) {
        return LaunchGram<T, true>(ctx, kernel, scale, x_dim, x, ld_x, n, x, ld_x, n, var, k, ld_k);
    }

    template<typename T>
    int
    LaunchKtest(Context *ctx, int kernel, T scale, long x_dim, const T *x1, long ld_x1, long n1, const T *x2, long ld_x2, long n2, T *k, long ld_k) {
        return LaunchGram<T, false>(ctx, kernel, scale, x_dim, x1, ld_x1, n1, x2, ld_x2, n2, nullptr, k, ld_k);
    }

    template int LaunchKtrain<float>(Context *, int, float, long, const float *, long, const float *, long, float *, long);
    template int LaunchKtrain<double>(Context *, int, double, long, const double *, long, const double *, long, double *, long);
    template int LaunchKtest<float>(Context *, int, float, long, const float *, long, long, const float *, long, long, float *, long);
    template int LaunchKtest<double>(Context *, int, double, long, const double *, long, long, const double *, long, long, double *, long);

}  // namespace erl_gp
