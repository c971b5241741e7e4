// Dense (HBM-resident) linear algebra for large VanillaGaussianProcess and SPGP systems.
//
// Replaces Eigen's `llt()` and `triangularView::solve` as used at src/vanilla_gp.cpp:499-502,
// :139-149 and src/sparse_pseudo_input_gp.cpp:341, 100-106, 304-309, 768-774, 840.
//
// Blocked right-looking Cholesky, 128-column panels inside 512-column block columns (Potrf):
//   1. DiagFactorKernel: one CTA factors the 128x128 diagonal block entirely in shared memory (the same
//      block-packed 16x16 machinery as the generic batched small-GP kernel) and also produces its INVERSE, so that
//   2. the panel solve  L21 = A21 * L11^-T  (in place) and every later triangular solve (predictive variance, SPGP)
//      are plain GEMMs against the kept 128x128 inverses, and
//   3. the trailing updates are lower-triangle-only GEMMs (rank 128 inside the block column, rank 512 to its right);
//   4. the latency-bound panel chain is looked ahead on two high-priority side streams (block column and next diagonal tile).
// FP64 GEMMs run on the tensor path (mma.sync.m8n8k4.f64, SASS DMMA; 16 warps per CTA, 32 x 32 warp tiles): the DFMA loop
// (GemmKernel<double>, ERL_GP_DENSE_FMA=1) has the same peak (36.2 vs 37.0 TFLOP/s measured, tools/mma_rate.cu) but is
// register-file bound at 52 % of it.  FP32 GEMMs are true-FP32 FFMA tiles (GemmKernel<float>).
// alpha = L^-T L^-1 y for a few right-hand sides is one wavefront kernel per direction (TrsvWavefrontKernel).
#include "erl_gp_dense.cuh"
#include "erl_gp_dense_mma.cuh"

#include <cstdlib>

#include "erl_gp_batched.cuh"
#include "erl_gp_rowgp64.cuh"

namespace erl_gp {

    // =========================================================================================
    // GEMM
    // =========================================================================================
    constexpr int kGemmBM = 128, kGemmBN = 128, kGemmBK = 16, kGemmThreads = 256, kGemmPad = 4;

    // load a (ROWS x BK) operand tile into registers: ROWS = 128 "outer" index, BK = 16 reduction index
    //   K_CONTIG = false: element (o, k) at src[o + k * ld]   (outer index contiguous)
    //   K_CONTIG = true : element (o, k) at src[k + o * ld]   (reduction index contiguous)
    template<typename T, bool K_CONTIG>
    __device__ __forceinline__ void
    GemmLoadTile(const T *__restrict__ src, const long ld, const long o0, const long o_lim, const long k0, const long k_lim, const int tid, T (&reg)[8]) {
        if (!K_CONTIG) {
            const int o = (tid & 31) * 4;
            const int kk = tid >> 5;  // 0..7
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const long k = k0 + kk + 8 * i;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const long oo = o0 + o + j;
                    reg[i * 4 + j] = (oo < o_lim && k < k_lim) ? src[oo + k * ld] : T(0);
                }
            }
        } else {
            const int kk = tid & 15;
            const int o = tid >> 4;  // 0..15
            const long k = k0 + kk;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long oo = o0 + o + 16 * i;
                reg[i] = (oo < o_lim && k < k_lim) ? src[k + oo * ld] : T(0);
            }
        }
    }

    template<typename T, bool K_CONTIG>
    __device__ __forceinline__ void
    GemmStoreTile(T *__restrict__ dst /* [BK][128 + pad] */, const int tid, const T (&reg)[8]) {
        constexpr int kLd = kGemmBM + kGemmPad;
        if (!K_CONTIG) {
            const int o = (tid & 31) * 4;
            const int kk = tid >> 5;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                T t4[4] = {reg[i * 4], reg[i * 4 + 1], reg[i * 4 + 2], reg[i * 4 + 3]};
                Store4(dst + (kk + 8 * i) * kLd + o, t4);
            }
        } else {
            const int kk = tid & 15;
            const int o = tid >> 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) { dst[kk * kLd + o + 16 * i] = reg[i]; }
        }
    }

    template<typename T, bool A_KC, bool B_KC>
    __global__ void __launch_bounds__(kGemmThreads, sizeof(T) == 4 ? 2 : 1)
    GemmKernel(
        const long m,
        const long n,
        const long k,
        const T alpha,
        const T *__restrict__ a,
        const long lda,
        const T *__restrict__ b,
        const long ldb,
        const T beta,
        T *__restrict__ c,
        const long ldc,
        const int lower_only) {
        constexpr int kLd = kGemmBM + kGemmPad;
        extern __shared__ __align__(16) unsigned char smem_raw[];
        T *as = reinterpret_cast<T *>(smem_raw);       // [2][BK][kLd]
        T *bs = as + 2 * kGemmBK * kLd;                // [2][BK][kLd]
        const long row0 = static_cast<long>(blockIdx.x) * kGemmBM;
        const long col0 = static_cast<long>(blockIdx.y) * kGemmBN;
        if (lower_only && col0 > row0 + kGemmBM - 1) { return; }  // tile strictly above the diagonal
        if (lower_only == 2 && blockIdx.x == 0 && blockIdx.y == 0) { return; }  // diagonal tile (0, 0) is updated by the look-ahead stream
        const int tid = threadIdx.x;
        const int tx = tid & 15, ty = tid >> 4;

        T acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[i][j] = T(0); }
        }
        T ra[8], rb[8];
        GemmLoadTile<T, A_KC>(a, lda, row0, m, 0, k, tid, ra);
        GemmLoadTile<T, B_KC>(b, ldb, col0, n, 0, k, tid, rb);
        GemmStoreTile<T, A_KC>(as, tid, ra);
        GemmStoreTile<T, B_KC>(bs, tid, rb);
        __syncthreads();
        const long num_kt = (k + kGemmBK - 1) / kGemmBK;
        for (long kt = 0; kt < num_kt; ++kt) {
            const int cur = static_cast<int>(kt & 1);
            if (kt + 1 < num_kt) {
                GemmLoadTile<T, A_KC>(a, lda, row0, m, (kt + 1) * kGemmBK, k, tid, ra);
                GemmLoadTile<T, B_KC>(b, ldb, col0, n, (kt + 1) * kGemmBK, k, tid, rb);
            }
            const T *at = as + cur * kGemmBK * kLd;
            const T *bt = bs + cur * kGemmBK * kLd;
#pragma unroll
            for (int kk = 0; kk < kGemmBK; ++kk) {
                T av[8], bv[8];
                T t4[4];
                Load4(at + kk * kLd + tx * 4, t4);
                av[0] = t4[0], av[1] = t4[1], av[2] = t4[2], av[3] = t4[3];
                Load4(at + kk * kLd + 64 + tx * 4, t4);
                av[4] = t4[0], av[5] = t4[1], av[6] = t4[2], av[7] = t4[3];
                Load4(bt + kk * kLd + ty * 4, t4);
                bv[0] = t4[0], bv[1] = t4[1], bv[2] = t4[2], bv[3] = t4[3];
                Load4(bt + kk * kLd + 64 + ty * 4, t4);
                bv[4] = t4[0], bv[5] = t4[1], bv[6] = t4[2], bv[7] = t4[3];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { acc[i][j] += av[i] * bv[j]; }
                }
            }
            if (kt + 1 < num_kt) {
                GemmStoreTile<T, A_KC>(as + (cur ^ 1) * kGemmBK * kLd, tid, ra);
                GemmStoreTile<T, B_KC>(bs + (cur ^ 1) * kGemmBK * kLd, tid, rb);
            }
            __syncthreads();
        }
        // epilogue
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const long col = col0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
            if (col >= n) { continue; }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long row = row0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
                if (row >= m || (lower_only && row < col)) { continue; }
                T *dst = c + row + col * ldc;
                const T prev = beta == T(0) ? T(0) : beta * (*dst);
                *dst = alpha * acc[i][j] + prev;
            }
        }
    }

    // FP64 on the tensor path (DMMA m8n8k4): same tiling / staging as GemmKernel, the 8 x 8 DFMA register tile is replaced
    // by the warp-level fragments of erl_gp_dense_mma.cuh (the DFMA loop is register-file bound at ~52 % of the FP64 peak)
    template<bool A_KC, bool B_KC>
    __global__ void __launch_bounds__(kGemmThreads, 1)
    GemmKernelDmma(const long m, const long n, const long k, const double alpha, const double *__restrict__ a, const long lda, const double *__restrict__ b, const long ldb, const double beta,
                   double *__restrict__ c, const long ldc, const int lower_only) {
        constexpr int kLd = kGemmBM + kGemmPad;
        static_assert(kLd == kMmaLd && kGemmBK == kMmaBk, "slab layout shared with erl_gp_dense_mma.cuh");
        extern __shared__ __align__(16) unsigned char smem_raw[];
        double *as = reinterpret_cast<double *>(smem_raw);
        double *bs = as + 2 * kGemmBK * kLd;
        const long row0 = static_cast<long>(blockIdx.x) * kGemmBM;
        const long col0 = static_cast<long>(blockIdx.y) * kGemmBN;
        if (lower_only && col0 > row0 + kGemmBM - 1) { return; }
        if (lower_only == 2 && blockIdx.x == 0 && blockIdx.y == 0) { return; }  // diagonal tile (0, 0) is updated by the look-ahead stream
        const int tid = threadIdx.x;
        const int lane = tid & 31, warp = tid >> 5;
        const int wm = warp & 1, wn = warp >> 1;
        const int g = lane >> 2, kq = lane & 3;
        double acc[8][4][2];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = acc[mi][ni][1] = 0.0; }
        }
        double ra[8], rb[8];
        GemmLoadTile<double, A_KC>(a, lda, row0, m, 0, k, tid, ra);
        GemmLoadTile<double, B_KC>(b, ldb, col0, n, 0, k, tid, rb);
        GemmStoreTile<double, A_KC>(as, tid, ra);
        GemmStoreTile<double, B_KC>(bs, tid, rb);
        __syncthreads();
        const long num_kt = (k + kGemmBK - 1) / kGemmBK;
        for (long kt = 0; kt < num_kt; ++kt) {
            const int cur = static_cast<int>(kt & 1);
            if (kt + 1 < num_kt) {
                GemmLoadTile<double, A_KC>(a, lda, row0, m, (kt + 1) * kGemmBK, k, tid, ra);
                GemmLoadTile<double, B_KC>(b, ldb, col0, n, (kt + 1) * kGemmBK, k, tid, rb);
            }
            SlabMma(acc, as + cur * kGemmBK * kLd, bs + cur * kGemmBK * kLd, wm, wn, lane);
            if (kt + 1 < num_kt) {
                GemmStoreTile<double, A_KC>(as + (cur ^ 1) * kGemmBK * kLd, tid, ra);
                GemmStoreTile<double, B_KC>(bs + (cur ^ 1) * kGemmBK * kLd, tid, rb);
            }
            __syncthreads();
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const long col = col0 + 32 * wn + 8 * ni + 2 * kq + e;
                if (col >= n) { continue; }
#pragma unroll
                for (int mi = 0; mi < 8; ++mi) {
                    const long row = row0 + 64 * wm + 8 * mi + g;
                    if (row >= m || (lower_only && row < col)) { continue; }
                    double *dst = c + row + col * ldc;
                    const double prev = beta == 0.0 ? 0.0 : beta * (*dst);
                    *dst = alpha * acc[mi][ni][e] + prev;
                }
            }
        }
    }

    // ---- the same GEMM with a 3-stage cp.async pipeline ---------------------------------------------------------------
    // ncu on the rank-512 update (profiles/r01k_dense_syrk_*): tensor pipe 73.5 % of the active cycles, the rest long-scoreboard
    // (the operand loads of slab k+1 are issued one slab = 1 us ahead, L2 hit rate 60 %) and fixed-latency stalls.  cp.async
    // (LDGSTS) keeps two slabs in flight, needs no staging registers and leaves one barrier per slab.  Measured slower than the
    // register-staged kernel (see GemmUseCpAsync): kept for A/B runs only.
    constexpr int kGemmStages = 3;

    __device__ __forceinline__ void
    CpAsync8(double *smem_dst, const double *gmem_src, const bool pred) {
        const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
        const int src_bytes = pred ? 8 : 0;  // 0: zero-fill, the source is not read
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(gmem_src), "r"(src_bytes) : "memory");
    }

    __device__ __forceinline__ void
    CpAsync16(double *smem_dst, const double *gmem_src, const bool pred) {
        const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
        const int src_bytes = pred ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(src_bytes) : "memory");
    }

    // one (128 x BK) operand slab: same element -> thread mapping as GemmLoadTile / GemmStoreTile.  vec16: the operand allows
    // 16-byte copies along the outer index (even leading dimension, 16-byte aligned base, even tile origin)
    template<bool K_CONTIG>
    __device__ __forceinline__ void
    GemmCpAsyncTile(double *__restrict__ dst /* [BK][kLd] */, const double *__restrict__ src, const long ld, const long o0, const long o_lim, const long k0, const long k_lim, const int tid,
                    const bool vec16) {
        constexpr int kLd = kGemmBM + kGemmPad;
        if (!K_CONTIG) {
            const int o = (tid & 31) * 4;
            const int kk = tid >> 5;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const long k = k0 + kk + 8 * i;
                if (vec16) {
#pragma unroll
                    for (int j = 0; j < 4; j += 2) {
                        const long oo = o0 + o + j;
                        if (oo + 1 < o_lim || oo >= o_lim) {
                            const bool ok = oo < o_lim && k < k_lim;
                            CpAsync16(dst + (kk + 8 * i) * kLd + o + j, ok ? src + oo + k * ld : src, ok);
                        } else {  // the pair straddles the edge
                            CpAsync8(dst + (kk + 8 * i) * kLd + o + j, k < k_lim ? src + oo + k * ld : src, k < k_lim);
                            CpAsync8(dst + (kk + 8 * i) * kLd + o + j + 1, src, false);
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const long oo = o0 + o + j;
                        const bool ok = oo < o_lim && k < k_lim;
                        CpAsync8(dst + (kk + 8 * i) * kLd + o + j, ok ? src + oo + k * ld : src, ok);
                    }
                }
            }
        } else {
            const int kk = tid & 15;
            const int o = tid >> 4;
            const long k = k0 + kk;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long oo = o0 + o + 16 * i;
                const bool ok = oo < o_lim && k < k_lim;
                CpAsync8(dst + kk * kLd + o + 16 * i, ok ? src + k + oo * ld : src, ok);
            }
        }
    }

    template<bool A_KC, bool B_KC>
    __global__ void __launch_bounds__(kGemmThreads, 1)
    GemmKernelDmmaAsync(const long m, const long n, const long k, const double alpha, const double *__restrict__ a, const long lda, const double *__restrict__ b, const long ldb, const double beta,
                        double *__restrict__ c, const long ldc, const int lower_only) {
        constexpr int kLd = kGemmBM + kGemmPad;
        constexpr int kSlab = kGemmBK * kLd;
        extern __shared__ __align__(16) unsigned char smem_raw[];
        double *as = reinterpret_cast<double *>(smem_raw);  // [stages][BK][kLd]
        double *bs = as + kGemmStages * kSlab;
        const long row0 = static_cast<long>(blockIdx.x) * kGemmBM;
        const long col0 = static_cast<long>(blockIdx.y) * kGemmBN;
        if (lower_only && col0 > row0 + kGemmBM - 1) { return; }
        if (lower_only == 2 && blockIdx.x == 0 && blockIdx.y == 0) { return; }
        const int tid = threadIdx.x;
        const int lane = tid & 31, warp = tid >> 5;
        const int wm = warp & 1, wn = warp >> 1;
        const int g = lane >> 2, kq = lane & 3;
        double acc[8][4][2];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = acc[mi][ni][1] = 0.0; }
        }
        const long num_kt = (k + kGemmBK - 1) / kGemmBK;
        const bool va = !A_KC && (lda & 1) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0;  // row0 / col0 are multiples of 128
        const bool vb = !B_KC && (ldb & 1) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0;
#pragma unroll
        for (int st = 0; st < kGemmStages - 1; ++st) {
            if (st < num_kt) {
                GemmCpAsyncTile<A_KC>(as + st * kSlab, a, lda, row0, m, st * kGemmBK, k, tid, va);
                GemmCpAsyncTile<B_KC>(bs + st * kSlab, b, ldb, col0, n, st * kGemmBK, k, tid, vb);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (long kt = 0; kt < num_kt; ++kt) {
            asm volatile("cp.async.wait_group %0;" ::"n"(kGemmStages - 2) : "memory");  // slab kt has landed (for this thread)
            __syncthreads();                                                              // ... for everybody; slab kt - 1 is consumed
            const long nxt = kt + kGemmStages - 1;
            if (nxt < num_kt) {
                const int st = static_cast<int>(nxt % kGemmStages);
                GemmCpAsyncTile<A_KC>(as + st * kSlab, a, lda, row0, m, nxt * kGemmBK, k, tid, va);
                GemmCpAsyncTile<B_KC>(bs + st * kSlab, b, ldb, col0, n, nxt * kGemmBK, k, tid, vb);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            const int cur = static_cast<int>(kt % kGemmStages);
            SlabMma(acc, as + cur * kSlab, bs + cur * kSlab, wm, wn, lane);
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const long col = col0 + 32 * wn + 8 * ni + 2 * kq + e;
                if (col >= n) { continue; }
#pragma unroll
                for (int mi = 0; mi < 8; ++mi) {
                    const long row = row0 + 64 * wm + 8 * mi + g;
                    if (row >= m || (lower_only && row < col)) { continue; }
                    double *dst = c + row + col * ldc;
                    const double prev = beta == 0.0 ? 0.0 : beta * (*dst);
                    *dst = alpha * acc[mi][ni][e] + prev;
                }
            }
        }
    }

    // ---- the same GEMM with 16 warps per CTA (32 x 32 warp tiles) --------------------------------------------------------
    // The 8-warp kernel keeps 2 warps per scheduler and its tensor pipe is 73.5 % active (ncu, rank-512 update): the rest are
    // fixed-latency and scoreboard stalls that nothing covers.  512 threads halve the accumulators per thread (32 doubles, <= 128
    // registers) and give every scheduler 4 warps; the operand fragments are read twice as often from shared memory (8 LDS.64
    // per 16 DMMAs instead of 12 per 32), which the shared-memory pipe has room for.
    constexpr int kGemmThreads512 = 512;

    template<bool K_CONTIG>
    __device__ __forceinline__ void
    GemmLoadTile512(const double *__restrict__ src, const long ld, const long o0, const long o_lim, const long k0, const long k_lim, const int tid, double (&reg)[4]) {
        if (!K_CONTIG) {
            const int o = (tid & 31) * 4;
            const long k = k0 + (tid >> 5);  // 0 .. 15
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long oo = o0 + o + j;
                reg[j] = (oo < o_lim && k < k_lim) ? src[oo + k * ld] : 0.0;
            }
        } else {
            const long k = k0 + (tid & 15);
            const int o = tid >> 4;  // 0 .. 31
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long oo = o0 + o + 32 * i;
                reg[i] = (oo < o_lim && k < k_lim) ? src[k + oo * ld] : 0.0;
            }
        }
    }

    template<bool K_CONTIG>
    __device__ __forceinline__ void
    GemmStoreTile512(double *__restrict__ dst /* [BK][128 + pad] */, const int tid, const double (&reg)[4]) {
        constexpr int kLd = kGemmBM + kGemmPad;
        if (!K_CONTIG) {
            Store4(dst + (tid >> 5) * kLd + (tid & 31) * 4, reg);
        } else {
            const int kk = tid & 15;
            const int o = tid >> 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) { dst[kk * kLd + o + 32 * i] = reg[i]; }
        }
    }

    template<bool A_KC, bool B_KC>
    __global__ void __launch_bounds__(kGemmThreads512, 1)
    GemmKernelDmma512(const long m, const long n, const long k, const double alpha, const double *__restrict__ a, const long lda, const double *__restrict__ b, const long ldb, const double beta,
                      double *__restrict__ c, const long ldc, const int lower_only) {
        constexpr int kLd = kGemmBM + kGemmPad;
        extern __shared__ __align__(16) unsigned char smem_raw[];
        double *as = reinterpret_cast<double *>(smem_raw);
        double *bs = as + 2 * kGemmBK * kLd;
        const long row0 = static_cast<long>(blockIdx.x) * kGemmBM;
        const long col0 = static_cast<long>(blockIdx.y) * kGemmBN;
        if (lower_only && col0 > row0 + kGemmBM - 1) { return; }
        if (lower_only == 2 && blockIdx.x == 0 && blockIdx.y == 0) { return; }
        const int tid = threadIdx.x;
        const int lane = tid & 31, warp = tid >> 5;
        const int wm = warp & 3, wn = warp >> 2;
        const int g = lane >> 2, kq = lane & 3;
        double acc[4][4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = acc[mi][ni][1] = 0.0; }
        }
        double ra[4], rb[4];
        GemmLoadTile512<A_KC>(a, lda, row0, m, 0, k, tid, ra);
        GemmLoadTile512<B_KC>(b, ldb, col0, n, 0, k, tid, rb);
        GemmStoreTile512<A_KC>(as, tid, ra);
        GemmStoreTile512<B_KC>(bs, tid, rb);
        __syncthreads();
        const long num_kt = (k + kGemmBK - 1) / kGemmBK;
        for (long kt = 0; kt < num_kt; ++kt) {
            const int cur = static_cast<int>(kt & 1);
            if (kt + 1 < num_kt) {
                GemmLoadTile512<A_KC>(a, lda, row0, m, (kt + 1) * kGemmBK, k, tid, ra);
                GemmLoadTile512<B_KC>(b, ldb, col0, n, (kt + 1) * kGemmBK, k, tid, rb);
            }
            const double *at = as + cur * kGemmBK * kLd;
            const double *bt = bs + cur * kGemmBK * kLd;
#pragma unroll
            for (int k4 = 0; k4 < kGemmBK / 4; ++k4) {
                const double *ap = at + (4 * k4 + kq) * kLd + 32 * wm + g;
                const double *bp = bt + (4 * k4 + kq) * kLd + 32 * wn + g;
                double av[4], bv[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) { av[mi] = ap[8 * mi]; }
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) { bv[ni] = bp[8 * ni]; }
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) { Dmma884(acc[mi][ni], av[mi], bv[ni]); }
                }
            }
            if (kt + 1 < num_kt) {
                GemmStoreTile512<A_KC>(as + (cur ^ 1) * kGemmBK * kLd, tid, ra);
                GemmStoreTile512<B_KC>(bs + (cur ^ 1) * kGemmBK * kLd, tid, rb);
            }
            __syncthreads();
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const long col = col0 + 32 * wn + 8 * ni + 2 * kq + e;
                if (col >= n) { continue; }
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    const long row = row0 + 32 * wm + 8 * mi + g;
                    if (row >= m || (lower_only && row < col)) { continue; }
                    double *dst = c + row + col * ldc;
                    const double prev = beta == 0.0 ? 0.0 : beta * (*dst);
                    *dst = alpha * acc[mi][ni][e] + prev;
                }
            }
        }
    }

    // ---- the 16-warp GEMM fed by the TMA engine ----------------------------------------------------------------------------
    // C = alpha A B^T + beta C with both operands outer-contiguous (the SYRK / GEMM trailing updates of the blocked Cholesky and the
    // panel updates of the triangular solves): column kk of a 128 x 16 operand slab is 1 KB of contiguous, 16-byte aligned HBM, so a
    // slab is 16 one-dimensional bulk copies (cp.async.bulk.shared::cluster.global, no tensor map needed) that land in the padded
    // [k][132] layout SlabMma-style fragment loads want, and complete on the stage's mbarrier (expect_tx = the bytes of both slabs).
    // Warp 0 issues the 32 copies of a stage, one per lane; nobody spends LDG / STS issue slots or staging registers on operands, and
    // kGemmTmaStages slabs are in flight instead of one.  A stage is reused after the CTA barrier that ends its k-step (measured:
    // per-stage `empty` mbarriers instead of that barrier, so that the warps may drift apart, made Potrf n = 16384 slower - 65.1 ms
    // against 59.8 ms - the 16 warps share one slab and run best in step).
    // Preconditions (checked by Gemm(), else the register-staged kernel runs): k % 16 == 0, m and n even, lda and ldb even,
    // 16-byte aligned operands.  Rows past the matrix edge are never copied: they only feed accumulators that are not stored.
    constexpr int kGemmTmaStages = 4;

    __device__ __forceinline__ void
    MbarWaitParity(const uint32_t bar, const uint32_t parity) {
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        } while (done == 0);
    }

    __global__ void __launch_bounds__(kGemmThreads512, 1)
    GemmKernelDmmaTma(const long m, const long n, const long k, const double alpha, const double *__restrict__ a, const long lda, const double *__restrict__ b, const long ldb, const double beta,
                      double *__restrict__ c, const long ldc, const int lower_only) {
        constexpr int kLd = kGemmBM + kGemmPad;
        constexpr int kSlab = kGemmBK * kLd;
        extern __shared__ __align__(128) unsigned char smem_raw[];
        double *as = reinterpret_cast<double *>(smem_raw);  // [stages][BK][kLd]
        double *bs = as + kGemmTmaStages * kSlab;
        const uint32_t bars = static_cast<uint32_t>(__cvta_generic_to_shared(bs + kGemmTmaStages * kSlab));  // one 8-byte mbarrier per stage
        const long row0 = static_cast<long>(blockIdx.x) * kGemmBM;
        const long col0 = static_cast<long>(blockIdx.y) * kGemmBN;
        if (lower_only && col0 > row0 + kGemmBM - 1) { return; }
        if (lower_only == 2 && blockIdx.x == 0 && blockIdx.y == 0) { return; }
        const int tid = threadIdx.x;
        const int lane = tid & 31, warp = tid >> 5;
        const int wm = warp & 3, wn = warp >> 2;
        const int g = lane >> 2, kq = lane & 3;
        if (tid == 0) {
#pragma unroll
            for (int st = 0; st < kGemmTmaStages; ++st) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * st) : "memory"); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        const uint32_t a_bytes = static_cast<uint32_t>(m - row0 < kGemmBM ? m - row0 : kGemmBM) * 8u;
        const uint32_t b_bytes = static_cast<uint32_t>(n - col0 < kGemmBN ? n - col0 : kGemmBN) * 8u;
        const long num_kt = k / kGemmBK;
        // lane l < 16: column l of the A slab, lane 16 + l: column l of the B slab
        const int kk = lane & 15;
        const double *src = lane < 16 ? a + row0 + kk * lda : b + col0 + kk * ldb;
        const long src_step = kGemmBK * (lane < 16 ? lda : ldb);
        const uint32_t dst0 = static_cast<uint32_t>(__cvta_generic_to_shared((lane < 16 ? as : bs) + kk * kLd));
        const uint32_t my_bytes = lane < 16 ? a_bytes : b_bytes;
        auto issue = [&](const long kt) {  // warp 0, all lanes
            const int st = static_cast<int>(kt % kGemmTmaStages);
            const uint32_t bar = bars + 8 * st;
            if (lane == 0) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kGemmBK * (a_bytes + b_bytes)) : "memory"); }
            __syncwarp();
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst0 + st * kSlab * 8), "l"(src + kt * src_step), "r"(my_bytes), "r"(bar)
                         : "memory");
        };
        if (warp == 0) {
            for (long kt = 0; kt < kGemmTmaStages && kt < num_kt; ++kt) { issue(kt); }
        }
        // The epilogue reads the C tile once, at the very end, and nothing hides that DRAM latency (ncu: C is 70 % of the DRAM reads
        // of a rank-512 update).  One bulk L2 prefetch per column now, so that the tile waits in L2 when the k loop is done.
        if (beta != 0.0 && tid >= 32 && tid < 32 + kGemmBN && (ldc & 1) == 0 && (reinterpret_cast<uintptr_t>(c) & 15) == 0) {
            const long col = col0 + (tid - 32);
            if (col < n) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(c + row0 + col * ldc), "r"(a_bytes) : "memory"); }
        }
        double acc[4][4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = acc[mi][ni][1] = 0.0; }
        }
        for (long kt = 0; kt < num_kt; ++kt) {
            const int st = static_cast<int>(kt % kGemmTmaStages);
            MbarWaitParity(bars + 8 * st, static_cast<uint32_t>(kt / kGemmTmaStages) & 1u);
            const double *at = as + st * kSlab;
            const double *bt = bs + st * kSlab;
#pragma unroll
            for (int k4 = 0; k4 < kGemmBK / 4; ++k4) {
                const double *ap = at + (4 * k4 + kq) * kLd + 32 * wm + g;
                const double *bp = bt + (4 * k4 + kq) * kLd + 32 * wn + g;
                double av[4], bv[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) { av[mi] = ap[8 * mi]; }
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) { bv[ni] = bp[8 * ni]; }
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) { Dmma884(acc[mi][ni], av[mi], bv[ni]); }
                }
            }
            __syncthreads();  // every warp is done with stage st
            if (warp == 0 && kt + kGemmTmaStages < num_kt) { issue(kt + kGemmTmaStages); }
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const long col = col0 + 32 * wn + 8 * ni + 2 * kq + e;
                if (col >= n) { continue; }
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    const long row = row0 + 32 * wm + 8 * mi + g;
                    if (row >= m || (lower_only && row < col)) { continue; }
                    double *dst = c + row + col * ldc;
                    const double prev = beta == 0.0 ? 0.0 : beta * (*dst);
                    *dst = alpha * acc[mi][ni][e] + prev;
                }
            }
        }
    }

    // ERL_GP_DENSE_TMA=0: the register-staged 16-warp kernel for every shape (A/B)
    static bool
    GemmUseTma() {
        static const char *env = std::getenv("ERL_GP_DENSE_TMA");
        return env == nullptr || std::atoi(env) != 0;
    }

    static bool
    GemmUse512() {
        static const bool off = std::getenv("ERL_GP_DENSE_256") != nullptr;  // A/B: the 8-warp kernel (Potrf n = 16384: 63.9 ms vs 61.9 ms)
        return !off;
    }

    // ERL_GP_DENSE_CP_ASYNC=1 selects the cp.async kernel.  Measured on the B200 at n = 16384 (Potrf, 63.9 ms with the
    // register-staged kernel): 69.7 ms with 8-byte copies, 66.5 ms with 16-byte copies - slower, so it is off by default.
    static bool
    GemmUseCpAsync() {
        static const bool on = std::getenv("ERL_GP_DENSE_CP_ASYNC") != nullptr;
        return on;
    }

    template<typename T, bool A_KC, bool B_KC>
    struct GemmSelect {
        static auto
        Get() {
            return GemmKernel<T, A_KC, B_KC>;
        }
    };
    template<bool A_KC, bool B_KC>
    struct GemmSelect<double, A_KC, B_KC> {
        static auto
        Get() {
            static const bool fma = std::getenv("ERL_GP_DENSE_FMA") != nullptr;      // A/B measurements of the DFMA loop
            return fma ? GemmKernel<double, A_KC, B_KC> : (GemmUseCpAsync() ? GemmKernelDmmaAsync<A_KC, B_KC> : (GemmUse512() ? GemmKernelDmma512<A_KC, B_KC> : GemmKernelDmma<A_KC, B_KC>));
        }
    };

    template<typename T>
    int
    Gemm(Context *ctx, int op_a, int op_b, long m, long n, long k, T alpha, const T *a, long lda, const T *b, long ldb, T beta, T *c, long ldc, int lower_only) {
        if (m <= 0 || n <= 0) { return ERL_GP_STATUS_OK; }
        const dim3 grid(static_cast<unsigned>(CeilDiv(m, kGemmBM)), static_cast<unsigned>(CeilDiv(n, kGemmBN)));
        // two stages of two operand slabs; the cp.async FP64 kernel uses kGemmStages
        const size_t smem = sizeof(T) * ((sizeof(T) == 8 && GemmUseCpAsync()) ? 2 * kGemmStages : 4) * kGemmBK * (kGemmBM + kGemmPad);
        // op(A)(m,k): N -> a[m + k lda] (outer contiguous), T -> a[k + m lda] (k contiguous)
        // op(B)(k,n): N -> b[k + n ldb] (k contiguous),     T -> b[n + k ldb] (outer contiguous)
#define ERL_GP_GEMM_LAUNCH(AKC, BKC)                                                                                          \
    {                                                                                                                         \
        auto kern = GemmSelect<T, AKC, BKC>::Get();                                                                               \
        ERL_GP_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
        kern<<<grid, (sizeof(T) == 8 && GemmUse512() && !GemmUseCpAsync() && std::getenv("ERL_GP_DENSE_FMA") == nullptr) ? kGemmThreads512 : kGemmThreads, smem, ctx->stream>>>(m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, lower_only);    \
    }
        if constexpr (sizeof(T) == 8) {
            if (op_a == kOpN && op_b == kOpT && GemmUseTma() && GemmUse512() && !GemmUseCpAsync() && std::getenv("ERL_GP_DENSE_FMA") == nullptr && k > 0 && (k % kGemmBK) == 0 &&
                ((m | n | lda | ldb) & 1) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0) {
                const size_t smem_tma = sizeof(double) * 2 * kGemmTmaStages * kGemmBK * (kGemmBM + kGemmPad) + 8 * kGemmTmaStages;
                ERL_GP_CUDA_OK(ctx, cudaFuncSetAttribute(GemmKernelDmmaTma, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_tma)));
                GemmKernelDmmaTma<<<grid, kGemmThreads512, smem_tma, ctx->stream>>>(m, n, k, alpha, reinterpret_cast<const double *>(a), lda, reinterpret_cast<const double *>(b), ldb, beta,
                                                                                   reinterpret_cast<double *>(c), ldc, lower_only);
                ctx->launches += 1;
                ERL_GP_CUDA_OK(ctx, cudaGetLastError());
                return ERL_GP_STATUS_OK;
            }
        }
        if (op_a == kOpN && op_b == kOpT) {
            ERL_GP_GEMM_LAUNCH(false, false)
        } else if (op_a == kOpN && op_b == kOpN) {
            ERL_GP_GEMM_LAUNCH(false, true)
        } else if (op_a == kOpT && op_b == kOpN) {
            ERL_GP_GEMM_LAUNCH(true, true)
        } else {
            ERL_GP_GEMM_LAUNCH(true, false)
        }
#undef ERL_GP_GEMM_LAUNCH
        ctx->launches += 1;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        return ERL_GP_STATUS_OK;
    }

    // =========================================================================================
    // 128 x 128 diagonal block: Cholesky + inverse in one CTA
    // =========================================================================================
    constexpr int kDiagBlocks = kPanel / kNB;  // 8

    // V = L^-1 (identity right-hand side), register-resident right-looking blocked substitution.
    // Thread (tr, tc) owns rows tr + 16 m and columns col0 + QPT * tc + j.
    template<typename T>
    __device__ void
    InverseColumns(const T *lp, const T *dinv, T *r_buf, T *s_buf, const int nblk, const int col0, T *__restrict__ out /* 128 x 128 col-major */) {
        using Smem = BatchSmem<T, 1, kDiagBlocks>;
        constexpr int kQpt = Smem::kQpt;
        constexpr int kTq = Smem::kTq;
        constexpr int kLd = DinvLd<T>::value;
        const int tid = threadIdx.x;
        const int tr = tid & 15;
        const int tc = tid >> 4;
        const int qbase = tc * kQpt;
        T v[kDiagBlocks][kQpt];
#pragma unroll
        for (int m = 0; m < kDiagBlocks; ++m) {
#pragma unroll
            for (int j = 0; j < kQpt; ++j) { v[m][j] = (tr + kNB * m == col0 + qbase + j) ? T(1) : T(0); }
        }
        for (int kb = col0 / kNB; kb < nblk; ++kb) {  // rows above col0 of these columns of the inverse are zero
            T rk[kQpt];
#pragma unroll
            for (int m = 0; m < kDiagBlocks; ++m) {
                if (m == kb) {
#pragma unroll
                    for (int j = 0; j < kQpt; ++j) { rk[j] = v[m][j]; }
                }
            }
#pragma unroll
            for (int j = 0; j < kQpt; j += 4) {
                T t4[4] = {rk[j], rk[j + 1], rk[j + 2], rk[j + 3]};
                Store4(r_buf + tr * kTq + qbase + j, t4);
            }
            __syncthreads();
            T res[kQpt];
#pragma unroll
            for (int j = 0; j < kQpt; ++j) { res[j] = T(0); }
            const T *dk = dinv + kb * kNB * kLd + tr * kLd;
#pragma unroll
            for (int p4 = 0; p4 < kNB; p4 += 4) {
                T d4[4];
                Load4(dk + p4, d4);
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
#pragma unroll
                    for (int j = 0; j < kQpt; j += 4) {
                        T r4[4];
                        Load4(r_buf + (p4 + pp) * kTq + qbase + j, r4);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) { res[j + jj] += d4[pp] * r4[jj]; }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kQpt; j += 4) {
                T t4[4] = {res[j], res[j + 1], res[j + 2], res[j + 3]};
                Store4(s_buf + tr * kTq + qbase + j, t4);
            }
#pragma unroll
            for (int j = 0; j < kQpt; ++j) { out[(kb * kNB + tr) + static_cast<long>(col0 + qbase + j) * kPanel] = res[j]; }
            __syncthreads();
            if (kb + 1 < nblk) {
#pragma unroll
                for (int p2 = 0; p2 < kNB; p2 += 2) {
                    T s2[2][kQpt];
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp) {
#pragma unroll
                        for (int j = 0; j < kQpt; j += 4) {
                            T t4[4];
                            Load4(s_buf + (p2 + pp) * kTq + qbase + j, t4);
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) { s2[pp][j + jj] = t4[jj]; }
                        }
                    }
#pragma unroll
                    for (int m = 1; m < kDiagBlocks; ++m) {
                        if (m > kb && m < nblk) {
                            const T *lrow = lp + LowerBlock(m, kb) + tr + kNB * p2;
                            const T l0 = lrow[0];
                            const T l1 = lrow[kNB];
#pragma unroll
                            for (int j = 0; j < kQpt; ++j) {
                                v[m][j] -= l0 * s2[0][j];
                                v[m][j] -= l1 * s2[1][j];
                            }
                        }
                    }
                }
            }
        }
    }

    // a: nk x nk block (lower triangle read, ld).  On exit a holds L11 (strict upper of the block zeroed),
    // linv the 128 x 128 inverse (identity padded).  info: set to col_offset + failing column if not SPD.
    template<typename T>
    __global__ void __launch_bounds__(kBatchThreads, 1)
    DiagFactorKernel(T *__restrict__ a, const long ld, const int nk, T *__restrict__ linv, int *__restrict__ info, const int col_offset) {
        using Smem = BatchSmem<T, 1, kDiagBlocks>;
        extern __shared__ __align__(16) unsigned char smem_raw[];
        T *smem = reinterpret_cast<T *>(smem_raw);
        T *lp = smem + Smem::kLp;
        T *dinv = smem + Smem::kDinv;
        T *r_buf = smem + Smem::kR;
        T *s_buf = smem + Smem::kS;
        int *s_fail = reinterpret_cast<int *>(smem + Smem::kEnd);
        constexpr int kLd = DinvLd<T>::value;
        const int tid = threadIdx.x;
        const int warp = tid >> 5;
        const int lane = tid & 31;
        const int nblk = (nk + kNB - 1) / kNB;
        const int npad = nblk * kNB;
        if (tid == 0) { *s_fail = 0; }
        // 8 loads in flight per thread (one load per loop trip left the single CTA waiting ~1 us of L2 / HBM latency per
        // column: 17 % of the kernel); a warp reads 32 consecutive rows of one column
        for (int i0 = 0; i0 < kPanel * kPanel / kBatchThreads; i0 += 8) {
            T val[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = tid + kBatchThreads * (i0 + u);
                const int r = e % kPanel, c = e / kPanel;
                if (r < nk && c < nk) {
                    val[u] = r >= c ? a[r + static_cast<long>(c) * ld] : T(0);
                } else {
                    val[u] = r == c ? T(1) : T(0);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = tid + kBatchThreads * (i0 + u);
                const int r = e % kPanel, c = e / kPanel;
                if ((r >> 4) >= (c >> 4) && r < npad) { lp[LowerBlock(r >> 4, c >> 4) + (r & 15) + kNB * (c & 15)] = val[u]; }
            }
        }
        __syncthreads();
        CholeskySmem(lp, dinv, nblk, s_fail);
        __syncthreads();
        if (*s_fail != 0) {
            if (tid == 0 && *info == 0) { *info = col_offset + *s_fail; }
        }
        // the last diagonal block's inverse is not produced inside CholeskySmem's loop when it breaks early: it is
        // (DiagInverse runs before the break), so dinv is complete here.
        for (int c = warp; c < nk; c += kBatchThreads / 32) {
            for (int r = (c & ~15) + lane; r < nk; r += 32) { a[r + static_cast<long>(c) * ld] = r >= c ? lp[LowerBlock(r >> 4, c >> 4) + (r & 15) + kNB * (c & 15)] : T(0); }
        }
        // zero + identity-pad the inverse, then fill the columns of the active blocks
        for (int e = tid; e < kPanel * kPanel; e += kBatchThreads) { linv[e] = (e % kPanel == e / kPanel && e / kPanel >= npad) ? T(1) : T(0); }
        __syncthreads();
        constexpr int kTq = Smem::kTq;
        for (int col0 = 0; col0 < npad; col0 += kTq) { InverseColumns<T>(lp, dinv, r_buf, s_buf, nblk, col0, linv); }
        (void) kLd;
    }

    template<typename T>
    __global__ void
    CopyLowerKernel(const long n, const T *__restrict__ k, const long ld_k, T *__restrict__ l, const long ld_l) {
        const long r = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
        const long c = blockIdx.y;
        if (r < n) { l[r + c * ld_l] = r >= c ? k[r + c * ld_k] : T(0); }
    }

    template<typename T>
    int
    CopyLower(Context *ctx, long n, const T *k, long ld_k, T *l, long ld_l) {
        const dim3 grid(static_cast<unsigned>(CeilDiv(n, 256)), static_cast<unsigned>(n));
        CopyLowerKernel<T><<<grid, 256, 0, ctx->stream>>>(n, k, ld_k, l, ld_l);
        ctx->launches += 1;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        return ERL_GP_STATUS_OK;
    }

    // C[0:nt, 0:nt] (lower triangle, nt <= 128) -= A[0:nt, 0:k] A[0:nt, 0:k]^T, spread over up to 10 CTAs of 32 x 32 outputs.
    // This is the tile that the next diagonal factorisation waits for: as one 128 x 128 GEMM tile it is a single CTA on
    // a single SM (30 us at k = 128, 89 us at k = 512); cut in 32 x 32 pieces it takes a few microseconds.
    template<typename T>
    __global__ void __launch_bounds__(256)
    SyrkTileKernel(const int nt, const long k, const T *__restrict__ a, const long lda, T *__restrict__ c, const long ldc) {
        constexpr int kSub = 32, kChunk = 64, kLdS = kSub + 2;
        __shared__ __align__(16) T as[kChunk][kLdS];
        __shared__ __align__(16) T bs[kChunk][kLdS];
        const int r0 = kSub * blockIdx.x, c0 = kSub * blockIdx.y;
        if (c0 > r0 || r0 >= nt) { return; }
        const int tid = threadIdx.x;
        const int tx = tid & 15, ty = tid >> 4;
        T acc[2][2] = {{T(0), T(0)}, {T(0), T(0)}};
        for (long k0 = 0; k0 < k; k0 += kChunk) {
            T va[8], vb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = tid + 256 * u;
                const int rr = e & 31;
                const long kk = k0 + (e >> 5);
                va[u] = (r0 + rr < nt && kk < k) ? a[(r0 + rr) + kk * lda] : T(0);
                vb[u] = (c0 + rr < nt && kk < k) ? a[(c0 + rr) + kk * lda] : T(0);
            }
            __syncthreads();  // the previous chunk has been consumed
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = tid + 256 * u;
                as[e >> 5][e & 31] = va[u];
                bs[e >> 5][e & 31] = vb[u];
            }
            __syncthreads();
#pragma unroll 16
            for (int kk = 0; kk < kChunk; ++kk) {
                T a2[2], b2[2];
                Load2(&as[kk][2 * tx], a2);
                Load2(&bs[kk][2 * ty], b2);
                acc[0][0] += a2[0] * b2[0];
                acc[0][1] += a2[0] * b2[1];
                acc[1][0] += a2[1] * b2[0];
                acc[1][1] += a2[1] * b2[1];
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = r0 + 2 * tx + i, col = c0 + 2 * ty + j;
                if (row < nt && col < nt && row >= col) { c[row + static_cast<long>(col) * ldc] -= acc[i][j]; }
            }
        }
    }

    // L11 (in place) and its inverse for one 128 x 128 diagonal block.  FP64: the DMMA kernel of erl_gp_rowgp64.cuh (round 2; the
    // generic shared-memory kernel above with ERL_GP_DIAG_LEGACY=1); FP32: the generic kernel.
    template<typename T>
    static void
    LaunchDiagFactor(cudaStream_t stream, T *a, const long ld, const int nk, T *linv, int *info, const int col_offset) {
        using Smem = BatchSmem<T, 1, kDiagBlocks>;
        if constexpr (sizeof(T) == 8) {
            static const bool legacy = std::getenv("ERL_GP_DIAG_LEGACY") != nullptr;
            static const bool attr = [] {
                cudaFuncSetAttribute(rowgp64::DiagFactor64Kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(rowgp64::DiagFactor64SmemBytes()));
                return true;
            }();
            (void) attr;
            if (!legacy) {
                rowgp64::DiagFactor64Kernel<8><<<1, rowgp64::kThreads, rowgp64::DiagFactor64SmemBytes(), stream>>>(a, ld, nk, linv, info, col_offset);
                return;
            }
        }
        DiagFactorKernel<T><<<1, kBatchThreads, Smem::kBytes, stream>>>(a, ld, nk, linv, info, col_offset);
    }

    // Look-ahead of one diagonal block on the stream `la`: once `chain` has produced the k columns at `a` (rows of the
    // block), update the nt x nt diagonal tile `c` with them and factor it; ev_la_done fires when L11 and its inverse exist.
    template<typename T>
    static int
    LookaheadDiag(Context *ctx, cudaStream_t chain, cudaStream_t la, const int nt, const long k, const T *a, const long lda, T *c, const long ldc, T *linv_k, int *info, const long col) {
        using Smem = BatchSmem<T, 1, kDiagBlocks>;
        ERL_GP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_la_start, chain));
        ERL_GP_CUDA_OK(ctx, cudaStreamWaitEvent(la, ctx->ev_la_start, 0));
        SyrkTileKernel<T><<<dim3(4, 4), 256, 0, la>>>(nt, k, a, lda, c, ldc);
        LaunchDiagFactor<T>(la, c, ldc, nt, linv_k, info, static_cast<int>(col));
        ctx->launches += 2;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        ERL_GP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_la_done, la));
        return ERL_GP_STATUS_OK;
    }

    // Factor the block column [j0, j1) (rows j0 .. n) right-looking with 128-column panels on ctx->stream:
    // DiagFactorKernel (L11 and its inverse), panel solve L21 <- A21 L11^-T as an IN-PLACE GEMM against the inverse (one
    // column tile: every CTA has consumed its 128 rows of A21 before its epilogue overwrites them), rank-128 update of the
    // columns of the block column to the right of the panel with L21 itself as the operand.
    // Panel look-ahead: the diagonal tile of the next panel is updated and factored on ctx->side_stream2 (LookaheadDiag)
    // while ctx->stream runs the rest of the rank-128 update, so the chain per panel is diag + solve + a few microseconds
    // instead of diag + solve + whole update.  first_diag_pending: the caller already did that for the first panel.
    template<typename T>
    static int
    FactorBlockColumn(Context *ctx, const long n, const long j0, const long j1, T *l, const long ld, T *linv, int *info, const bool first_diag_pending) {
        using Smem = BatchSmem<T, 1, kDiagBlocks>;
        static const bool no_panel_la = std::getenv("ERL_GP_POTRF_NO_PANEL_LOOKAHEAD") != nullptr;  // A/B measurements
        cudaStream_t chain = ctx->stream;
        cudaStream_t la = ctx->side_stream2;
        bool diag_done = first_diag_pending;  // the diagonal block of the panel about to be processed is factored on `la`
        int rc = ERL_GP_STATUS_OK;
        for (long k0 = j0; k0 < j1; k0 += kPanel) {
            const long nk = n - k0 < kPanel ? n - k0 : kPanel;
            T *linv_k = linv + (k0 / kPanel) * kPanel * kPanel;
            if (diag_done) {
                ERL_GP_CUDA_OK(ctx, cudaStreamWaitEvent(chain, ctx->ev_la_done, 0));
                diag_done = false;
            } else {
                LaunchDiagFactor<T>(chain, l + k0 + k0 * ld, ld, static_cast<int>(nk), linv_k, info, static_cast<int>(k0));
                ctx->launches += 1;
                ERL_GP_CUDA_OK(ctx, cudaGetLastError());
            }
            const long k1 = k0 + nk;
            const long m = n - k1;
            if (m <= 0) { break; }
            T *l21 = l + k1 + k0 * ld;
            rc = Gemm<T>(ctx, kOpN, kOpT, m, nk, nk, T(1), l21, ld, linv_k, kPanel, T(0), l21, ld, 0);
            if (rc != ERL_GP_STATUS_OK) { return rc; }
            const long rest = j1 - k1;  // columns of the block column still to the right of this panel
            if (rest <= 0) { break; }
            T *c = l + k1 + k1 * ld;
            int skip_first_tile = 1;
            if (!no_panel_la) {
                const long next_nk = m < kPanel ? m : kPanel;
                rc = LookaheadDiag<T>(ctx, chain, la, static_cast<int>(next_nk), nk, l21, ld, c, ld, linv + (k1 / kPanel) * kPanel * kPanel, info, k1);
                if (rc != ERL_GP_STATUS_OK) { return rc; }
                diag_done = true;
                skip_first_tile = 2;
            }
            rc = Gemm<T>(ctx, kOpN, kOpT, m, rest, nk, T(-1), l21, ld, l21, ld, T(1), c, ld, skip_first_tile);  // 2: all lower tiles but (0, 0)
            if (rc != ERL_GP_STATUS_OK) { return rc; }
        }
        if (diag_done) { ERL_GP_CUDA_OK(ctx, cudaStreamWaitEvent(chain, ctx->ev_la_done, 0)); }
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    int
    Potrf(Context *ctx, long n, T *l, long ld, T *linv, T *panel, int *info) {
        using Smem = BatchSmem<T, 1, kDiagBlocks>;
        ERL_GP_CUDA_OK(ctx, cudaFuncSetAttribute(DiagFactorKernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(Smem::kBytes)));
        ERL_GP_CUDA_OK(ctx, cudaMemsetAsync(info, 0, sizeof(int), ctx->stream));
        // Two-level blocking: a block column of kOuter columns is factored with 128-column panels whose rank-128 updates
        // stay inside the block column; everything to its right gets ONE rank-kOuter update per block column (4x less
        // read-modify-write traffic on the trailing matrix and 4x longer reduction loops than a rank-128 SYRK per panel:
        // the rank-512 SYRK runs at 26.5 of 37 TFLOP/s, the rank-128 one at 19).
        // Look-ahead over block columns: the panel chain of a block column is a sequence of small latency-bound kernels
        // (4 x (one-CTA diagonal factorisation + panel solve + rank-128 update), ~0.7 ms) that left the GPU idle for a third
        // of the factorisation.  So, once block column b is final, the update of block column b + 1 and its panel chain run
        // on a high-priority side stream while the main stream applies the rank-kOuter update to the columns right of block
        // column b + 1; the two join before the next step.  The first diagonal block of block column b + 1 is itself looked
        // ahead (LookaheadDiag) while the side stream updates the rest of that block column.
        constexpr long kOuter = 4 * kPanel;
        if (ctx->side_stream == nullptr) {
            // highest priority: the chain's small kernels must get the first SM slots that free up while the bulk update still
            // has thousands of CTAs queued (with equal priorities they only start when the bulk kernel has drained)
            int prio_lo = 0, prio_hi = 0;
            ERL_GP_CUDA_OK(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            ERL_GP_CUDA_OK(ctx, cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, prio_hi));
            ERL_GP_CUDA_OK(ctx, cudaStreamCreateWithPriority(&ctx->side_stream2, cudaStreamNonBlocking, prio_hi));
            for (cudaEvent_t *ev: {&ctx->ev_panel, &ctx->ev_diag, &ctx->ev_la_start, &ctx->ev_la_done}) { ERL_GP_CUDA_OK(ctx, cudaEventCreateWithFlags(ev, cudaEventDisableTiming)); }
        }
        (void) panel;  // workspace of the former out-of-place panel solve
        cudaStream_t main_stream = ctx->stream;
        cudaStream_t side = ctx->side_stream;
        static const bool no_lookahead = std::getenv("ERL_GP_POTRF_NO_LOOKAHEAD") != nullptr;  // A/B measurements
        int rc = FactorBlockColumn<T>(ctx, n, 0, kOuter < n ? kOuter : n, l, ld, linv, info, false);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        for (long j0 = 0; j0 + kOuter < n; j0 += kOuter) {
            // block column [j0, j1) is final and visible to the main stream
            const long j1 = j0 + kOuter;
            const long m = n - j1;                     // rows / columns of the trailing matrix
            const long w = kOuter < m ? kOuter : m;    // width of the next block column
            const T *a1 = l + j1 + j0 * ld;            // L[j1:, j0:j1]
            T *c1 = l + j1 + j1 * ld;
            const bool fork = m > w && !no_lookahead;
            cudaStream_t chain = fork ? side : main_stream;
            if (fork) {
                ERL_GP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_panel, main_stream));
                ERL_GP_CUDA_OK(ctx, cudaStreamWaitEvent(side, ctx->ev_panel, 0));
            }
            // next block column: A[j1:, j1:j1+w] -= L[j1:, j0:j1] L[j1:j1+w, j0:j1]^T (lower tiles; its diagonal tile (0, 0) on
            // the look-ahead stream, followed by the first diagonal factorisation), then its panel chain
            rc = LookaheadDiag<T>(ctx, chain, ctx->side_stream2, static_cast<int>(m < kPanel ? m : kPanel), kOuter, a1, ld, c1, ld, linv + (j1 / kPanel) * kPanel * kPanel, info, j1);
            if (rc != ERL_GP_STATUS_OK) { return rc; }
            ctx->stream = chain;
            rc = Gemm<T>(ctx, kOpN, kOpT, m, w, kOuter, T(-1), a1, ld, a1, ld, T(1), c1, ld, 2);
            if (rc == ERL_GP_STATUS_OK) { rc = FactorBlockColumn<T>(ctx, n, j1, j1 + w, l, ld, linv, info, true); }
            if (fork && rc == ERL_GP_STATUS_OK && cudaEventRecord(ctx->ev_diag, side) != cudaSuccess) { rc = ERL_GP_STATUS_CUDA_ERROR; }
            ctx->stream = main_stream;
            if (rc != ERL_GP_STATUS_OK) { return rc; }
            if (m > w) {
                // the rest of the trailing matrix: A[j1+w:, j1+w:] -= L[j1+w:, j0:j1] L[j1+w:, j0:j1]^T (lower tiles)
                const T *a2 = a1 + w;
                rc = Gemm<T>(ctx, kOpN, kOpT, m - w, m - w, kOuter, T(-1), a2, ld, a2, ld, T(1), c1 + w + w * ld, ld, 1);
                if (rc != ERL_GP_STATUS_OK) { return rc; }
            }
            if (fork) { ERL_GP_CUDA_OK(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_diag, 0)); }
        }
        return ERL_GP_STATUS_OK;
    }

    // acc[j] += sum_i s[i + j * lds]^2 for i < rows
    template<typename T>
    __global__ void
    ColSumSqKernel(const long rows, const long t, const T *__restrict__ s, const long lds, T *__restrict__ acc) {
        const long j = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const int lane = threadIdx.x & 31;
        if (j >= t) { return; }
        T sum = 0;
        for (long i = lane; i < rows; i += 32) {
            const T v = s[i + j * lds];
            sum += v * v;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, off); }
        if (lane == 0) { acc[j] += sum; }
    }

    template<typename T>
    int
    TrsmLower(Context *ctx, long n, long t, const T *l, long ld, const T *linv, T *w, long ldw, T *s_buf, T *sumsq, bool keep) {
        for (long k0 = 0, kb = 0; k0 < n; k0 += kPanel, ++kb) {
            const long nk = n - k0 < kPanel ? n - k0 : kPanel;
            // S = L_kk^-1 * W_k
            int rc = Gemm<T>(ctx, kOpN, kOpN, nk, t, nk, T(1), linv + kb * kPanel * kPanel, kPanel, w + k0, ldw, T(0), s_buf, kPanel, false);
            if (rc != ERL_GP_STATUS_OK) { return rc; }
            if (sumsq != nullptr) {
                ColSumSqKernel<T><<<static_cast<unsigned>(CeilDiv(t, 8)), 256, 0, ctx->stream>>>(nk, t, s_buf, kPanel, sumsq);
                ctx->launches += 1;
                ERL_GP_CUDA_OK(ctx, cudaGetLastError());
            }
            if (keep) { ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(w + k0, sizeof(T) * ldw, s_buf, sizeof(T) * kPanel, sizeof(T) * nk, t, cudaMemcpyDeviceToDevice, ctx->stream)); }
            const long m = n - k0 - nk;
            if (m <= 0) { break; }
            // W_below -= L(below, k) * S
            rc = Gemm<T>(ctx, kOpN, kOpN, m, t, nk, T(-1), l + (k0 + nk) + k0 * ld, ld, s_buf, kPanel, T(1), w + k0 + nk, ldw, false);
            if (rc != ERL_GP_STATUS_OK) { return rc; }
        }
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    int
    TrsmLowerTrans(Context *ctx, long n, long t, const T *l, long ld, const T *linv, T *z, long ldz, T *s_buf) {
        const long num_panels = CeilDiv(n, kPanel);
        for (long kb = num_panels - 1; kb >= 0; --kb) {
            const long k0 = kb * kPanel;
            const long nk = n - k0 < kPanel ? n - k0 : kPanel;
            // S = L_kk^-T * Z_k
            int rc = Gemm<T>(ctx, kOpT, kOpN, nk, t, nk, T(1), linv + kb * kPanel * kPanel, kPanel, z + k0, ldz, T(0), s_buf, kPanel, false);
            if (rc != ERL_GP_STATUS_OK) { return rc; }
            ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(z + k0, sizeof(T) * ldz, s_buf, sizeof(T) * kPanel, sizeof(T) * nk, t, cudaMemcpyDeviceToDevice, ctx->stream));
            if (k0 <= 0) { break; }
            // Z_above -= L(k, above)^T * S
            rc = Gemm<T>(ctx, kOpT, kOpN, k0, t, nk, T(-1), l + k0, ld, s_buf, kPanel, T(1), z, ldz, false);
            if (rc != ERL_GP_STATUS_OK) { return rc; }
        }
        return ERL_GP_STATUS_OK;
    }

    // =========================================================================================
    // alpha = L^-T L^-1 y for a few right-hand sides (y_dim <= 4): blocked substitution with vector kernels.
    // (Running these solves through 128-wide GEMM tiles cost 35 % of Train() at n = 4096: 4 launches per panel, each
    // computing 127 useless columns.)  Memory-bound: L is read once per direction.
    // =========================================================================================
    constexpr int kTrsvYmax = 4;

    // y_k <- M * y_k with M = Linv_k (forward) or Linv_k^T (backward); one CTA of 1024 threads: thread (i, jq) sums 16
    // columns of row i (all loads independent), the 8 partial sums are combined through shared memory
    template<typename T>
    __global__ void __launch_bounds__(1024)
    TrsvDiagKernel(const T *__restrict__ linv_k, const int nk, T *__restrict__ y, const long ldy, const int y_dim, const int trans) {
        __shared__ T ys[kPanel * kTrsvYmax];
        __shared__ T part[8][kPanel * kTrsvYmax];
        const int i = threadIdx.x & (kPanel - 1);
        const int jq = threadIdx.x >> 7;  // 0..7
        if (threadIdx.x < kPanel) {
            for (int c = 0; c < kTrsvYmax; ++c) { ys[i + c * kPanel] = (i < nk && c < y_dim) ? y[i + c * ldy] : T(0); }
        }
        __syncthreads();
        T sum[kTrsvYmax];
#pragma unroll
        for (int c = 0; c < kTrsvYmax; ++c) { sum[c] = T(0); }
        T mv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int j = 16 * jq + u;
            mv[u] = trans ? linv_k[j + i * kPanel] : linv_k[i + j * kPanel];  // identity-padded beyond nk
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int c = 0; c < kTrsvYmax; ++c) { sum[c] += mv[u] * ys[16 * jq + u + c * kPanel]; }
        }
#pragma unroll
        for (int c = 0; c < kTrsvYmax; ++c) { part[jq][i + c * kPanel] = sum[c]; }
        __syncthreads();
        if (threadIdx.x < kPanel && i < nk) {
            for (int c = 0; c < y_dim; ++c) {
                T tot = T(0);
#pragma unroll
                for (int q = 0; q < 8; ++q) { tot += part[q][i + c * kPanel]; }
                y[i + c * ldy] = tot;
            }
        }
    }

    // forward: y[r] -= sum_c L[r, k0 + c] z[c] for the rows r below the panel.  CTA = 64 rows x 4 column quarters
    // (coalesced along r, 32 independent loads per thread), partial sums combined through shared memory
    template<typename T>
    __global__ void __launch_bounds__(256)
    TrsvUpdateBelowKernel(const T *__restrict__ l, const long ld, const long n, const long k0, const int nk, T *__restrict__ y, const long ldy, const int y_dim) {
        __shared__ T zs[kPanel * kTrsvYmax];
        __shared__ T part[4][64 * kTrsvYmax];
        for (int e = threadIdx.x; e < kPanel * kTrsvYmax; e += blockDim.x) {
            const int j = e % kPanel, c = e / kPanel;
            zs[e] = (j < nk && c < y_dim) ? y[k0 + j + c * ldy] : T(0);
        }
        __syncthreads();
        const int rl = threadIdx.x & 63;
        const int cq = threadIdx.x >> 6;
        const long r = k0 + nk + static_cast<long>(blockIdx.x) * 64 + rl;
        T sum[kTrsvYmax];
#pragma unroll
        for (int c = 0; c < kTrsvYmax; ++c) { sum[c] = T(0); }
        if (r < n) {
            const T *lr = l + r + (k0 + 32 * cq) * ld;
            T lv[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) { lv[u] = 32 * cq + u < nk ? lr[static_cast<long>(u) * ld] : T(0); }
#pragma unroll
            for (int u = 0; u < 32; ++u) {
#pragma unroll
                for (int c = 0; c < kTrsvYmax; ++c) { sum[c] += lv[u] * zs[32 * cq + u + c * kPanel]; }
            }
        }
#pragma unroll
        for (int c = 0; c < kTrsvYmax; ++c) { part[cq][rl + c * 64] = sum[c]; }
        __syncthreads();
        if (cq == 0 && r < n) {
            for (int c = 0; c < y_dim; ++c) { y[r + c * ldy] -= part[0][rl + c * 64] + part[1][rl + c * 64] + part[2][rl + c * 64] + part[3][rl + c * 64]; }
        }
    }

    // backward: y[j] -= sum_c L[k0 + c, j] z[c] for the columns j left of the panel; warp = column (coalesced along c)
    template<typename T>
    __global__ void __launch_bounds__(256)
    TrsvUpdateAboveKernel(const T *__restrict__ l, const long ld, const long k0, const int nk, T *__restrict__ y, const long ldy, const int y_dim) {
        __shared__ T zs[kPanel * kTrsvYmax];
        for (int e = threadIdx.x; e < kPanel * kTrsvYmax; e += blockDim.x) {
            const int j = e % kPanel, c = e / kPanel;
            zs[e] = (j < nk && c < y_dim) ? y[k0 + j + c * ldy] : T(0);
        }
        __syncthreads();
        const long j = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const int lane = threadIdx.x & 31;
        if (j >= k0) { return; }
        T sum[kTrsvYmax];
#pragma unroll
        for (int c = 0; c < kTrsvYmax; ++c) { sum[c] = T(0); }
        const T *lc = l + k0 + j * ld;
        for (int i = lane; i < nk; i += 32) {
            const T lv = lc[i];
#pragma unroll
            for (int c = 0; c < kTrsvYmax; ++c) { sum[c] += lv * zs[i + c * kPanel]; }
        }
#pragma unroll
        for (int c = 0; c < kTrsvYmax; ++c) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) { sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], off); }
        }
        if (lane == 0) {
            for (int c = 0; c < y_dim; ++c) { y[j + c * ldy] -= sum[c]; }
        }
    }

    // ---- wavefront TRSV: the whole substitution in ONE launch per direction -------------------------------------------------
    // One CTA per 128-row block.  Forward (TRANS = false): CTA i accumulates sum_{k < i} L[i, k] z_k as the z_k appear
    // (a flag per block, set by the CTA that produced it), then z_i = Linv_i (y_i - sum) and raises its own flag; backward
    // (TRANS = true) is the mirror image with the column blocks L[k, i]^T, k > i.  The critical path per block is one flag
    // round trip + two 128 x 128 mat-vecs (~2-3 us) instead of two dependent kernel launches (~16 us); the off-critical
    // blocks of L stream in behind the wavefront (register double buffering, one block ahead).
    // Blocks are handed out by a ticket counter in the order the CTAs actually start, and a CTA only ever waits for blocks
    // with smaller tickets, so the spin loops cannot deadlock however many CTAs are resident.
    constexpr int kWaveThreads = 512;

    template<typename T, bool TRANS>
    struct WaveFrag {  // this thread's 32 entries of a 128 x 128 block M (leading dimension ldm): out = M v (or M^T v)
        T m[32];

        // non-TRANS: thread (r = tid & 127, q = tid >> 7) holds M[r, 32 q + j];  TRANS: lane c, warp w hold M[c + 32 i, w + 16 j] (slot 4 j + i)
        __device__ __forceinline__ void
        Load(const T *__restrict__ mat, const long ldm, const long row_lim, const long col_lim, const int tid) {
            if (!TRANS) {
                const int r = tid & 127, q = tid >> 7;
#pragma unroll
                for (int j = 0; j < 32; ++j) { m[j] = (r < row_lim && 32 * q + j < col_lim) ? mat[r + (32 * q + j) * ldm] : T(0); }
            } else {
                const int c = tid & 31, w = tid >> 5;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { m[4 * j + i] = (c + 32 * i < row_lim && w + 16 * j < col_lim) ? mat[(c + 32 * i) + (w + 16 * j) * ldm] : T(0); }
                }
            }
        }

        // adds this block's contribution to acc_s[yy][r] (shared, TRANS) or to the thread's partial sums (non-TRANS)
        __device__ __forceinline__ void
        Apply(const T (*vs)[kPanel], const int y_dim, T (&acc)[kTrsvYmax], T (*acc_s)[kPanel], const int tid) const {
            if (!TRANS) {
                const int q = tid >> 7;
#pragma unroll
                for (int yy = 0; yy < kTrsvYmax; ++yy) {
                    if (yy < y_dim) {
                        T s0 = T(0), s1 = T(0);
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            s0 += m[j] * vs[yy][32 * q + j];
                            s1 += m[j + 1] * vs[yy][32 * q + j + 1];
                        }
                        acc[yy] += s0 + s1;
                    }
                }
            } else {
                const int c = tid & 31, w = tid >> 5;
#pragma unroll
                for (int yy = 0; yy < kTrsvYmax; ++yy) {
                    if (yy < y_dim) {
                        T v4[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) { v4[i] = vs[yy][c + 32 * i]; }
                        T mine = T(0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            T sj = (m[4 * j] * v4[0] + m[4 * j + 1] * v4[1]) + (m[4 * j + 2] * v4[2] + m[4 * j + 3] * v4[3]);
#pragma unroll
                            for (int off = 16; off > 0; off >>= 1) { sj += __shfl_xor_sync(0xffffffffu, sj, off); }
                            if (c == j) { mine = sj; }
                        }
                        if (c < 8) { acc_s[yy][w + 16 * c] += mine; }  // column w + 16 c belongs to this warp only
                    }
                }
            }
        }
    };

    template<typename T, bool TRANS>
    __global__ void __launch_bounds__(kWaveThreads, 1)
    TrsvWavefrontKernel(const T *__restrict__ l, const long ld, const T *__restrict__ linv, const long n, T *y, const long ldy, const int y_dim, int *sync_buf) {
        __shared__ int s_step;
        __shared__ T vs[kTrsvYmax][kPanel];        // the vector the current block is applied to
        __shared__ T acc_s[kTrsvYmax][kPanel];     // TRANS: running sums; non-TRANS: reduction scratch
        __shared__ T red[3][kTrsvYmax][kPanel];    // non-TRANS: partial sums of the column quarters 1..3
        const int tid = threadIdx.x;
        if (tid == 0) { s_step = atomicAdd(sync_buf, 1); }
        for (int e = tid; e < kTrsvYmax * kPanel; e += kWaveThreads) { acc_s[e / kPanel][e % kPanel] = T(0); }
        __syncthreads();
        const int nblk = static_cast<int>((n + kPanel - 1) / kPanel);
        const int step = s_step;
        if (step >= nblk) { return; }
        volatile int *ready = sync_buf + 1;
        const int blk = TRANS ? nblk - 1 - step : step;
        const long r0 = static_cast<long>(blk) * kPanel;
        const long nr = n - r0 < kPanel ? n - r0 : kPanel;  // rows of this block

        T acc[kTrsvYmax];
#pragma unroll
        for (int yy = 0; yy < kTrsvYmax; ++yy) { acc[yy] = T(0); }

        // FP32: the next block is loaded into a second register set before waiting for the current flag; FP64 (64 doubles would
        // not fit the 128-register budget of 512 threads) pulls the next block into L2 with prefetch instructions instead
        constexpr bool kDoubleBuffer = sizeof(T) == 4;
        WaveFrag<T, TRANS> cur;
        WaveFrag<T, (kDoubleBuffer ? TRANS : false)> nxt_storage[kDoubleBuffer ? 1 : 0 + 1];
        auto block_ptr = [&](const int kk) {
            const int kb = TRANS ? nblk - 1 - kk : kk;
            const long k0 = static_cast<long>(kb) * kPanel;
            return TRANS ? l + k0 + r0 * ld : l + r0 + k0 * ld;  // L[kb rows, blk columns]^T : L[blk rows, kb columns]
        };
        auto load_block = [&](WaveFrag<T, TRANS> &f, const int kk) {
            const int kb = TRANS ? nblk - 1 - kk : kk;
            const long k0 = static_cast<long>(kb) * kPanel;
            // non-TRANS: kb < blk is a full block, my rows may be short; TRANS: the last row block may be short
            f.Load(block_ptr(kk), ld, TRANS ? n - k0 : nr, kPanel, tid);
        };
        auto prefetch_block = [&](const int kk) {  // 128 columns x 1 KiB: 2 lines per thread
            const int kb = TRANS ? nblk - 1 - kk : kk;
            const long rows = TRANS ? n - static_cast<long>(kb) * kPanel : nr;
            const T *base = block_ptr(kk);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = tid + kWaveThreads * u;  // 0 .. 1023
                const int col = e >> 3, seg = (e & 7) * 16;
                if (seg < rows) { asm volatile("prefetch.global.L2 [%0];" ::"l"(base + seg + col * ld)); }
            }
        };
        if (step > 0) { load_block(cur, 0); }
        for (int kk = 0; kk < step; ++kk) {
            if (kk + 1 < step) {
                if constexpr (kDoubleBuffer) {
                    load_block(nxt_storage[0], kk + 1);
                } else {
                    prefetch_block(kk + 1);
                }
            }
            const int kb = TRANS ? nblk - 1 - kk : kk;
            if (tid == 0) {
                while (ready[kb] == 0) {}
                __threadfence();
            }
            __syncthreads();  // also: everybody is done with vs of the previous block
            const long k0 = static_cast<long>(kb) * kPanel;
            for (int e = tid; e < kTrsvYmax * kPanel; e += kWaveThreads) {
                const int yy = e / kPanel, c = e % kPanel;
                vs[yy][c] = (yy < y_dim && k0 + c < n) ? __ldcg(y + k0 + c + yy * ldy) : T(0);
            }
            __syncthreads();
            cur.Apply(vs, y_dim, acc, acc_s, tid);
            if (kk + 1 < step) {
                if constexpr (kDoubleBuffer) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) { cur.m[j] = nxt_storage[0].m[j]; }
                } else {
                    load_block(cur, kk + 1);
                }
            }
        }
        // v = y_blk - sum, then z_blk = Linv_blk v (forward) / Linv_blk^T v (backward)
        cur.Load(linv + static_cast<long>(blk) * kPanel * kPanel, kPanel, kPanel, kPanel, tid);
        __syncthreads();
        if (!TRANS) {
            const int r = tid & 127, q = tid >> 7;
#pragma unroll
            for (int yy = 0; yy < kTrsvYmax; ++yy) {
                if (q > 0) {
                    red[q - 1][yy][r] = acc[yy];
                } else {
                    acc_s[yy][r] = acc[yy];
                }
            }
            __syncthreads();
            for (int e = tid; e < kTrsvYmax * kPanel; e += kWaveThreads) {
                const int yy = e / kPanel, c = e % kPanel;
                acc_s[yy][c] += (red[0][yy][c] + red[1][yy][c]) + red[2][yy][c];
            }
            __syncthreads();
        }
        for (int e = tid; e < kTrsvYmax * kPanel; e += kWaveThreads) {
            const int yy = e / kPanel, c = e % kPanel;
            vs[yy][c] = (yy < y_dim && c < nr) ? y[r0 + c + yy * ldy] - acc_s[yy][c] : T(0);
        }
        __syncthreads();
        for (int e = tid; e < kTrsvYmax * kPanel; e += kWaveThreads) { acc_s[e / kPanel][e % kPanel] = T(0); }
#pragma unroll
        for (int yy = 0; yy < kTrsvYmax; ++yy) { acc[yy] = T(0); }
        __syncthreads();
        cur.Apply(vs, y_dim, acc, acc_s, tid);
        __syncthreads();
        if (!TRANS) {
            const int r = tid & 127, q = tid >> 7;
#pragma unroll
            for (int yy = 0; yy < kTrsvYmax; ++yy) {
                if (q > 0) {
                    red[q - 1][yy][r] = acc[yy];
                } else {
                    acc_s[yy][r] = acc[yy];
                }
            }
            __syncthreads();
            for (int e = tid; e < kTrsvYmax * kPanel; e += kWaveThreads) {
                const int yy = e / kPanel, c = e % kPanel;
                acc_s[yy][c] += (red[0][yy][c] + red[1][yy][c]) + red[2][yy][c];
            }
            __syncthreads();
        }
        for (int e = tid; e < kTrsvYmax * kPanel; e += kWaveThreads) {
            const int yy = e / kPanel, c = e % kPanel;
            if (yy < y_dim && c < nr) { y[r0 + c + yy * ldy] = acc_s[yy][c]; }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            ready[blk] = 1;
        }
    }

    template<typename T>
    int
    TrsvSolve(Context *ctx, long n, long y_dim, const T *l, long ld, const T *linv, T *y, long ldy) {
        if (y_dim < 1 || y_dim > kTrsvYmax) { return ERL_GP_STATUS_UNSUPPORTED; }
        static const bool legacy_trsv = std::getenv("ERL_GP_TRSV_LEGACY") != nullptr;  // A/B measurements: two launches per panel
        if (!legacy_trsv) {
            const int nblk = static_cast<int>(CeilDiv(n, kPanel));
            if (ctx->sync_ints_capacity < 2 * (nblk + 1)) {
                if (ctx->sync_ints != nullptr) {
                    ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
                    ERL_GP_CUDA_OK(ctx, cudaFree(ctx->sync_ints));
                    ctx->sync_ints = nullptr;
                    ctx->sync_ints_capacity = 0;
                }
                const int cap = 2 * (nblk + 1) < 1024 ? 1024 : 2 * (nblk + 1);
                ERL_GP_CUDA_OK(ctx, cudaMalloc(&ctx->sync_ints, sizeof(int) * cap));
                ctx->sync_ints_capacity = cap;
            }
            ERL_GP_CUDA_OK(ctx, cudaMemsetAsync(ctx->sync_ints, 0, sizeof(int) * 2 * (nblk + 1), ctx->stream));
            TrsvWavefrontKernel<T, false><<<nblk, kWaveThreads, 0, ctx->stream>>>(l, ld, linv, n, y, ldy, static_cast<int>(y_dim), ctx->sync_ints);
            TrsvWavefrontKernel<T, true><<<nblk, kWaveThreads, 0, ctx->stream>>>(l, ld, linv, n, y, ldy, static_cast<int>(y_dim), ctx->sync_ints + nblk + 1);
            ctx->launches += 2;
            ERL_GP_CUDA_OK(ctx, cudaGetLastError());
            return ERL_GP_STATUS_OK;
        }
        const long num_panels = CeilDiv(n, kPanel);
        const int yd = static_cast<int>(y_dim);
        for (long kb = 0; kb < num_panels; ++kb) {  // z = L^-1 y
            const long k0 = kb * kPanel;
            const int nk = static_cast<int>(n - k0 < kPanel ? n - k0 : kPanel);
            TrsvDiagKernel<T><<<1, 1024, 0, ctx->stream>>>(linv + kb * kPanel * kPanel, nk, y + k0, ldy, yd, 0);
            const long m = n - k0 - nk;
            if (m > 0) { TrsvUpdateBelowKernel<T><<<static_cast<unsigned>(CeilDiv(m, 64)), 256, 0, ctx->stream>>>(l, ld, n, k0, nk, y, ldy, yd); }
            ctx->launches += m > 0 ? 2 : 1;
        }
        for (long kb = num_panels - 1; kb >= 0; --kb) {  // alpha = L^-T z
            const long k0 = kb * kPanel;
            const int nk = static_cast<int>(n - k0 < kPanel ? n - k0 : kPanel);
            TrsvDiagKernel<T><<<1, 1024, 0, ctx->stream>>>(linv + kb * kPanel * kPanel, nk, y + k0, ldy, yd, 1);
            if (k0 > 0) { TrsvUpdateAboveKernel<T><<<static_cast<unsigned>(CeilDiv(k0, 8)), 256, 0, ctx->stream>>>(l, ld, k0, nk, y, ldy, yd); }
            ctx->launches += k0 > 0 ? 2 : 1;
        }
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    __global__ void
    GemvTKernel(const long n, const long t, const T *__restrict__ w, const long ldw, const T *__restrict__ alpha, const long ld_a, const long y_dim, T *__restrict__ out, const long ld_out) {
        const long j = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const int lane = threadIdx.x & 31;
        if (j >= t) { return; }
        for (long c = 0; c < y_dim; ++c) {
            T sum = 0;
            for (long i = lane; i < n; i += 32) { sum += w[i + j * ldw] * alpha[i + c * ld_a]; }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, off); }
            if (lane == 0) { out[j + c * ld_out] = sum; }
        }
    }

    template<typename T>
    int
    GemvT(Context *ctx, long n, long t, const T *w, long ldw, const T *alpha, long ld_a, long y_dim, T *out, long ld_out) {
        GemvTKernel<T><<<static_cast<unsigned>(CeilDiv(t, 8)), 256, 0, ctx->stream>>>(n, t, w, ldw, alpha, ld_a, y_dim, out, ld_out);
        ctx->launches += 1;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    __global__ void
    VarianceFinalizeKernel(const long t, const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ var) {
        const long j = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
        if (j < t) { var[j] = b != nullptr ? T(1) - a[j] + b[j] : T(1) - a[j]; }  // literal prior 1.0f: src/vanilla_gp.cpp:121, src/sparse_pseudo_input_gp.cpp:291
    }

    template<typename T>
    int
    VarianceFinalize(Context *ctx, long t, const T *a, const T *b, T *var) {
        VarianceFinalizeKernel<T><<<static_cast<unsigned>(CeilDiv(t, 256)), 256, 0, ctx->stream>>>(t, a, b, var);
        ctx->launches += 1;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        return ERL_GP_STATUS_OK;
    }

#define ERL_GP_INSTANTIATE_DENSE(T)                                                                                            \
    template int Gemm<T>(Context *, int, int, long, long, long, T, const T *, long, const T *, long, T, T *, long, int);      \
    template int Potrf<T>(Context *, long, T *, long, T *, T *, int *);                                                        \
    template int TrsmLower<T>(Context *, long, long, const T *, long, const T *, T *, long, T *, T *, bool);                   \
    template int TrsmLowerTrans<T>(Context *, long, long, const T *, long, const T *, T *, long, T *);                         \
    template int CopyLower<T>(Context *, long, const T *, long, T *, long);                                                    \
    template int GemvT<T>(Context *, long, long, const T *, long, const T *, long, long, T *, long);                           \
    template int TrsvSolve<T>(Context *, long, long, const T *, long, const T *, T *, long);                                    \
    template int VarianceFinalize<T>(Context *, long, const T *, const T *, T *);
    ERL_GP_INSTANTIATE_DENSE(float)
    ERL_GP_INSTANTIATE_DENSE(double)
#undef ERL_GP_INSTANTIATE_DENSE

}  // namespace erl_gp
