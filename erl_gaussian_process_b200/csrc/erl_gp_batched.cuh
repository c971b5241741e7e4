// One-CTA-per-GP batched train / predict for the thousands of small partition GPs.
//
// Replaces, for every GP of a batch at once,
//   VanillaGaussianProcess::UpdateKtrain + Solve        src/vanilla_gp.cpp:476-505
//   VanillaGaussianProcess::ComputeKtest                src/vanilla_gp.cpp:521-552
//   TestResult::GetMean / GetVariance                   src/vanilla_gp.cpp:61-150
// as they are driven per partition by
//   LidarGaussianProcess2D::Train / TestResult          src/lidar_gp_2d.cpp:366-392, 47-167
//   RangeSensorGaussianProcess3D::Train / TestResult    src/range_sensor_gp_3d.cpp:334-360, 58-178
//   BatchGaussianProcessUpdateTorch::Solve              src/batch_gp_update_torch.cpp:74-82
//
// Design (B200, 148 SMs, 227 KB shared memory / CTA, FP32 71 TFLOP/s, FP64 36 TFLOP/s measured):
//   * the Gram matrix is built straight into shared memory (fused distance + covariance + noise
//     diagonal), as 16x16 blocks of the lower triangle only ("block-packed": n = 256 floats fit);
//   * blocked right-looking Cholesky entirely in shared memory: a warp factors the 16x16
//     diagonal block in registers with shuffles, its inverse is kept (16x16, row-major), the
//     panel solve and the trailing update are register-tiled GEMMs with 128-bit shared loads;
//   * alpha = L^-T L^-1 y by one warp while the other warps stream L to HBM (coalesced);
//   * predict: a 16 x 16 thread grid owns an (n x TQ) tile of V = L^-1 Kt in REGISTERS
//     (rows cyclic over the 16x16 blocks), Ktest entries are generated in registers, the
//     blocked forward substitution is right-looking so the only shared-memory traffic is the
//     16 x TQ block row being solved; mean and ||v||^2 are reduced with warp shuffles.
//   * padding to a multiple of 16 uses the identity (K tail = I => L tail = I, alpha tail = 0),
//     the idea of src/batch_gp_update_torch.cpp:61-69.
// HBM traffic per GP is the algorithmic minimum: x, y, var in; L, alpha out; queries in;
// mean / var out.
#pragma once

#include "erl_gp_internal.cuh"
#include "erl_gp_rowgp.cuh"

namespace erl_gp {
    namespace rowgp64 {
        // FP64 row-GP kernel on DMMA m8n8k4, n <= 128 (erl_gp_rowgp64.cuh, instantiated in erl_gp_rowgp64_x<dim>.cu)
        template<int XDIM>
        int
        Launch(Context *ctx, const BatchParams<double> &params, int mode, int tiles_per_gp);
        extern template int Launch<1>(Context *, const BatchParams<double> &, int, int);
        extern template int Launch<2>(Context *, const BatchParams<double> &, int, int);
        extern template int Launch<3>(Context *, const BatchParams<double> &, int, int);
    }  // namespace rowgp64

    namespace rowgp_tc {
        // tcgen05 / TMEM fused train + predict kernel, n <= 128 (erl_gp_rowgp_tc.cuh, instantiated in erl_gp_rowgp_tc_x<dim>.cu)
        template<int XDIM>
        int
        Launch(Context *ctx, const BatchParams<float> &params);
        extern template int Launch<1>(Context *, const BatchParams<float> &);
        extern template int Launch<2>(Context *, const BatchParams<float> &);
        extern template int Launch<3>(Context *, const BatchParams<float> &);
    }  // namespace rowgp_tc
}  // namespace erl_gp

#include <cstdlib>

namespace erl_gp {

    constexpr int kBatchThreads = 256;

    template<typename T>
    struct DinvLd {  // row stride of the 16x16 inverse blocks (16B-aligned rows, bank-spread)
        static constexpr int value = sizeof(T) == 4 ? 20 : 18;
    };

    __host__ __device__ __forceinline__ int
    LowerBlock(const int bi, const int bj) {
        return (bi * (bi + 1) / 2 + bj) * (kNB * kNB);
    }

    template<typename T>
    __device__ __forceinline__ void
    Load4(const T *p, T (&v)[4]);

    template<>
    __device__ __forceinline__ void
    Load4<float>(const float *p, float (&v)[4]) {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    }

    template<>
    __device__ __forceinline__ void
    Load4<double>(const double *p, double (&v)[4]) {
        const double2 a = *reinterpret_cast<const double2 *>(p);
        const double2 b = *reinterpret_cast<const double2 *>(p + 2);
        v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
    }

    template<typename T>
    __device__ __forceinline__ void
    Store4(T *p, const T (&v)[4]);

    template<>
    __device__ __forceinline__ void
    Store4<float>(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }

    template<>
    __device__ __forceinline__ void
    Store4<double>(double *p, const double (&v)[4]) {
        *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2 *>(p + 2) = make_double2(v[2], v[3]);
    }

    template<typename T>
    __device__ __forceinline__ void
    Load2(const T *p, T (&v)[2]);

    template<>
    __device__ __forceinline__ void
    Load2<float>(const float *p, float (&v)[2]) {
        const float2 t = *reinterpret_cast<const float2 *>(p);
        v[0] = t.x, v[1] = t.y;
    }

    template<>
    __device__ __forceinline__ void
    Load2<double>(const double *p, double (&v)[2]) {
        const double2 t = *reinterpret_cast<const double2 *>(p);
        v[0] = t.x, v[1] = t.y;
    }

    __device__ __forceinline__ float
    Sqrt(float v) {
        return sqrtf(v);
    }

    __device__ __forceinline__ double
    Sqrt(double v) {
        return sqrt(v);
    }

    // Shared-memory carve-up (in elements of T) for a kernel instance that supports up to
    // MROWS diagonal blocks (n <= 16 * MROWS).
    template<typename T, int XDIM, int MROWS>
    struct BatchSmem {
        static constexpr int kQpt = (sizeof(T) == 4 && MROWS <= 8) ? 8 : 4;  // queries per thread
        static constexpr int kTq = 16 * kQpt;                                // queries per tile
        static constexpr int kNpad = kNB * MROWS;
        static constexpr int kLp = 0;
        static constexpr int kDinv = kLp + MROWS * (MROWS + 1) / 2 * kNB * kNB;
        static constexpr int kXs = kDinv + MROWS * kNB * DinvLd<T>::value;
        static constexpr int kAl = kXs + ((kNpad * XDIM + 3) / 4) * 4;
        static constexpr int kR = kAl + kNpad;
        static constexpr int kS = kR + kNB * kTq;
        static constexpr int kQx = kS + kNB * kTq;
        static constexpr int kEnd = kQx + ((kTq * XDIM + 3) / 4) * 4;
        static constexpr size_t kBytes = static_cast<size_t>(kEnd) * sizeof(T) + 16;  // + flags
        static_assert(kNpad <= kNB * kTq, "var scratch aliases the R buffer");
    };

    // 1 / sqrt(v) off the slow paths: FP64 sqrt + division are ~40 dependent instructions each and sat on the critical
    // path of every pivot (16 per block); rsqrt(double) is MUFU.RSQ64H + Newton (<= 1 ulp), the diagonal is v * rsqrt(v).
    __device__ __forceinline__ float
    InvSqrt(float v) {
        return 1.0f / sqrtf(v);
    }

    __device__ __forceinline__ double
    InvSqrt(double v) {
        const double r = rsqrt(v);
        return fma(fma(-v * r, r, 1.0), 0.5 * r, r);  // one more Newton step: correctly rounded in practice
    }

    // ---- 16x16 diagonal block: Cholesky in registers of one warp --------------------------
    // Lane r (< 16) owns row r.  Returns 0 or the 1-based failing column (warp-uniform); invs[k] = 1 / L_kk (every lane).
    template<typename T>
    __device__ __forceinline__ int
    DiagCholesky(T *d, const int lane, T (&invs)[kNB]) {
        constexpr unsigned kFull = 0xffffffffu;
        T a[kNB];
#pragma unroll
        for (int c = 0; c < kNB; ++c) { a[c] = lane < kNB ? d[lane + kNB * c] : T(0); }
        int fail = 0;
#pragma unroll
        for (int k = 0; k < kNB; ++k) {
            const T akk = __shfl_sync(kFull, a[k], k);
            if (!(akk > T(0)) && fail == 0) { fail = k + 1; }
            const T inv = InvSqrt(akk);
            invs[k] = inv;
            const T lk = lane == k ? akk * inv : a[k] * inv;
            a[k] = lk;
#pragma unroll
            for (int j = k + 1; j < kNB; ++j) {
                const T ljk = __shfl_sync(kFull, lk, j);
                a[j] -= lk * ljk;  // lanes < j compute strictly-upper garbage, zeroed below
            }
        }
        if (lane < kNB) {
#pragma unroll
            for (int c = 0; c < kNB; ++c) { d[lane + kNB * c] = c <= lane ? a[c] : T(0); }
        }
        return fail;
    }

    // inverse of a factored lower 16x16 block (col-major d) into row-major dinv (stride DinvLd);
    // lane c computes column c of the inverse by forward substitution on e_c; invs[i] = 1 / d[i, i].
    template<typename T>
    __device__ __forceinline__ void
    DiagInverse(const T *d, T *dinv, const int lane, const T (&invs)[kNB]) {
        constexpr int kLd = DinvLd<T>::value;
        if (lane < kNB) {
            T xc[kNB];
#pragma unroll
            for (int i = 0; i < kNB; ++i) {
                T s = i == lane ? T(1) : T(0);
#pragma unroll
                for (int p = 0; p < i; ++p) { s -= d[i + kNB * p] * xc[p]; }
                xc[i] = s * invs[i];
            }
#pragma unroll
            for (int i = 0; i < kNB; ++i) { dinv[i * kLd + lane] = xc[i]; }
        }
    }

    // same, reciprocals of the diagonal computed here (one division per lane, shared by shuffles); whole warp must call
    template<typename T>
    __device__ __forceinline__ void
    DiagInverse(const T *d, T *dinv, const int lane) {
        const T mine = T(1) / d[(lane & (kNB - 1)) * (kNB + 1)];
        T invs[kNB];
#pragma unroll
        for (int i = 0; i < kNB; ++i) { invs[i] = __shfl_sync(0xffffffffu, mine, i); }
        DiagInverse(d, dinv, lane, invs);
    }

    // ---- 16x16 diagonal block AND its inverse in one elimination (round 2) ------------------
    // Lanes 0 .. 15 own the rows of the block, lanes 16 + j start from the unit vector e_j and run the very same instructions: with
    // sc = a[c] / d the update a[cc] -= sc A[cc][c] is, on e_j, the forward substitution of L x = e_j, so lane 16 + j ends with column j
    // of L^-1 - the inverse costs no instruction of its own (DiagInverse was a second 136-step dependent chain).  The shuffles read
    // the RAW entries A[cc][c] (symmetric tile), so only the reciprocal of the pivot (MUFU seed + Newton) sits on the serial chain; the
    // square roots that scale the eliminated entries into L are taken after the loop, one per lane.  d: col-major 16 x 16, full
    // symmetric tile on entry; dinv: row-major, stride DinvLd.  Returns 0 or the 1-based failing column (warp-uniform).
    __device__ __forceinline__ float
    RcpChain(const float d) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
        return r * fmaf(-d, r, 2.0f);
    }

    __device__ __forceinline__ double
    RcpChain(const double d) {
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
        r = fma(r, fma(-d, r, 1.0), r);
        return fma(r, fma(-d, r, 1.0), r);
    }

    template<typename T>
    __device__ __forceinline__ int
    DiagCholeskyInverse(T *d, T *dinv, const int lane) {
        constexpr unsigned kFull = 0xffffffffu;
        constexpr int kLd = DinvLd<T>::value;
        const int r = lane & (kNB - 1);
        T a[kNB], l[kNB];
#pragma unroll
        for (int c = 0; c < kNB; ++c) {
            const T v = r >= c ? d[r + kNB * c] : d[c + kNB * r];  // the block holds its lower triangle: read the mirror entry above the diagonal
            a[c] = lane < kNB ? v : (c == r ? T(1) : T(0));
        }
        int fail = 0;
        T dmine = T(1);
#pragma unroll
        for (int c = 0; c < kNB; ++c) {
            const T dc = __shfl_sync(kFull, a[c], c);
            T t[kNB];
#pragma unroll
            for (int cc = c + 1; cc < kNB; ++cc) { t[cc] = __shfl_sync(kFull, a[c], cc); }
            if (!(dc > T(0)) && fail == 0) { fail = c + 1; }
            const T sc = a[c] * RcpChain(dc);
#pragma unroll
            for (int cc = c + 1; cc < kNB; ++cc) { a[cc] -= sc * t[cc]; }
            l[c] = a[c];
            if (r == c) { dmine = dc; }
        }
        const T rsv = InvSqrt(dmine);
#pragma unroll
        for (int c = 0; c < kNB; ++c) { l[c] *= __shfl_sync(kFull, rsv, c); }
        __syncwarp();  // every lane has read the raw block
        if (lane < kNB) {
#pragma unroll
            for (int c = 0; c < kNB; ++c) { d[lane + kNB * c] = c <= lane ? l[c] : T(0); }
        } else {
#pragma unroll
            for (int c = 0; c < kNB; ++c) { dinv[c * kLd + r] = l[c]; }  // column r of the inverse: Dinv[c][r]
        }
        return fail;
    }

    // ---- blocked right-looking Cholesky of the block-packed lower triangle in smem ---------
    // On return lp holds L (diagonal blocks with zero strict upper), dinv the inverses of the
    // diagonal blocks.  *s_fail (shared) = 0 or the 1-based failing column.
    template<typename T>
    __device__ void
    CholeskySmem(T *lp, T *dinv, const int nblk, int *s_fail) {
        constexpr int kLd = DinvLd<T>::value;
        const int tid = threadIdx.x;
        const int warp = tid >> 5;
        const int lane = tid & 31;
        for (int kb = 0; kb < nblk; ++kb) {
            T *dkk = lp + LowerBlock(kb, kb);
            T *dinv_k = dinv + kb * kNB * kLd;
            if (warp == 0) {
#ifdef ERL_GP_DIAG_TWO_PASS  // A/B: round-1 version (Cholesky with the scaled entries on the shuffle chain, then a separate inverse)
                T invs[kNB];
                const int fail = DiagCholesky(dkk, lane, invs);
                if (fail != 0 && lane == 0 && *s_fail == 0) { *s_fail = kb * kNB + fail; }
                __syncwarp();
                DiagInverse(dkk, dinv_k, lane, invs);
#else
                const int fail = DiagCholeskyInverse(dkk, dinv_k, lane);
                if (fail != 0 && lane == 0 && *s_fail == 0) { *s_fail = kb * kNB + fail; }
#endif
            }
            __syncthreads();
            const int m = nblk - kb - 1;  // block rows below the diagonal block
            if (m == 0) { break; }
            // panel solve: X(bi, kb) = A(bi, kb) * L_kk^-T, one thread per row
            if (tid < m * kNB) {
                T *blk = lp + LowerBlock(kb + 1 + (tid >> 4), kb) + (tid & 15);
                T arow[kNB];
#pragma unroll
                for (int p = 0; p < kNB; ++p) { arow[p] = blk[kNB * p]; }
#pragma unroll
                for (int c = 0; c < kNB; ++c) {
                    T s = 0;
#pragma unroll
                    for (int p = 0; p <= c; ++p) { s += arow[p] * dinv_k[c * kLd + p]; }
                    blk[kNB * c] = s;
                }
            }
            __syncthreads();
            // trailing update: C(bi, bj) -= X(bi, kb) X(bj, kb)^T for kb < bj <= bi, one warp per block,
            // lane -> 4 rows x 2 columns
            const int num_pairs = m * (m + 1) / 2;
            const int rg = (lane & 3) * 4;
            const int cg = (lane >> 2) * 2;
            for (int q = warp; q < num_pairs; q += kBatchThreads / 32) {
                int ii = 0;
                while ((ii + 1) * (ii + 2) / 2 <= q) { ++ii; }
                const int jj = q - ii * (ii + 1) / 2;
                const int bi = kb + 1 + ii, bj = kb + 1 + jj;
                const T *xa = lp + LowerBlock(bi, kb) + rg;
                const T *xb = lp + LowerBlock(bj, kb) + cg;
                T acc[4][2];
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc[i][0] = acc[i][1] = T(0); }
#pragma unroll
                for (int p = 0; p < kNB; ++p) {
                    T a4[4], b2[2];
                    Load4(xa + kNB * p, a4);
                    Load2(xb + kNB * p, b2);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[i][0] += a4[i] * b2[0];
                        acc[i][1] += a4[i] * b2[1];
                    }
                }
                T *c = lp + LowerBlock(bi, bj) + rg + kNB * cg;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    T c4[4];
                    Load4(c + kNB * j, c4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { c4[i] -= acc[i][j]; }
                    Store4(c + kNB * j, c4);
                }
            }
            __syncthreads();
        }
    }

    // ---- alpha = L^-T L^-1 y by one warp (al holds y on entry, alpha on exit) --------------
    template<typename T>
    __device__ void
    AlphaSolveWarp(const T *lp, const T *dinv, T *al, const int nblk, const int lane) {
        constexpr int kLd = DinvLd<T>::value;
        constexpr unsigned kFull = 0xffffffffu;
        const int npad = nblk * kNB;
        // forward: z = L^-1 y
        for (int kb = 0; kb < nblk; ++kb) {
            const T *dk = dinv + kb * kNB * kLd;
            T z = 0;
            if (lane < kNB) {
#pragma unroll
                for (int p = 0; p < kNB; ++p) { z += dk[lane * kLd + p] * al[kb * kNB + p]; }  // dk strict upper = 0
            }
            __syncwarp();
            if (lane < kNB) { al[kb * kNB + lane] = z; }
            __syncwarp();
            for (int i = (kb + 1) * kNB + lane; i < npad; i += 32) {
                const T *row = lp + LowerBlock(i >> 4, kb) + (i & 15);
                T s = 0;
#pragma unroll
                for (int p = 0; p < kNB; ++p) { s += row[kNB * p] * al[kb * kNB + p]; }
                al[i] -= s;
            }
            __syncwarp();
        }
        // backward: alpha = L^-T z
        for (int kb = nblk - 1; kb >= 0; --kb) {
            T s[kNB];
#pragma unroll
            for (int c = 0; c < kNB; ++c) { s[c] = T(0); }
            for (int i = (kb + 1) * kNB + lane; i < npad; i += 32) {
                const T *row = lp + LowerBlock(i >> 4, kb) + (i & 15);
                const T ai = al[i];
#pragma unroll
                for (int c = 0; c < kNB; ++c) { s[c] += row[kNB * c] * ai; }
            }
#pragma unroll
            for (int c = 0; c < kNB; ++c) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) { s[c] += __shfl_xor_sync(kFull, s[c], off); }
            }
            const T *dk = dinv + kb * kNB * kLd;
            T a = 0;
            if (lane < kNB) {
#pragma unroll
                for (int p = 0; p < kNB; ++p) { a += dk[p * kLd + lane] * (al[kb * kNB + p] - s[p]); }  // Dinv^T
            }
            __syncwarp();
            if (lane < kNB) { al[kb * kNB + lane] = a; }
            __syncwarp();
        }
    }

    // ---- predict one tile of TQ queries; V = L^-1 Kt lives in registers ---------------------
    template<typename T, int XDIM, int MROWS>
    __device__ void
    PredictTile(
        const BatchParams<T> &p,
        const T *lp,
        const T *dinv,
        const T *xs,
        const T *al,
        T *r_buf,
        T *s_buf,
        T *qx,
        const int n,
        const int nblk,
        const long q_begin,
        const int nq) {
        using Smem = BatchSmem<T, XDIM, MROWS>;
        constexpr int kQpt = Smem::kQpt;
        constexpr int kTq = Smem::kTq;
        constexpr int kLd = DinvLd<T>::value;
        constexpr unsigned kFull = 0xffffffffu;
        const int tid = threadIdx.x;
        const int tr = tid & 15;
        const int tc = tid >> 4;
        const int qbase = tc * kQpt;

        // stage the tile's query points (coalesced: the [T][x_dim] list is contiguous)
        for (int e = tid; e < kTq * XDIM; e += kBatchThreads) { qx[e] = e < nq * XDIM ? p.q_x[q_begin * XDIM + e] : T(0); }
        __syncthreads();

        // Ktest entries of this thread: rows tr + 16 m, queries qbase + j
        T v[MROWS][kQpt];
        T macc[kQpt];
#pragma unroll
        for (int j = 0; j < kQpt; ++j) { macc[j] = T(0); }
        T xq[kQpt][XDIM];
#pragma unroll
        for (int j = 0; j < kQpt; ++j) {
#pragma unroll
            for (int d = 0; d < XDIM; ++d) { xq[j][d] = qx[(qbase + j) * XDIM + d]; }
        }
#pragma unroll
        for (int m = 0; m < MROWS; ++m) {
            const int i = tr + kNB * m;
            if (m < nblk && i < n) {
                T xi[XDIM];
#pragma unroll
                for (int d = 0; d < XDIM; ++d) { xi[d] = xs[i * XDIM + d]; }
                const T a = al[i];
#pragma unroll
                for (int j = 0; j < kQpt; ++j) {
                    const T kv = p.cov(SquaredDistance<T, XDIM>(xi, xq[j]));
                    v[m][j] = kv;
                    macc[j] += kv * a;
                }
            } else {
#pragma unroll
                for (int j = 0; j < kQpt; ++j) { v[m][j] = T(0); }
            }
        }

        T sumsq[kQpt];
#pragma unroll
        for (int j = 0; j < kQpt; ++j) { sumsq[j] = T(0); }

        for (int kb = 0; kb < nblk; ++kb) {
            // 1. residual of block row kb -> R
            T rk[kQpt];
#pragma unroll
            for (int m = 0; m < MROWS; ++m) {
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
            // 2. V_kb = Dinv_kb * R   (Dinv strict upper is zero)
            T out[kQpt];
#pragma unroll
            for (int j = 0; j < kQpt; ++j) { out[j] = T(0); }
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
                        for (int jj = 0; jj < 4; ++jj) { out[j + jj] += d4[pp] * r4[jj]; }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kQpt; j += 4) {
                T t4[4] = {out[j], out[j + 1], out[j + 2], out[j + 3]};
                Store4(s_buf + tr * kTq + qbase + j, t4);
            }
#pragma unroll
            for (int j = 0; j < kQpt; ++j) { sumsq[j] += out[j] * out[j]; }
            __syncthreads();
            // 3. right-looking update of the rows below: V(m) -= L(m, kb) * V_kb
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
                    for (int m = 1; m < MROWS; ++m) {
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

        // reduce mean / ||v||^2 over the 16 row-owners of each query (half-warp)
#pragma unroll
        for (int j = 0; j < kQpt; ++j) {
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) {
                macc[j] += __shfl_xor_sync(kFull, macc[j], off);
                sumsq[j] += __shfl_xor_sync(kFull, sumsq[j], off);
            }
        }
        if (tr == 0) {
#pragma unroll
            for (int j = 0; j < kQpt; ++j) {
                const int q = qbase + j;
                if (q < nq) {
                    const long src = q_begin + q;
                    const long dst = p.q_out_index != nullptr ? p.q_out_index[src] : src;
                    if (p.mean != nullptr) {
                        T f = macc[j];
                        if (p.mapping != ERL_GP_MAPPING_NONE) { f = MappingInv<T>(p.mapping, p.mapping_scale, f); }
                        p.mean[dst] = f;
                    }
                    if (p.variance != nullptr) { p.variance[dst] = T(1) - sumsq[j]; }  // literal prior 1.0f, src/vanilla_gp.cpp:121
                    if (p.valid != nullptr) { p.valid[dst] = 1; }
                }
            }
        }
    }

    template<typename T, int XDIM, int MROWS, int MODE>
    __global__ void __launch_bounds__(kBatchThreads, (MROWS <= 8) ? 2 : 1)
    BatchedGpKernel(const BatchParams<T> p) {
        using Smem = BatchSmem<T, XDIM, MROWS>;
        extern __shared__ __align__(16) unsigned char smem_raw[];
        T *smem = reinterpret_cast<T *>(smem_raw);
        T *lp = smem + Smem::kLp;
        T *dinv = smem + Smem::kDinv;
        T *xs = smem + Smem::kXs;
        T *al = smem + Smem::kAl;
        T *r_buf = smem + Smem::kR;
        T *s_buf = smem + Smem::kS;
        T *qx = smem + Smem::kQx;
        int *s_fail = reinterpret_cast<int *>(smem + Smem::kEnd);

        const int g = blockIdx.x;
        const int tid = threadIdx.x;
        const int warp = tid >> 5;
        const int lane = tid & 31;
        const int n = p.n_train[g];
        const long q0 = (MODE & kBatchPredict) ? p.q_offsets[g] : 0;
        const long q1 = (MODE & kBatchPredict) ? p.q_offsets[g + 1] : 0;

        if (MODE & kBatchTrain) {
            if (n <= p.min_train || n <= 0) {  // the reference's `cnt > min_num_samples_per_group` / `cnt > 0` gate
                if (tid == 0) { p.info[g] = -1; }
                if ((MODE & kBatchPredict) && p.valid != nullptr) {
                    for (long q = q0 + tid; q < q1; q += kBatchThreads) { p.valid[p.q_out_index != nullptr ? p.q_out_index[q] : q] = 0; }
                }
                return;
            }
        } else {
            if (p.info[g] != 0 || q1 <= q0) { return; }  // untrained / failed GP: outputs stay untouched
        }
        const int nblk = (n + kNB - 1) / kNB;
        const int npad = nblk * kNB;
        constexpr int kLd = DinvLd<T>::value;

        // ---- stage the training set ----
        const T *gx = p.x + static_cast<long>(g) * p.max_n * XDIM;
        for (int e = tid; e < npad * XDIM; e += kBatchThreads) { xs[e] = e < n * XDIM ? gx[e] : T(0); }
        if (tid == 0) { *s_fail = 0; }

        if (MODE & kBatchTrain) {
            T *svar = r_buf;  // scratch, dead before the predict phase
            const T *gy = p.y + static_cast<long>(g) * p.max_n;
            const T *gv = p.var + static_cast<long>(g) * p.max_n;
            for (int e = tid; e < npad; e += kBatchThreads) {
                al[e] = e < n ? gy[e] : T(0);
                svar[e] = e < n ? gv[e] : T(0);
            }
            __syncthreads();
            // ---- Gram, lower blocks: one element per thread per block ----
            {
                const int r = tid & 15, c = tid >> 4;
                for (int bi = 0; bi < nblk; ++bi) {
                    const int i = bi * kNB + r;
                    T xi[XDIM];
#pragma unroll
                    for (int d = 0; d < XDIM; ++d) { xi[d] = xs[i * XDIM + d]; }
                    for (int bj = 0; bj <= bi; ++bj) {
                        const int j = bj * kNB + c;
                        T val;
                        if (i == j) {
                            val = i < n ? T(1) + svar[i] : T(1);
                        } else if (i > j && i < n) {
                            T xj[XDIM];
#pragma unroll
                            for (int d = 0; d < XDIM; ++d) { xj[d] = xs[j * XDIM + d]; }
                            val = p.cov(SquaredDistance<T, XDIM>(xi, xj));
                        } else {
                            val = T(0);
                        }
                        lp[LowerBlock(bi, bj) + r + kNB * c] = val;
                    }
                }
            }
            __syncthreads();
            CholeskySmem(lp, dinv, nblk, s_fail);
            __syncthreads();
            const int fail = *s_fail;
            if (fail != 0) {
                if (tid == 0) { p.info[g] = fail; }
                if ((MODE & kBatchPredict) && p.valid != nullptr) {
                    for (long q = q0 + tid; q < q1; q += kBatchThreads) { p.valid[p.q_out_index != nullptr ? p.q_out_index[q] : q] = 0; }
                }
                return;
            }
            // ---- alpha (warp 0) overlapped with the L write-back (other warps) ----
            if (warp == 0) {
                AlphaSolveWarp(lp, dinv, al, nblk, lane);
                if (lane == 0) { p.info[g] = 0; }
            } else if (p.write_l) {
                T *gl = p.l + static_cast<long>(g) * p.max_n * p.max_n;
                for (int c = warp - 1; c < n; c += kBatchThreads / 32 - 1) {
                    for (int r = lane; r < n; r += 32) { gl[r + static_cast<long>(c) * p.max_n] = r >= c ? lp[LowerBlock(r >> 4, c >> 4) + (r & 15) + kNB * (c & 15)] : T(0); }
                }
            }
            __syncthreads();
            T *ga = p.alpha + static_cast<long>(g) * p.max_n;
            for (int e = tid; e < n; e += kBatchThreads) { ga[e] = al[e]; }
        } else {
            // ---- predict-only: reload L / alpha, rebuild the diagonal-block inverses ----
            const T *gl = p.l + static_cast<long>(g) * p.max_n * p.max_n;
            const T *ga = p.alpha + static_cast<long>(g) * p.max_n;
            for (int e = tid; e < npad; e += kBatchThreads) { al[e] = e < n ? ga[e] : T(0); }
            for (int c = warp; c < npad; c += kBatchThreads / 32) {
                for (int r = (c & ~15) + lane; r < npad; r += 32) {
                    T val;
                    if (r < n && c < n) {
                        val = r >= c ? gl[r + static_cast<long>(c) * p.max_n] : T(0);
                    } else {
                        val = r == c ? T(1) : T(0);
                    }
                    lp[LowerBlock(r >> 4, c >> 4) + (r & 15) + kNB * (c & 15)] = val;
                }
            }
            __syncthreads();
            for (int kb = warp; kb < nblk; kb += kBatchThreads / 32) { DiagInverse(lp + LowerBlock(kb, kb), dinv + kb * kNB * kLd, lane); }
        }

        if (MODE & kBatchPredict) {
            __syncthreads();
            constexpr int kTq = Smem::kTq;
            for (long qb = q0 + static_cast<long>(blockIdx.y) * kTq; qb < q1; qb += static_cast<long>(gridDim.y) * kTq) {
                const int nq = static_cast<int>(q1 - qb < kTq ? q1 - qb : kTq);
                PredictTile<T, XDIM, MROWS>(p, lp, dinv, xs, al, r_buf, s_buf, qx, n, nblk, qb, nq);
                __syncthreads();
            }
        }
    }

    template<typename T, int XDIM, int MROWS, int MODE>
    static int
    LaunchBatchInstance(Context *ctx, const BatchParams<T> &params, const int tiles_per_gp) {
        using Smem = BatchSmem<T, XDIM, MROWS>;
        auto kernel = BatchedGpKernel<T, XDIM, MROWS, MODE>;
        if (static_cast<int>(Smem::kBytes) > ctx->max_smem_optin) {
            return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "batched kernel needs %zu B of shared memory, device allows %d", Smem::kBytes, ctx->max_smem_optin);
        }
        ERL_GP_CUDA_OK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(Smem::kBytes)));
        const dim3 grid(static_cast<unsigned>(params.num_gps), static_cast<unsigned>(tiles_per_gp < 1 ? 1 : tiles_per_gp));
        kernel<<<grid, kBatchThreads, Smem::kBytes, ctx->stream>>>(params);
        ctx->launches += 1;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        return ERL_GP_STATUS_OK;
    }

    template<typename T, int XDIM, int MROWS>
    static int
    LaunchBatchMode(Context *ctx, const BatchParams<T> &params, const int mode, const int tiles_per_gp) {
        switch (mode) {
            case kBatchTrain: return LaunchBatchInstance<T, XDIM, MROWS, kBatchTrain>(ctx, params, 1);
            case kBatchPredict: return LaunchBatchInstance<T, XDIM, MROWS, kBatchPredict>(ctx, params, tiles_per_gp);
            case kBatchTrainPredict: return LaunchBatchInstance<T, XDIM, MROWS, kBatchTrainPredict>(ctx, params, 1);
            default: return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "batch: bad mode %d", mode);
        }
    }

    template<typename T, int XDIM>
    int
    LaunchBatchXdim(Context *ctx, const BatchParams<T> &params, const int mode, const int tiles_per_gp) {
        const int max_n = params.max_n;
        if constexpr (sizeof(T) == 4) {
            // FP32, n <= 128: the register-resident "row GP" kernel (erl_gp_rowgp.cuh); ERL_GP_BATCH_LEGACY=1 keeps the
            // generic shared-memory kernel below (A/B measurements and tests of the generic path)
            static const bool legacy = std::getenv("ERL_GP_BATCH_LEGACY") != nullptr;
            static const bool legacy_large = std::getenv("ERL_GP_BATCH_LEGACY_LARGE") != nullptr;  // generic kernel for 128 < n <= 256 only
            // fused train + predict, n <= 128: the tcgen05 / TMEM kernel (erl_gp_rowgp_tc.cuh) when the context asks for it
            // (erl_gp_context_set_rowgp_tc) or ERL_GP_ROWGP_TC=1; the mma.sync kernel below is the default: it is the faster of
            // the two on the C4 stream (4.90 vs 5.30 ms)
            static const bool tc_env = std::getenv("ERL_GP_ROWGP_TC") != nullptr && std::atoi(std::getenv("ERL_GP_ROWGP_TC")) != 0;
            const bool tc_on = ctx->rowgp_tc >= 0 ? ctx->rowgp_tc != 0 : tc_env;
            if (max_n <= 128 && !legacy && tc_on && mode == kBatchTrainPredict && params.q_out_index == nullptr && params.mapping == ERL_GP_MAPPING_NONE) {
                return rowgp_tc::Launch<XDIM>(ctx, params);
            }
            if (max_n <= 128 && !legacy) { return rowgp::Launch<XDIM>(ctx, params, mode, tiles_per_gp); }
            if (max_n <= 256 && !legacy && !legacy_large) { return rowgp::Launch<XDIM>(ctx, params, mode, tiles_per_gp); }
        }
        if constexpr (sizeof(T) == 8) {
            // FP64, n <= 192: the DMMA row-GP kernel (erl_gp_rowgp64.cuh); ERL_GP_BATCH_LEGACY=1 keeps the generic kernel below
            static const bool legacy64 = std::getenv("ERL_GP_BATCH_LEGACY") != nullptr;
            static const bool legacy64_large = std::getenv("ERL_GP_BATCH_LEGACY_LARGE") != nullptr;  // only 128 < n <= 192 on the generic kernel
            if (max_n <= 128 && !legacy64) { return rowgp64::Launch<XDIM>(ctx, params, mode, tiles_per_gp); }
            if (max_n <= 192 && !legacy64 && !legacy64_large) { return rowgp64::Launch<XDIM>(ctx, params, mode, tiles_per_gp); }
        }
        if (max_n <= 64) { return LaunchBatchMode<T, XDIM, 4>(ctx, params, mode, tiles_per_gp); }
        if (max_n <= 128) { return LaunchBatchMode<T, XDIM, 8>(ctx, params, mode, tiles_per_gp); }
        if (max_n <= 192) { return LaunchBatchMode<T, XDIM, 12>(ctx, params, mode, tiles_per_gp); }
        if constexpr (sizeof(T) == 4) {
            if (max_n <= 256) { return LaunchBatchMode<T, XDIM, 16>(ctx, params, mode, tiles_per_gp); }
        }
        return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "batch: max_n=%d exceeds the one-CTA-per-GP limit (%ld)", max_n, BatchMaxN<T>());
    }

}  // namespace erl_gp
