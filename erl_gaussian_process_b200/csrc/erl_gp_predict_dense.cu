// Fused predict kernels of the dense (HBM-resident) VanillaGaussianProcess.
//
//   TestResult::GetMean      src/vanilla_gp.cpp:61-104    mean[j, c] = sum_i k(x_i, x*_j) alpha[i, c]
//   TestResult::GetVariance  src/vanilla_gp.cpp:106-150   var[j]     = 1 - || L^-1 k(X, x*_j) ||^2
//
// The reference materialises Ktest (n x T) and V = L^-1 Ktest (n x T).  Here Ktest entries are generated
// where they are consumed and the triangular solve is LEFT-looking with one CTA per 128 test points:
//
//   for every 128-row panel p of L (top to bottom), CTA c (test points 128 c .. 128 c + 127):
//       W    = Ktest[panel p rows, my points]            generated in registers (fused distance + covariance)
//       W   -= L[panel p rows, 0 : 128 p] * V[0 : 128 p, my points]      the GEMM that carries all the flops, K = 128 p
//       V_p  = Linv_p * W                                 128 x 128 x 128 GEMM against the kept inverse of L's diagonal block
//       sumsq += column sums of V_p^2;  V_p -> HBM slab of this CTA (operand of the later panels)
//
// Compared with the right-looking sequence of (diagonal GEMM, column-sum kernel, trailing GEMM) per panel this
// has no dependency between CTAs (one launch per sweep instead of 3 per panel), reads / writes every V entry
// once instead of re-reading and re-writing the whole trailing block per panel, and keeps all 148 SMs busy
// during the 128-row diagonal solves.  All CTAs walk L in step, so a panel row of L is fetched from HBM once
// and served to the other CTAs from L2.  Bound: FP64 / FP32 FMA pipe; useful flops = T (n^2 + n).
#include "erl_gp_dense.cuh"
#include "erl_gp_dense_mma.cuh"

#include <cstdlib>

namespace erl_gp {

    namespace {

        constexpr int kTile = 128;    // C tile edge: panel rows x test points
        constexpr int kBk = 16;       // reduction slab
        constexpr int kPad = 4;
        constexpr int kLd = kTile + kPad;
        static_assert(kLd == kMmaLd && kBk == kMmaBk, "slab layout shared with erl_gp_dense_mma.cuh");
        constexpr int kThreads = 256;

        template<typename T>
        __device__ __forceinline__ void
        LdVec4(const T *p, T (&v)[4]);
        template<>
        __device__ __forceinline__ void
        LdVec4<float>(const float *p, float (&v)[4]) {
            const float4 t = *reinterpret_cast<const float4 *>(p);
            v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
        }
        template<>
        __device__ __forceinline__ void
        LdVec4<double>(const double *p, double (&v)[4]) {
            const double2 a = *reinterpret_cast<const double2 *>(p);
            const double2 b = *reinterpret_cast<const double2 *>(p + 2);
            v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
        }
        template<typename T>
        __device__ __forceinline__ void
        StVec4(T *p, const T (&v)[4]);
        template<>
        __device__ __forceinline__ void
        StVec4<float>(float *p, const float (&v)[4]) {
            *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
        }
        template<>
        __device__ __forceinline__ void
        StVec4<double>(double *p, const double (&v)[4]) {
            *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
            *reinterpret_cast<double2 *>(p + 2) = make_double2(v[2], v[3]);
        }

        // A slab (128 rows x 16 k) of a column-major matrix (rows contiguous) -> registers.  Thread: 4 consecutive rows, 2 k's.
        template<typename T>
        __device__ __forceinline__ void
        LoadA(const T *__restrict__ a, const long lda, const long row0, const long row_lim, const long k0, const long k_lim, const bool vec, const int tid, T (&reg)[8]) {
            const int r = (tid & 31) * 4;
            const int kk = tid >> 5;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const long k = k0 + kk + 8 * i;
                const long row = row0 + r;
                T v4[4] = {T(0), T(0), T(0), T(0)};
                if (k < k_lim) {
                    if (vec && row + 3 < row_lim) {
                        LdVec4<T>(a + row + k * lda, v4);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (row + j < row_lim) { v4[j] = a[row + j + k * lda]; }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) { reg[4 * i + j] = v4[j]; }
            }
        }

        template<typename T>
        __device__ __forceinline__ void
        StoreA(T *__restrict__ dst /* [16][kLd] */, const int tid, const T (&reg)[8]) {
            const int r = (tid & 31) * 4;
            const int kk = tid >> 5;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const T v4[4] = {reg[4 * i], reg[4 * i + 1], reg[4 * i + 2], reg[4 * i + 3]};
                StVec4<T>(dst + (kk + 8 * i) * kLd + r, v4);
            }
        }

        // B slab (16 k x 128 columns) of V (k contiguous: element (k, j) at v[k + j * ldv]) -> registers.
        // Thread: 4 consecutive k's of 2 columns.
        template<typename T>
        __device__ __forceinline__ void
        LoadB(const T *__restrict__ v, const long ldv, const long k0, const int tid, T (&reg)[8]) {
            const int k4 = (tid & 3) * 4;
            const int c = tid >> 2;  // 0..63
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                T v4[4];
                LdVec4<T>(v + k0 + k4 + static_cast<long>(c + 64 * i) * ldv, v4);  // k0 is a multiple of 16, slabs are fully inside the solved rows
#pragma unroll
                for (int j = 0; j < 4; ++j) { reg[4 * i + j] = v4[j]; }
            }
        }

        template<typename T>
        __device__ __forceinline__ void
        StoreB(T *__restrict__ dst /* [16][kLd] */, const int tid, const T (&reg)[8]) {
            const int k4 = (tid & 3) * 4;
            const int c = tid >> 2;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { dst[(k4 + j) * kLd + c + 64 * i] = reg[4 * i + j]; }
            }
        }

        // acc[i][j] += sum_kk at[kk][rows(i)] * bt[kk][cols(j)]  over one 16-deep slab
        template<typename T>
        __device__ __forceinline__ void
        SlabFma(T (&acc)[8][8], const T *__restrict__ at, const T *__restrict__ bt, const int tx, const int ty) {
#pragma unroll
            for (int kk = 0; kk < kBk; ++kk) {
                T av[8], bv[8], t4[4];
                LdVec4<T>(at + kk * kLd + tx * 4, t4);
                av[0] = t4[0], av[1] = t4[1], av[2] = t4[2], av[3] = t4[3];
                LdVec4<T>(at + kk * kLd + 64 + tx * 4, t4);
                av[4] = t4[0], av[5] = t4[1], av[6] = t4[2], av[7] = t4[3];
                LdVec4<T>(bt + kk * kLd + ty * 4, t4);
                bv[0] = t4[0], bv[1] = t4[1], bv[2] = t4[2], bv[3] = t4[3];
                LdVec4<T>(bt + kk * kLd + 64 + ty * 4, t4);
                bv[4] = t4[0], bv[5] = t4[1], bv[6] = t4[2], bv[7] = t4[3];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { acc[i][j] += av[i] * bv[j]; }
                }
            }
        }

        template<typename T, int XDIM>
        __global__ void __launch_bounds__(kThreads, 1)
        PredictVarianceKernel(
            const Covariance<T> cov,
            const long n,
            const long t,
            const T *__restrict__ x_train,  // [n][XDIM]
            const T *__restrict__ x_test,   // [t][XDIM]
            const T *__restrict__ l,
            const long ldl,
            const T *__restrict__ linv,  // per panel 128 x 128, identity padded
            T *__restrict__ v_slabs,     // gridDim.x slabs of n_pad x 128 (ld = n_pad)
            const long n_pad,
            T *__restrict__ sumsq) {  // [t]
            extern __shared__ __align__(16) unsigned char smem_raw[];
            T *as = reinterpret_cast<T *>(smem_raw);  // [2][16][kLd]
            T *bs = as + 2 * kBk * kLd;               // [2][16][kLd]
            T *wt = bs + 2 * kBk * kLd;               // [128][kLd]  W tile, row (k) major: operand B of the diagonal solve
            T *xqs = wt + kTile * kLd;                // [128][XDIM] test points of the current column tile
            const int tid = threadIdx.x;
            const int tx = tid & 15, ty = tid >> 4;
            T *vs = v_slabs + static_cast<long>(blockIdx.x) * n_pad * kTile;
            const long num_panels = (n + kTile - 1) / kTile;
            const bool vec_l = (ldl & 3) == 0 && (reinterpret_cast<uintptr_t>(l) & 31) == 0;
            const long num_ct = (t + kTile - 1) / kTile;

            for (long ct = blockIdx.x; ct < num_ct; ct += gridDim.x) {
                const long col0 = ct * kTile;
                // the tile's test points -> shared (they are re-read once per panel; registers are needed for the 8 x 8 tile)
                __syncthreads();
                for (int e = tid; e < kTile * XDIM; e += kThreads) { xqs[e] = col0 * XDIM + e < t * XDIM ? x_test[col0 * XDIM + e] : T(0); }
                __syncthreads();
                T ssq[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { ssq[j] = T(0); }

                for (long p = 0; p < num_panels; ++p) {
                    const long k0 = p * kTile;
                    T acc[8][8];
                    // ---- acc = -Ktest[panel rows, my points] (the sign makes the update below a plain += ) ----
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const long row = k0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
                        T xi[XDIM];
#pragma unroll
                        for (int d = 0; d < XDIM; ++d) { xi[d] = row < n ? x_train[row * XDIM + d] : T(0); }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int c = j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4);
                            acc[i][j] = (row < n && col0 + c < t) ? -cov(SquaredDistance<T, XDIM>(xi, xqs + c * XDIM)) : T(0);
                        }
                    }
                    // ---- acc += L[panel rows, 0:k0] * V[0:k0, my points] ----
                    if (k0 > 0) {
                        T ra[8], rb[8];
                        LoadA<T>(l, ldl, k0, n, 0, k0, vec_l, tid, ra);
                        LoadB<T>(vs, n_pad, 0, tid, rb);
                        StoreA<T>(as, tid, ra);
                        StoreB<T>(bs, tid, rb);
                        __syncthreads();
                        const long num_kt = k0 / kBk;
                        for (long kt = 0; kt < num_kt; ++kt) {
                            const int cur = static_cast<int>(kt & 1);
                            if (kt + 1 < num_kt) {
                                LoadA<T>(l, ldl, k0, n, (kt + 1) * kBk, k0, vec_l, tid, ra);
                                LoadB<T>(vs, n_pad, (kt + 1) * kBk, tid, rb);
                            }
                            SlabFma<T>(acc, as + cur * kBk * kLd, bs + cur * kBk * kLd, tx, ty);
                            if (kt + 1 < num_kt) {
                                StoreA<T>(as + (cur ^ 1) * kBk * kLd, tid, ra);
                                StoreB<T>(bs + (cur ^ 1) * kBk * kLd, tid, rb);
                            }
                            __syncthreads();
                        }
                    }
                    // ---- W tile -> shared (row major over the panel rows), negated back ----
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4);
                        const T lo[4] = {-acc[i][0], -acc[i][1], -acc[i][2], -acc[i][3]};
                        const T hi[4] = {-acc[i][4], -acc[i][5], -acc[i][6], -acc[i][7]};
                        StVec4<T>(wt + row * kLd + ty * 4, lo);
                        StVec4<T>(wt + row * kLd + 64 + ty * 4, hi);
                    }
                    // ---- V_p = Linv_p * W  (A = Linv_p from HBM / L2 through the slab buffers, B = W tile in shared) ----
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) { acc[i][j] = T(0); }
                    }
                    {
                        const T *lip = linv + p * kTile * kTile;
                        T ra[8];
                        LoadA<T>(lip, kTile, 0, kTile, 0, kTile, true, tid, ra);
                        StoreA<T>(as, tid, ra);
                        __syncthreads();  // also publishes the W tile
                        for (int kt = 0; kt < kTile / kBk; ++kt) {
                            const int cur = kt & 1;
                            if (kt + 1 < kTile / kBk) { LoadA<T>(lip, kTile, 0, kTile, (kt + 1) * kBk, kTile, true, tid, ra); }
                            SlabFma<T>(acc, as + cur * kBk * kLd, wt + kt * kBk * kLd, tx, ty);
                            if (kt + 1 < kTile / kBk) { StoreA<T>(as + (cur ^ 1) * kBk * kLd, tid, ra); }
                            __syncthreads();
                        }
                    }
                    // ---- column sums of squares; V_p -> my slab (operand of the later panels) ----
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4);
                        T s = T(0);
#pragma unroll
                        for (int i = 0; i < 8; ++i) { s += acc[i][j] * acc[i][j]; }
                        ssq[j] += s;
                        if (p + 1 < num_panels) {
                            const T lo[4] = {acc[0][j], acc[1][j], acc[2][j], acc[3][j]};
                            const T hi[4] = {acc[4][j], acc[5][j], acc[6][j], acc[7][j]};
                            StVec4<T>(vs + k0 + tx * 4 + static_cast<long>(c) * n_pad, lo);
                            StVec4<T>(vs + k0 + 64 + tx * 4 + static_cast<long>(c) * n_pad, hi);
                        }
                    }
                    __syncthreads();  // the slab rows written above are read by this CTA in the next panel
                }
                // ---- reduce the column sums over the 16 row-owners (tx) and write ----
#pragma unroll
                for (int j = 0; j < 8; ++j) {
#pragma unroll
                    for (int off = 8; off > 0; off >>= 1) { ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], off); }
                }
                if (tx == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const long col = col0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
                        if (col < t) { sumsq[col] = ssq[j]; }
                    }
                }
            }
        }

        // ---------------------------------------------------------------------------------------------
        // FP64 variant on the tensor path: the same left-looking sweep, but the 128 x 128 tile is accumulated with
        // mma.sync.m8n8k4.f64 (SASS DMMA).  A DFMA needs two fresh 64-bit register operands per MAC and the FMA loop
        // above saturates the register file at ~52 % of the FP64 peak (19.3 of 37 TFLOP/s measured, the same
        // ratio as FFMA with fresh operands in tools/fma_lds_rate.cu); a DMMA does 8 MACs per lane from four operand
        // registers.  Warp w owns rows 64 (w & 1) .. +63 and columns 32 (w >> 1) .. +31 of the tile:
        //   acc[mi][ni][e] = C[64 wm + 8 mi + lane / 4][32 wn + 8 ni + 2 (lane % 4) + e]
        // ---------------------------------------------------------------------------------------------
        template<int XDIM>
        __global__ void __launch_bounds__(kThreads, 1)
        PredictVarianceKernelDmma(
            const Covariance<double> cov,
            const long n,
            const long t,
            const double *__restrict__ x_train,
            const double *__restrict__ x_test,
            const double *__restrict__ l,
            const long ldl,
            const double *__restrict__ linv,
            double *__restrict__ v_slabs,
            const long n_pad,
            double *__restrict__ sumsq) {
            using T = double;
            extern __shared__ __align__(16) unsigned char smem_raw[];
            T *as = reinterpret_cast<T *>(smem_raw);  // [2][16][kLd]
            T *bs = as + 2 * kBk * kLd;               // [2][16][kLd]
            T *wt = bs + 2 * kBk * kLd;               // [128][kLd]  W tile, row (k) major
            T *xqs = wt + kTile * kLd;                // [128][XDIM]
            T *ssq_s = xqs + kTile * 3;               // [128] column sums of the tile
            const int tid = threadIdx.x;
            const int lane = tid & 31;
            const int warp = tid >> 5;
            const int wm = warp & 1, wn = warp >> 1;
            const int g = lane >> 2, kq = lane & 3;
            T *vs = v_slabs + static_cast<long>(blockIdx.x) * n_pad * kTile;
            const long num_panels = (n + kTile - 1) / kTile;
            const bool vec_l = (ldl & 3) == 0 && (reinterpret_cast<uintptr_t>(l) & 31) == 0;
            const long num_ct = (t + kTile - 1) / kTile;

            for (long ct = blockIdx.x; ct < num_ct; ct += gridDim.x) {
                const long col0 = ct * kTile;
                __syncthreads();
                for (int e = tid; e < kTile * XDIM; e += kThreads) { xqs[e] = col0 * XDIM + e < t * XDIM ? x_test[col0 * XDIM + e] : T(0); }
                if (tid < kTile) { ssq_s[tid] = T(0); }
                __syncthreads();
                T ssq[4][2];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) { ssq[ni][0] = ssq[ni][1] = T(0); }

                for (long p = 0; p < num_panels; ++p) {
                    const long k0 = p * kTile;
                    T acc[8][4][2];
                    // ---- acc = -Ktest[panel rows, my points] ----
#pragma unroll
                    for (int mi = 0; mi < 8; ++mi) {
                        const long row = k0 + 64 * wm + 8 * mi + g;
                        T xi[XDIM];
#pragma unroll
                        for (int d = 0; d < XDIM; ++d) { xi[d] = row < n ? x_train[row * XDIM + d] : T(0); }
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int c = 32 * wn + 8 * ni + 2 * kq + e;
                                acc[mi][ni][e] = (row < n && col0 + c < t) ? -cov(SquaredDistance<T, XDIM>(xi, xqs + c * XDIM)) : T(0);
                            }
                        }
                    }
                    // ---- acc += L[panel rows, 0:k0] * V[0:k0, my points] ----
                    if (k0 > 0) {
                        T ra[8], rb[8];
                        LoadA<T>(l, ldl, k0, n, 0, k0, vec_l, tid, ra);
                        LoadB<T>(vs, n_pad, 0, tid, rb);
                        StoreA<T>(as, tid, ra);
                        StoreB<T>(bs, tid, rb);
                        __syncthreads();
                        const long num_kt = k0 / kBk;
                        for (long kt = 0; kt < num_kt; ++kt) {
                            const int cur = static_cast<int>(kt & 1);
                            if (kt + 1 < num_kt) {
                                LoadA<T>(l, ldl, k0, n, (kt + 1) * kBk, k0, vec_l, tid, ra);
                                LoadB<T>(vs, n_pad, (kt + 1) * kBk, tid, rb);
                            }
                            SlabMma(acc, as + cur * kBk * kLd, bs + cur * kBk * kLd, wm, wn, lane);
                            if (kt + 1 < num_kt) {
                                StoreA<T>(as + (cur ^ 1) * kBk * kLd, tid, ra);
                                StoreB<T>(bs + (cur ^ 1) * kBk * kLd, tid, rb);
                            }
                            __syncthreads();
                        }
                    }
                    // ---- W tile -> shared, negated back ----
#pragma unroll
                    for (int mi = 0; mi < 8; ++mi) {
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) {
                            *reinterpret_cast<double2 *>(wt + (64 * wm + 8 * mi + g) * kLd + 32 * wn + 8 * ni + 2 * kq) = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
                            acc[mi][ni][0] = acc[mi][ni][1] = T(0);
                        }
                    }
                    // ---- V_p = Linv_p * W ----
                    {
                        const T *lip = linv + p * kTile * kTile;
                        T ra[8];
                        LoadA<T>(lip, kTile, 0, kTile, 0, kTile, true, tid, ra);
                        StoreA<T>(as, tid, ra);
                        __syncthreads();  // also publishes the W tile
                        for (int kt = 0; kt < kTile / kBk; ++kt) {
                            const int cur = kt & 1;
                            if (kt + 1 < kTile / kBk) { LoadA<T>(lip, kTile, 0, kTile, (kt + 1) * kBk, kTile, true, tid, ra); }
                            SlabMma(acc, as + cur * kBk * kLd, wt + kt * kBk * kLd, wm, wn, lane);
                            if (kt + 1 < kTile / kBk) { StoreA<T>(as + (cur ^ 1) * kBk * kLd, tid, ra); }
                            __syncthreads();
                        }
                    }
                    // ---- column sums of squares; V_p -> my slab ----
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int c = 32 * wn + 8 * ni + 2 * kq + e;
                            T s = T(0);
#pragma unroll
                            for (int mi = 0; mi < 8; ++mi) { s += acc[mi][ni][e] * acc[mi][ni][e]; }
                            ssq[ni][e] += s;
                            if (p + 1 < num_panels) {
#pragma unroll
                                for (int mi = 0; mi < 8; ++mi) { vs[k0 + 64 * wm + 8 * mi + g + static_cast<long>(c) * n_pad] = acc[mi][ni][e]; }
                            }
                        }
                    }
                    __syncthreads();  // the slab rows written above are read by this CTA in the next panel
                }
                // ---- reduce the column sums over the 8 row groups of the warp (lane / 4) and the two row-halves (wm) ----
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        T s = ssq[ni][e];
                        s += __shfl_xor_sync(0xffffffffu, s, 4);
                        s += __shfl_xor_sync(0xffffffffu, s, 8);
                        s += __shfl_xor_sync(0xffffffffu, s, 16);
                        if (g == 0) { atomicAdd(ssq_s + 32 * wn + 8 * ni + 2 * kq + e, s); }
                    }
                }
                __syncthreads();
                if (tid < kTile && col0 + tid < t) { sumsq[col0 + tid] = ssq_s[tid]; }
            }
        }

        // ---- 16-warp variant of the DMMA kernel: 32 x 32 warp tiles, 4 warps per scheduler instead of 2 (the 8-warp kernel keeps the
        // tensor pipe 70.6 % active; the same change took the GEMM of the factorisation from 73.5 % upwards) -----------------------
        constexpr int kThreads16 = 512;

        template<typename T>
        __device__ __forceinline__ void
        LoadA16(const T *__restrict__ a, const long lda, const long row0, const long row_lim, const long k0, const long k_lim, const bool vec, const int tid, T (&reg)[4]) {
            const long row = row0 + (tid & 31) * 4;
            const long k = k0 + (tid >> 5);  // 0 .. 15
            T v4[4] = {T(0), T(0), T(0), T(0)};
            if (k < k_lim) {
                if (vec && row + 3 < row_lim) {
                    LdVec4<T>(a + row + k * lda, v4);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (row + j < row_lim) { v4[j] = a[row + j + k * lda]; }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { reg[j] = v4[j]; }
        }

        template<typename T>
        __device__ __forceinline__ void
        StoreA16(T *__restrict__ dst /* [16][kLd] */, const int tid, const T (&reg)[4]) {
            StVec4<T>(dst + (tid >> 5) * kLd + (tid & 31) * 4, reg);
        }

        template<typename T>
        __device__ __forceinline__ void
        LoadB16(const T *__restrict__ v, const long ldv, const long k0, const int tid, T (&reg)[4]) {
            LdVec4<T>(v + k0 + (tid & 3) * 4 + static_cast<long>(tid >> 2) * ldv, reg);  // column tid / 4 (0 .. 127), 4 consecutive k
        }

        template<typename T>
        __device__ __forceinline__ void
        StoreB16(T *__restrict__ dst /* [16][kLd] */, const int tid, const T (&reg)[4]) {
            const int k4 = (tid & 3) * 4;
            const int c = tid >> 2;
#pragma unroll
            for (int j = 0; j < 4; ++j) { dst[(k4 + j) * kLd + c] = reg[j]; }
        }

        __device__ __forceinline__ void
        SlabMma16(double (&acc)[4][4][2], const double *__restrict__ at, const double *__restrict__ bt, const int wm, const int wn, const int lane) {
            const int kq = lane & 3;
            const int g = lane >> 2;
#pragma unroll
            for (int k4 = 0; k4 < kBk / 4; ++k4) {
                const double *ap = at + (4 * k4 + kq) * kLd + 32 * wm + g;
                const double *bp = bt + (4 * k4 + kq) * kLd + 32 * wn + g;
                double a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) { a[mi] = ap[8 * mi]; }
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) { b[ni] = bp[8 * ni]; }
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) { Dmma884(acc[mi][ni], a[mi], b[ni]); }
                }
            }
        }

        template<int XDIM>
        __global__ void __launch_bounds__(kThreads16, 1)
        PredictVarianceKernelDmma16(
            const Covariance<double> cov,
            const long n,
            const long t,
            const double *__restrict__ x_train,
            const double *__restrict__ x_test,
            const double *__restrict__ l,
            const long ldl,
            const double *__restrict__ linv,
            double *__restrict__ v_slabs,
            const long n_pad,
            double *__restrict__ sumsq) {
            using T = double;
            extern __shared__ __align__(16) unsigned char smem_raw[];
            T *as = reinterpret_cast<T *>(smem_raw);  // [2][16][kLd]
            T *bs = as + 2 * kBk * kLd;               // [2][16][kLd]
            T *wt = bs + 2 * kBk * kLd;               // [128][kLd]  W tile, row (k) major
            T *xqs = wt + kTile * kLd;                // [128][XDIM]
            T *ssq_s = xqs + kTile * 3;               // [128] column sums of the tile
            const int tid = threadIdx.x;
            const int lane = tid & 31;
            const int warp = tid >> 5;
            const int wm = warp & 3, wn = warp >> 2;  // 16 warps: rows 32 wm .. + 31, columns 32 wn .. + 31
            const int g = lane >> 2, kq = lane & 3;
            T *vs = v_slabs + static_cast<long>(blockIdx.x) * n_pad * kTile;
            const long num_panels = (n + kTile - 1) / kTile;
            const bool vec_l = (ldl & 3) == 0 && (reinterpret_cast<uintptr_t>(l) & 31) == 0;
            const long num_ct = (t + kTile - 1) / kTile;

            for (long ct = blockIdx.x; ct < num_ct; ct += gridDim.x) {
                const long col0 = ct * kTile;
                __syncthreads();
                for (int e = tid; e < kTile * XDIM; e += kThreads16) { xqs[e] = col0 * XDIM + e < t * XDIM ? x_test[col0 * XDIM + e] : T(0); }
                if (tid < kTile) { ssq_s[tid] = T(0); }
                __syncthreads();
                T ssq[4][2];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) { ssq[ni][0] = ssq[ni][1] = T(0); }

                for (long p = 0; p < num_panels; ++p) {
                    const long k0 = p * kTile;
                    T acc[4][4][2];
                    // ---- acc = -Ktest[panel rows, my points] ----
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        const long row = k0 + 32 * wm + 8 * mi + g;
                        T xi[XDIM];
#pragma unroll
                        for (int d = 0; d < XDIM; ++d) { xi[d] = row < n ? x_train[row * XDIM + d] : T(0); }
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int c = 32 * wn + 8 * ni + 2 * kq + e;
                                acc[mi][ni][e] = (row < n && col0 + c < t) ? -cov(SquaredDistance<T, XDIM>(xi, xqs + c * XDIM)) : T(0);
                            }
                        }
                    }
                    // ---- acc += L[panel rows, 0:k0] * V[0:k0, my points] ----
                    if (k0 > 0) {
                        T ra[4], rb[4];
                        LoadA16<T>(l, ldl, k0, n, 0, k0, vec_l, tid, ra);
                        LoadB16<T>(vs, n_pad, 0, tid, rb);
                        StoreA16<T>(as, tid, ra);
                        StoreB16<T>(bs, tid, rb);
                        __syncthreads();
                        const long num_kt = k0 / kBk;
                        for (long kt = 0; kt < num_kt; ++kt) {
                            const int cur = static_cast<int>(kt & 1);
                            if (kt + 1 < num_kt) {
                                LoadA16<T>(l, ldl, k0, n, (kt + 1) * kBk, k0, vec_l, tid, ra);
                                LoadB16<T>(vs, n_pad, (kt + 1) * kBk, tid, rb);
                            }
                            SlabMma16(acc, as + cur * kBk * kLd, bs + cur * kBk * kLd, wm, wn, lane);
                            if (kt + 1 < num_kt) {
                                StoreA16<T>(as + (cur ^ 1) * kBk * kLd, tid, ra);
                                StoreB16<T>(bs + (cur ^ 1) * kBk * kLd, tid, rb);
                            }
                            __syncthreads();
                        }
                    }
                    // ---- W tile -> shared, negated back ----
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) {
                            *reinterpret_cast<double2 *>(wt + (32 * wm + 8 * mi + g) * kLd + 32 * wn + 8 * ni + 2 * kq) = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
                            acc[mi][ni][0] = acc[mi][ni][1] = T(0);
                        }
                    }
                    // ---- V_p = Linv_p * W ----
                    {
                        const T *lip = linv + p * kTile * kTile;
                        T ra[4];
                        LoadA16<T>(lip, kTile, 0, kTile, 0, kTile, true, tid, ra);
                        StoreA16<T>(as, tid, ra);
                        __syncthreads();  // also publishes the W tile
                        for (int kt = 0; kt < kTile / kBk; ++kt) {
                            const int cur = kt & 1;
                            if (kt + 1 < kTile / kBk) { LoadA16<T>(lip, kTile, 0, kTile, (kt + 1) * kBk, kTile, true, tid, ra); }
                            SlabMma16(acc, as + cur * kBk * kLd, wt + kt * kBk * kLd, wm, wn, lane);
                            if (kt + 1 < kTile / kBk) { StoreA16<T>(as + (cur ^ 1) * kBk * kLd, tid, ra); }
                            __syncthreads();
                        }
                    }
                    // ---- column sums of squares; V_p -> my slab ----
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int c = 32 * wn + 8 * ni + 2 * kq + e;
                            T s = T(0);
#pragma unroll
                            for (int mi = 0; mi < 4; ++mi) { s += acc[mi][ni][e] * acc[mi][ni][e]; }
                            ssq[ni][e] += s;
                            if (p + 1 < num_panels) {
#pragma unroll
                                for (int mi = 0; mi < 4; ++mi) { vs[k0 + 32 * wm + 8 * mi + g + static_cast<long>(c) * n_pad] = acc[mi][ni][e]; }
                            }
                        }
                    }
                    __syncthreads();  // the slab rows written above are read by this CTA in the next panel
                }
                // ---- reduce the column sums over the 8 row groups of the warp (lane / 4) and the two row-halves (wm) ----
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        T s = ssq[ni][e];
                        s += __shfl_xor_sync(0xffffffffu, s, 4);
                        s += __shfl_xor_sync(0xffffffffu, s, 8);
                        s += __shfl_xor_sync(0xffffffffu, s, 16);
                        if (g == 0) { atomicAdd(ssq_s + 32 * wn + 8 * ni + 2 * kq + e, s); }
                    }
                }
                __syncthreads();
                if (tid < kTile && col0 + tid < t) { sumsq[col0 + tid] = ssq_s[tid]; }
            }
        }

        // mean[j + c * ld_out] = sum_i k(x_i, x*_j) alpha[i + c * ld_a]; one thread per test point and split of the
        // training set (blockIdx.y), partial sums combined with atomics only when the set is split
        template<typename T, int XDIM, int YMAX>
        __global__ void __launch_bounds__(128)
        PredictMeanKernel(const Covariance<T> cov, const long n, const long t, const T *__restrict__ x_train, const T *__restrict__ x_test, const T *__restrict__ alpha, const long ld_a,
                          const int y_dim, T *__restrict__ out, const long ld_out, const long rows_per_split) {
            __shared__ T xs[128 * XDIM];
            __shared__ T as_[128 * YMAX];
            const long j = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
            const long r_begin = static_cast<long>(blockIdx.y) * rows_per_split;
            const long r_end = r_begin + rows_per_split < n ? r_begin + rows_per_split : n;
            T xq[XDIM];
#pragma unroll
            for (int d = 0; d < XDIM; ++d) { xq[d] = j < t ? x_test[j * XDIM + d] : T(0); }
            T sum[YMAX];
#pragma unroll
            for (int c = 0; c < YMAX; ++c) { sum[c] = T(0); }
            for (long r0 = r_begin; r0 < r_end; r0 += 128) {
                const long cnt = r_end - r0 < 128 ? r_end - r0 : 128;
                __syncthreads();
                for (int e = threadIdx.x; e < cnt * XDIM; e += blockDim.x) { xs[e] = x_train[r0 * XDIM + e]; }
                for (int e = threadIdx.x; e < cnt * YMAX; e += blockDim.x) {
                    const int c = e / static_cast<int>(cnt), i = e % static_cast<int>(cnt);
                    as_[i * YMAX + c] = c < y_dim ? alpha[r0 + i + c * ld_a] : T(0);
                }
                __syncthreads();
                for (int i = 0; i < cnt; ++i) {
                    const T kv = cov(SquaredDistance<T, XDIM>(xs + i * XDIM, xq));
#pragma unroll
                    for (int c = 0; c < YMAX; ++c) { sum[c] += kv * as_[i * YMAX + c]; }
                }
            }
            if (j < t) {
                for (int c = 0; c < y_dim && c < YMAX; ++c) {
                    if (gridDim.y == 1) {
                        out[j + c * ld_out] = sum[c];
                    } else {
                        atomicAdd(out + j + c * ld_out, sum[c]);
                    }
                }
            }
        }

        // ERL_GP_DENSE_256=1: the 8-warp kernels (A/B measurements)
        static bool
        PredictUse16Warps() {
            static const bool on = std::getenv("ERL_GP_DENSE_256") == nullptr && std::getenv("ERL_GP_DENSE_FMA") == nullptr;
            return on;
        }

        // float: FFMA tile kernel; double: DMMA tile kernel (ERL_GP_DENSE_FMA=1 keeps the DFMA loop for A/B measurements)
        template<typename T, int XDIM>
        struct PredictVarianceSelect {
            static auto
            Get() {
                return PredictVarianceKernel<T, XDIM>;
            }
        };
        template<int XDIM>
        struct PredictVarianceSelect<double, XDIM> {
            static auto
            Get() {
                static const bool fma = std::getenv("ERL_GP_DENSE_FMA") != nullptr;
                return fma ? PredictVarianceKernel<double, XDIM> : (PredictUse16Warps() ? PredictVarianceKernelDmma16<XDIM> : PredictVarianceKernelDmma<XDIM>);
            }
        };

    }  // namespace

    template<typename T>
    size_t
    PredictVarianceSlabElems(const Context *ctx, const long n, const long t) {
        const long n_pad = CeilDiv(n, 128) * 128;
        const long num_ct = CeilDiv(t, kTile);
        const long ctas = num_ct < ctx->sm_count ? num_ct : ctx->sm_count;
        return static_cast<size_t>(ctas) * n_pad * kTile;
    }

    template<typename T>
    int
    PredictVariance(Context *ctx, int kernel, T scale, long x_dim, long n, long t, const T *x_train, const T *x_test, const T *l, long ldl, const T *linv, T *v_slabs, T *sumsq) {
        if (t <= 0) { return ERL_GP_STATUS_OK; }
        const long n_pad = CeilDiv(n, 128) * 128;
        const long num_ct = CeilDiv(t, kTile);
        const unsigned ctas = static_cast<unsigned>(num_ct < ctx->sm_count ? num_ct : ctx->sm_count);
        const size_t smem = sizeof(T) * (4 * kBk * kLd + kTile * kLd + kTile * 3 + kTile);
        const Covariance<T> cov = Covariance<T>::Make(kernel, scale);
#define ERL_GP_LAUNCH_PV(XD)                                                                                                     \
    {                                                                                                                            \
        auto kern = PredictVarianceSelect<T, XD>::Get();                                                                         \
        ERL_GP_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));    \
        kern<<<ctas, (sizeof(T) == 8 && PredictUse16Warps()) ? kThreads16 : kThreads, smem, ctx->stream>>>(cov, n, t, x_train, x_test, l, ldl, linv, v_slabs, n_pad, sumsq); \
    }
        switch (x_dim) {
            case 1: ERL_GP_LAUNCH_PV(1) break;
            case 2: ERL_GP_LAUNCH_PV(2) break;
            case 3: ERL_GP_LAUNCH_PV(3) break;
            default: return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "predict: x_dim=%ld (supported: 1, 2, 3)", x_dim);
        }
#undef ERL_GP_LAUNCH_PV
        ctx->launches += 1;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    int
    PredictMean(Context *ctx, int kernel, T scale, long x_dim, long n, long t, const T *x_train, const T *x_test, const T *alpha, long ld_a, long y_dim, T *out, long ld_out) {
        if (t <= 0) { return ERL_GP_STATUS_OK; }
        constexpr int kYmax = 4;
        if (y_dim > kYmax) { return ERL_GP_STATUS_UNSUPPORTED; }  // caller falls back to the materialised Ktest path
        const Covariance<T> cov = Covariance<T>::Make(kernel, scale);
        const long blocks_x = CeilDiv(t, 128);
        // enough CTAs to fill the machine: split the training set when there are few test points
        long splits = 1;
        while (blocks_x * splits < 2L * ctx->sm_count && CeilDiv(n, splits * 2) >= 512) { splits *= 2; }
        const long rows_per_split = CeilDiv(CeilDiv(n, splits), 128) * 128;
        splits = CeilDiv(n, rows_per_split);
        if (splits > 1) { ERL_GP_CUDA_OK(ctx, cudaMemset2DAsync(out, sizeof(T) * ld_out, 0, sizeof(T) * t, y_dim, ctx->stream)); }
        const dim3 grid(static_cast<unsigned>(blocks_x), static_cast<unsigned>(splits));
#define ERL_GP_LAUNCH_PM(XD) PredictMeanKernel<T, XD, kYmax><<<grid, 128, 0, ctx->stream>>>(cov, n, t, x_train, x_test, alpha, ld_a, static_cast<int>(y_dim), out, ld_out, rows_per_split);
        switch (x_dim) {
            case 1: ERL_GP_LAUNCH_PM(1) break;
            case 2: ERL_GP_LAUNCH_PM(2) break;
            case 3: ERL_GP_LAUNCH_PM(3) break;
            default: return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "predict: x_dim=%ld (supported: 1, 2, 3)", x_dim);
        }
#undef ERL_GP_LAUNCH_PM
        ctx->launches += 1;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        return ERL_GP_STATUS_OK;
    }

#define ERL_GP_INSTANTIATE_PREDICT(T)                                                                                                             \
    template size_t PredictVarianceSlabElems<T>(const Context *, long, long);                                                                   \
    template int PredictVariance<T>(Context *, int, T, long, long, long, const T *, const T *, const T *, long, const T *, T *, T *);            \
    template int PredictMean<T>(Context *, int, T, long, long, long, const T *, const T *, const T *, long, long, T *, long);
    ERL_GP_INSTANTIATE_PREDICT(float)
    ERL_GP_INSTANTIATE_PREDICT(double)
#undef ERL_GP_INSTANTIATE_PREDICT

}  // namespace erl_gp
