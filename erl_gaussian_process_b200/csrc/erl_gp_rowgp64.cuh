// FP64 twin of the row-GP kernel (erl_gp_rowgp.cuh) on the FP64 tensor path: mma.sync.m8n8k4.f64 (SASS DMMA), n <= 128.
//
// One CTA (256 threads, 8 warps, 2 CTAs per SM) per GP runs, for LidarGaussianProcess2D<double> / RangeSensorGaussianProcess3D<double>
// partitions and the double variant of the batched stream, what the reference runs per partition:
//   VanillaGaussianProcess::UpdateKtrain + Solve        src/vanilla_gp.cpp:476-505
//   ComputeKtest + TestResult::GetMean / GetVariance    src/vanilla_gp.cpp:521-552, 61-150
// The design is the FP32 kernel's, without the 3xTF32 split (double needs none):
//   * L lives in shared memory, column-major, packed by 16-column blocks (block b keeps rows >= 16 b) with a column stride == 2 (mod
//     16) doubles: "one row per lane", "one column per lane" and both MMA fragment patterns are bank-conflict free for 64-bit loads;
//   * the Gram entries are generated up front by all threads into L's own storage (no dependency, ~60 instructions each in double);
//   * train = left-looking blocked Cholesky in place, 16-column panels, one 16 x 16 tile per warp: (A) P = K - L[tile] L[pivot rows]^T
//     as 8 x 8 x 4 DMMAs whose reduction index is permuted (k-slot t of group e <-> column 2 t + e) so that A and B fragments are
//     single conflict-free LDS.64; (B) warp 0 factorises the 16 x 16
//     pivot tile with shuffles, lanes 16 .. 31 run the same elimination on the unit vectors and end up with the columns of its
//     inverse Dinv, z = L^-1 y rides along; (C) L_i = P_i Dinv^T straight from the accumulators: an 8 x 8 accumulator tile (row g,
//     columns 2 t, 2 t + 1) IS the A operand of two DMMAs under the same k permutation;
//   * alpha = L^-T z blocked through Dinv; L goes to HBM one column per warp and step (coalesced);
//   * predict = transposed substitution V^T = Kt^T L^-T, one warp per 8 queries with its 8 x 128 accumulators resident (64 registers):
//     V_j = X_j Dinv_j^T, X_i -= V_j L_ij^T, accumulator -> A operand without data movement, Ktest entries generated in the accumulator
//     layout, mean = k*^T alpha on the way, ||v||^2 from the finished blocks.
// Work per GP at n = 128 with 128 queries: ~6000 DMMAs (512 flop each); the DMMA pipe (37 TFLOP/s measured) bounds the C4 stream
// in double at 3.8 ms per 50 000 GPs.
#pragma once

#include "erl_gp_dense_mma.cuh"
#include "erl_gp_internal.cuh"

#include <cstdlib>
#include <type_traits>

#ifdef ERL_GP_ROWGP64_TIMING  // per-phase cycle counters, printed by one CTA (kernel experiments only)
#define ERL_GP64_TICK(acc_) { const long long now_ = clock64(); acc_ += now_ - tm_t; tm_t = now_; }
#else
#define ERL_GP64_TICK(acc_)
#endif

namespace erl_gp {
    namespace rowgp64 {

        constexpr unsigned kFull = 0xffffffffu;
        constexpr int kThreads = 256;
        constexpr int kWarps = kThreads / 32;
        constexpr int kQTile = 8 * kWarps;  // queries per pass: 8 per warp

        template<int NBLK>
        struct Layout {
            static constexpr int kNp = 16 * NBLK;

            __host__ __device__ static constexpr int
            Stride(const int cb) {  // doubles between consecutive columns of column block cb (== 2 mod 16)
                return kNp - 16 * cb + 2;
            }

            __host__ __device__ static constexpr int
            Base(const int cb) {  // first double of column block cb: sum_{b < cb} 16 * Stride(b)
                return 16 * cb * (kNp + 2) - 128 * cb * (cb - 1);
            }

            static constexpr int kDinvLd = 18;  // column stride of a 16 x 16 inverse block (== 2 mod 16)
            static constexpr int kL = 0;
            static constexpr int kPts = kL + Base(NBLK);          // [kNp][4]: x, y, z, -
            static constexpr int kAl = kPts + 4 * kNp;            // y -> z -> alpha
            static constexpr int kVar = kAl + kNp;                // noise variances
            static constexpr int kDinv = kVar + kNp;              // inverses of the 16 x 16 diagonal blocks, column-major
            static constexpr int kMisc = kDinv + NBLK * 16 * kDinvLd;
            static constexpr int kEnd = kMisc + 2;
            static constexpr size_t kBytes = static_cast<size_t>(kEnd) * sizeof(double);
        };

        template<int XDIM>
        __device__ __forceinline__ double
        Dist2(const double *__restrict__ a, const double (&b)[XDIM]) {
            double r2 = 0;
#pragma unroll
            for (int d = 0; d < XDIM; ++d) {
                const double diff = a[d] - b[d];
                r2 += diff * diff;
            }
            return r2;
        }

        // The three erl_covariance kernels in double with hand-rolled exp / sqrt.  In double a library covariance entry costs ~100 FP64
        // instructions (IEEE sqrt + exp with their special cases) and the entries (n^2 / 2 Gram + n t Ktest per GP) were half of the run
        // time of the first version of this kernel; these take ~40, accurate to a few ulp on the arguments that occur (r2 >= 0, exponent
        // argument in [-745, 0]):
        //   sqrt(r2) = r2 * rsqrt(r2) with one Newton correction (0 at r2 = 0);
        //   exp(x)   = 2^k p(r), k = rint(x log2 e), r = x - k ln2 (two-term Cody-Waite), p = Taylor polynomial of degree 12 on |r| <= 0.347
        //              (truncation error 1.7e-16), 2^k assembled in the exponent field.
        struct Cov64 {
            int type;
            double a;  // Matern32: sqrt(3) / l    OU: 1 / l    RBF: 1 / (2 l^2)

            __device__ __forceinline__ explicit Cov64(const Covariance<double> &cov) {
                type = cov.type;
                a = cov.type == ERL_GP_KERNEL_MATERN32 ? cov.c0 : 1.0 / cov.c0;
            }

            __device__ __forceinline__ static double
            Sqrt(const double r2) {
                return FastSqrt(r2);  // erl_gp_common.cuh
            }

            __device__ __forceinline__ static double
            Exp(const double x) {
                return FastExpNeg(x);
            }

            // (a __noinline__ call instead of 32 inlined copies per predict pass was measured: no gain, 16.04 vs 16.08 ms)
            __device__ __forceinline__ double
            operator()(const double r2) const {
                if (type == ERL_GP_KERNEL_RBF) { return Exp(-r2 * a); }
                const double ar = a * Sqrt(r2);
                const double e = Exp(-ar);
                return type == ERL_GP_KERNEL_MATERN32 ? fma(ar, e, e) : e;
            }
        };

        // 16 x 16 pivot tile: lanes 0 .. 15 own its rows, lanes 16 + j start from the unit vector e_j and run the very same elimination
        // (with sc = a[c] / d the update a[cc] -= sc A[cc][c] is the forward substitution of L x = e_j: the scaled entries l[c] that
        // lanes 0 .. 15 read as row r of L are, in lane 16 + j, column j of L^-1).  z = L^-1 y rides along in zacc.
        __device__ __forceinline__ void
        PivotBlock(double (&a)[16], double &zacc, double (&l)[16], const int c0, const int lane, int &fail, double *__restrict__ al) {
            // Only the reciprocal of the pivot sits on the serial chain (MUFU.RCP64H seed + two Newton steps); the square roots that turn
            // the eliminated entries into L (l = a / sqrt(d), z = zc / sqrt(d)) are taken after the loop: lane c keeps d_c, ONE rsqrt per
            // lane in parallel instead of sixteen in sequence on every lane.
            double dmine = 1.0, zraw = 0.0;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const double d = __shfl_sync(kFull, a[c], c);
                const double zc = __shfl_sync(kFull, zacc, c);
                double t[16];
#pragma unroll
                for (int cc = c + 1; cc < 16; ++cc) { t[cc] = __shfl_sync(kFull, a[c], cc); }  // A[cc][c] = A[c][cc] from lane cc
                if (!(d > 0.0) && fail == 0) { fail = c0 + c + 1; }
                double inv;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(d));
                inv = fma(inv, fma(-d, inv, 1.0), inv);
                inv = fma(inv, fma(-d, inv, 1.0), inv);
                const double sc = a[c] * inv;
#pragma unroll
                for (int cc = c + 1; cc < 16; ++cc) { a[cc] = fma(-sc, t[cc], a[cc]); }
                zacc = fma(-sc, zc, zacc);
                l[c] = a[c];  // unscaled: a[c] is final here
                if ((lane & 15) == c) { dmine = d, zraw = zc; }
            }
            const double rsv = rsqrt(dmine);  // lane c (and c + 16): 1 / sqrt(d_c)
#pragma unroll
            for (int c = 0; c < 16; ++c) { l[c] *= __shfl_sync(kFull, rsv, c); }
            if (lane < 16) { al[c0 + lane] = zraw * rsv; }
        }

        // Gram matrix straight into L's own packed storage (block b: rows >= 16 b of its 16 columns, the diagonal tile in full): in double
        // a covariance entry costs ~60 instructions (exp + sqrt), far more than its share of the factorisation, and has no dependency
        // on anything - all 256 threads generate them up front instead of the panel's warps on the critical path of the factorisation.
        // K[i][i] = 1 + var[i]; rows / columns >= n are identity padding.
        template<int XDIM, int NBLK>
        __device__ __forceinline__ void
        FillGram(const Cov64 &cov, double *__restrict__ smem, const int n, const int nblk) {
            using Lay = Layout<NBLK>;
            double *lp = smem + Lay::kL;
            const double *pts = smem + Lay::kPts;
            const double *sv = smem + Lay::kVar;
            const int npr = 16 * nblk;
            for (int cb = 0; cb < nblk; ++cb) {
                const int rows = npr - 16 * cb;  // rows of this column block (multiple of 16)
                double *blk = lp + Lay::Base(cb);
                const int stride = Lay::Stride(cb);
                for (int e = threadIdx.x; e < rows * 16; e += kThreads) {
                    const int c = e / rows, r = e - c * rows;  // consecutive threads -> consecutive rows of one column
                    const int row = 16 * cb + r, col = 16 * cb + c;
                    double kv = 0.0;
                    if (row == col) {
                        kv = row < n ? 1.0 + sv[row] : 1.0;
                    } else if (row < n && col < n) {
                        double xr[XDIM];
#pragma unroll
                        for (int d = 0; d < XDIM; ++d) { xr[d] = pts[4 * row + d]; }
                        kv = cov(Dist2<XDIM>(pts + 4 * col, xr));
                    }
                    blk[c * stride + r] = kv;
                }
            }
        }

        // blocked left-looking Cholesky, in place: the panels hold the Gram entries on entry (FillGram)
        template<int XDIM, int NBLK>
        __device__ __forceinline__ int
        Factorize(double *__restrict__ smem, const int n, const int nblk) {
            using Lay = Layout<NBLK>;
            double *lp = smem + Lay::kL;
            double *al = smem + Lay::kAl;
            double *dinv = smem + Lay::kDinv;
            (void) n;
            const int tid = threadIdx.x;
            const int warp = __shfl_sync(kFull, tid >> 5, 0);  // warp-uniform by construction
            const int lane = tid & 31;
            const int g = lane >> 2, t = lane & 3;
            constexpr int kTpw = (NBLK + kWarps - 1) / kWarps;  // 16-row tiles of a panel per warp
            int fail = 0;
#ifdef ERL_GP_ROWGP64_TIMING
            long long tm_a = 0, tm_b = 0, tm_w1 = 0, tm_c = 0, tm_w2 = 0, tm_t = clock64();
#endif
            for (int kb = 0; kb < nblk; ++kb) {
                const int c0 = 16 * kb;
                const int mt = nblk - kb;  // 16-row tiles of this panel: tile w + 8 ti -> warp w (one tile per warp up to n = 128, two beyond)
                const int stride_k = Lay::Stride(kb);
                double *panel = lp + Lay::Base(kb);  // element (row c0, column c0)
                double acc_t[kTpw][2][2][2];
#pragma unroll
                for (int ti = 0; ti < kTpw; ++ti) {
                double (&acc)[2][2][2] = acc_t[ti];
                const int rel = 16 * (warp + kWarps * ti);  // my tile's first row, relative to c0
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) { acc[rh][ch][0] = acc[rh][ch][1] = 0.0; }
                }
                if (warp + kWarps * ti < mt) {
                    // ---- A: P = K[tile, panel] - L[tile, 0:c0] L[pivot rows, 0:c0]^T ----
                    for (int jb = 0; jb < kb; ++jb) {
                        const int stride = Lay::Stride(jb);
                        const double *blk = lp + Lay::Base(jb) + (c0 - 16 * jb) + g;  // (row c0 + g, column 16 jb)
#pragma unroll
                        for (int ck = 0; ck < 2; ++ck) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const double *col = blk + (8 * ck + 2 * t + e) * stride;
                                const double b0 = col[0], b1 = col[8];
                                const double a0 = col[rel], a1 = col[rel + 8];
                                Dmma884(acc[0][0], a0, b0);
                                Dmma884(acc[0][1], a0, b1);
                                Dmma884(acc[1][0], a1, b0);
                                Dmma884(acc[1][1], a1, b1);
                            }
                        }
                    }
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh) {  // P = K - update: the Gram entries wait in the panel's own storage (FillGram)
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) { acc[rh][ch][e] = panel[(8 * ch + 2 * t + e) * stride_k + rel + 8 * rh + g] - acc[rh][ch][e]; }
                        }
                    }
                }
                }
                ERL_GP64_TICK(tm_a)
                // ---- B: pivot tile (warp 0 holds it) ----
                if (warp == 0) {
                    double (&acc)[2][2][2] = acc_t[0];
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) { panel[(8 * ch + 2 * t + e) * stride_k + 8 * rh + g] = acc[rh][ch][e]; }
                        }
                    }
                    const int r = lane & 15, hh = lane >> 4;
                    // z: sum_{j < c0} L[c0 + r][j] z_j (two lanes per row: even / odd column blocks)
                    double zp0 = 0.0, zp1 = 0.0;
                    for (int jb = hh; jb < kb; jb += 2) {
                        const int stride = Lay::Stride(jb);
                        const double *rowp = lp + Lay::Base(jb) + (c0 + r - 16 * jb);
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            zp0 = fma(rowp[j * stride], al[16 * jb + j], zp0);
                            zp1 = fma(rowp[(j + 1) * stride], al[16 * jb + j + 1], zp1);
                        }
                    }
                    double zs = zp0 + zp1;
                    zs += __shfl_xor_sync(kFull, zs, 16);
                    __syncwarp();
                    double zacc = lane < 16 ? al[c0 + r] - zs : 0.0;  // al[c0 + r] still holds y
                    double prow[16], l[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const double pv = panel[c * stride_k + r];
                        prow[c] = lane < 16 ? pv : (c == r ? 1.0 : 0.0);
                    }
                    __syncwarp();  // everybody has read y and the raw tile
                    PivotBlock(prow, zacc, l, c0, lane, fail, al);
                    if (lane < 16) {
#pragma unroll
                        for (int c = 0; c < 16; ++c) { panel[c * stride_k + r] = c > r ? 0.0 : l[c]; }
                    } else {
                        double *dst = dinv + kb * 16 * Lay::kDinvLd + r * Lay::kDinvLd;  // column r of Dinv (zero above the diagonal)
#pragma unroll
                        for (int c = 0; c < 16; ++c) { dst[c] = l[c]; }
                    }
                }
                ERL_GP64_TICK(tm_b)
                __syncthreads();  // #1: pivot tile, Dinv, z of this panel are published
                ERL_GP64_TICK(tm_w1)
                if (mt > 1) {
#pragma unroll
                    for (int ti = 0; ti < kTpw; ++ti) {
                    double (&acc)[2][2][2] = acc_t[ti];
                    const int rel = 16 * (warp + kWarps * ti);
                    if (warp + kWarps * ti > 0 && warp + kWarps * ti < mt) {
                        // ---- C: L_i = P_i Dinv^T; Dinv is lower triangular: the (ck = 1, ch = 0) block is zero ----
                        const double *dv = dinv + kb * 16 * Lay::kDinvLd + g;
                        double out[2][2][2];
#pragma unroll
                        for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
                            for (int ch = 0; ch < 2; ++ch) { out[rh][ch][0] = out[rh][ch][1] = 0.0; }
                        }
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
                            for (int ck = 0; ck <= ch; ++ck) {
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const double b = dv[(8 * ck + 2 * t + e) * Lay::kDinvLd + 8 * ch];  // Dinv[8 ch + g][8 ck + 2 t + e]
                                    Dmma884(out[0][ch], acc[0][ck][e], b);
                                    Dmma884(out[1][ch], acc[1][ck][e], b);
                                }
                            }
                        }
#pragma unroll
                        for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
                            for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
                                for (int e = 0; e < 2; ++e) { panel[(8 * ch + 2 * t + e) * stride_k + rel + 8 * rh + g] = out[rh][ch][e]; }
                            }
                        }
                    }
                    }
                    ERL_GP64_TICK(tm_c)
                    __syncthreads();  // #2: the whole panel is published
                    ERL_GP64_TICK(tm_w2)
                }
            }
#ifdef ERL_GP_ROWGP64_TIMING
            if (blockIdx.x == 20000 && lane == 0 && warp < 3) { printf("rowgp64 cta %d warp %d: A %lld  B %lld  wait1 %lld  C %lld  wait2 %lld\n", blockIdx.x, warp, tm_a, tm_b, tm_w1, tm_c, tm_w2); }
#endif
            return fail;
        }

        // alpha = L^-T z (al holds z on entry, alpha on exit); thread = column, blocked from the bottom through Dinv
        template<int NBLK>
        __device__ __forceinline__ void
        BackSolve(double *__restrict__ smem, const int nblk) {
            using Lay = Layout<NBLK>;
            const double *lp = smem + Lay::kL;
            double *al = smem + Lay::kAl;
            const double *dinv = smem + Lay::kDinv;
            const int tid = threadIdx.x;
            const int warp = __shfl_sync(kFull, tid >> 5, 0);
            const int lane = tid & 31;
            double s = 0.0;  // sum_{r > my block} L[r][tid] alpha[r]
            for (int kb = nblk - 1; kb >= 0; --kb) {
                const int c0 = 16 * kb;
                if (warp == (c0 >> 5)) {
                    const int lb = c0 & 31;
                    const bool mine = lane >= lb && lane < lb + 16;
                    const int jj = mine ? lane - lb : 0;
                    const double vj = al[c0 + jj] - s;
                    const double *dcol = dinv + kb * 16 * Lay::kDinvLd + jj * Lay::kDinvLd;  // column jj of Dinv
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int r = 0; r < 16; r += 2) {  // alpha_blk = Dinv^T (z_blk - s_blk); rows r < jj of the column are zero
                        a0 = fma(dcol[r], __shfl_sync(kFull, vj, lb + r), a0);
                        a1 = fma(dcol[r + 1], __shfl_sync(kFull, vj, lb + r + 1), a1);
                    }
                    if (mine) { al[c0 + jj] = a0 + a1; }
                }
                __syncthreads();
                if (tid < c0) {
                    const int cb = tid >> 4;
                    const double *colp = lp + Lay::Base(cb) + (tid & 15) * Lay::Stride(cb) + (c0 - 16 * cb);
#pragma unroll
                    for (int k = 0; k < 16; ++k) { s = fma(colp[k], al[c0 + k], s); }
                }
            }
        }

        // inverses of the 16 x 16 diagonal blocks (predict-only mode: L was reloaded from HBM): lane c < 16 solves L x = e_c
        template<int NBLK>
        __device__ __forceinline__ void
        ComputeDinv(double *__restrict__ smem, const int nblk) {
            using Lay = Layout<NBLK>;
            const double *lp = smem + Lay::kL;
            double *dinv = smem + Lay::kDinv;
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            if (lane >= 16) { return; }
            for (int kb = warp; kb < nblk; kb += kWarps) {
                const double *blk = lp + Lay::Base(kb);
                const int stride = Lay::Stride(kb);
                double x[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    double sum = r == lane ? 1.0 : 0.0;
#pragma unroll
                    for (int k = 0; k < r; ++k) { sum = fma(-blk[k * stride + r], x[k], sum); }  // L[r][k] x[k]
                    x[r] = r < lane ? 0.0 : sum / blk[r * stride + r];
                }
                double *dst = dinv + kb * 16 * Lay::kDinvLd + lane * Lay::kDinvLd;
#pragma unroll
                for (int r = 0; r < 16; ++r) { dst[r] = x[r]; }
            }
        }

        // compile-time loop: the body sees its index as a constant, so that register arrays indexed by it stay in registers whatever the
        // unroller's size limits are (with #pragma unroll the 66 block updates of the n <= 192 predict were only partly unrolled and its
        // accumulators went to local memory)
        template<int I, int N, typename F>
        __device__ __forceinline__ void
        StaticFor(F &&f) {
            if constexpr (I < N) {
                f(std::integral_constant<int, I>{});
                StaticFor<I + 1, N>(f);
            }
        }

        // one pass: 8 queries per warp, V^T = Kt^T L^-T with the 8 x (16 NBLK) accumulators resident
        template<int XDIM, int NBLK>
        __device__ __forceinline__ void
        PredictTile(const BatchParams<double> &p, const double *__restrict__ smem, const int n, const int nblk, const long q_begin, const int nq) {
            using Lay = Layout<NBLK>;
            const double *lp = smem + Lay::kL;
            const double *pts = smem + Lay::kPts;
            const double *al = smem + Lay::kAl;
            const double *dinv = smem + Lay::kDinv;
            const int tid = threadIdx.x;
            const int warp = __shfl_sync(kFull, tid >> 5, 0);
            const int lane = tid & 31;
            const int g = lane >> 2, t = lane & 3;
            if (8 * warp >= nq) { return; }  // (no barrier below)
            const int qi = 8 * warp + g;
            const bool active = qi < nq;
            double xq[XDIM];
#pragma unroll
            for (int d = 0; d < XDIM; ++d) { xq[d] = active ? p.q_x[(q_begin + qi) * XDIM + d] : 0.0; }
            // Ktest entries in the accumulator layout: tile ct holds (query g, training points 8 ct + 2 t, 8 ct + 2 t + 1)
            const Cov64 cov(p.cov);
            double x[2 * NBLK][2];
            double mean = 0.0;
#pragma unroll
            for (int ct = 0; ct < 2 * NBLK; ++ct) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 8 * ct + 2 * t + e;
                    double kv = 0.0;
                    if (ct < 2 * nblk && col < n) {
                        kv = cov(Dist2<XDIM>(pts + 4 * col, xq));
                        mean = fma(kv, al[col], mean);
                    }
                    x[ct][e] = kv;
                }
            }
            double ss = 0.0;
            StaticFor<0, NBLK>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                if (j < nblk) {
                    // V_j = X_j Dinv_j^T
                    const double *dv = dinv + j * 16 * Lay::kDinvLd + g;
                    double v[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
                        for (int ck = 0; ck <= ch; ++ck) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) { Dmma884(v[ch], x[2 * j + ck][e], dv[(8 * ck + 2 * t + e) * Lay::kDinvLd + 8 * ch]); }
                        }
                    }
                    ss = fma(v[0][0], v[0][0], ss);
                    ss = fma(v[0][1], v[0][1], ss);
                    ss = fma(v[1][0], v[1][0], ss);
                    ss = fma(v[1][1], v[1][1], ss);
                    const double nv[2][2] = {{-v[0][0], -v[0][1]}, {-v[1][0], -v[1][1]}};
                    // X_i -= V_j L_ij^T for the block rows below
                    const int stride = Lay::Stride(j);
                    const double *base = lp + Lay::Base(j) + g;
                    StaticFor<j + 1, NBLK>([&](auto ic) {
                        constexpr int i = decltype(ic)::value;
                        if (i < nblk) {
#pragma unroll
                            for (int ck = 0; ck < 2; ++ck) {
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const double *col = base + (8 * ck + 2 * t + e) * stride + 16 * (i - j);  // L[16 i + g (+ 8)][16 j + 8 ck + 2 t + e]
                                    Dmma884(x[2 * i], nv[ck][e], col[0]);
                                    Dmma884(x[2 * i + 1], nv[ck][e], col[8]);
                                }
                            }
                        }
                    });
                }
            });
            mean += __shfl_xor_sync(kFull, mean, 1);
            mean += __shfl_xor_sync(kFull, mean, 2);
            ss += __shfl_xor_sync(kFull, ss, 1);
            ss += __shfl_xor_sync(kFull, ss, 2);
            if (t == 0 && active) {
                const long src = q_begin + qi;
                const long dst = p.q_out_index != nullptr ? p.q_out_index[src] : src;
                if (p.mean != nullptr) {
                    double f = mean;
                    if (p.mapping != ERL_GP_MAPPING_NONE) { f = MappingInv<double>(p.mapping, p.mapping_scale, f); }
                    p.mean[dst] = f;
                }
                if (p.variance != nullptr) { p.variance[dst] = 1.0 - ss; }  // literal prior 1.0f, src/vanilla_gp.cpp:121
                if (p.valid != nullptr) { p.valid[dst] = 1; }
            }
        }

        // ------------------------------------------------------------------------------------------------------------
        // The 128 x 128 diagonal block of the dense blocked Cholesky (erl_gp_dense.cu: Potrf keeps the INVERSE of every diagonal block so
        // that the panel solve and every later triangular solve are GEMMs).  Same machinery as a partition GP: the tile is loaded into the
        // packed layout, factorised in place by Factorize (DMMA updates, shuffle pivot tiles, Dinv of the 16 x 16 blocks), and the full
        // inverse is the transposed substitution of the predict applied to the identity: V^T = I L^-T, so the warp that owns the eight
        // "queries" e_q ... e_{q+7} ends with columns q ... q + 7 of L^-1 (zero above the diagonal: the walk starts at q's own block).
        // ------------------------------------------------------------------------------------------------------------
        template<int NBLK>
        __device__ __forceinline__ void
        InverseFromIdentity(const double *__restrict__ smem, const int nblk, double *__restrict__ linv, const int ld_inv) {
            using Lay = Layout<NBLK>;
            const double *lp = smem + Lay::kL;
            const double *dinv = smem + Lay::kDinv;
            const int tid = threadIdx.x;
            const int warp = __shfl_sync(kFull, tid >> 5, 0);
            const int lane = tid & 31;
            const int g = lane >> 2, t = lane & 3;
            for (int pass = 0; pass < 2 * NBLK / kWarps; ++pass) {
                const int q0 = 8 * (kWarps * pass + warp);  // my eight columns of the inverse
                if (q0 >= 16 * nblk) { continue; }
                const int jq = q0 >> 4;
                double x[2 * NBLK][2];
#pragma unroll
                for (int ct = 0; ct < 2 * NBLK; ++ct) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) { x[ct][e] = (8 * ct + 2 * t + e == q0 + g) ? 1.0 : 0.0; }
                }
#pragma unroll
                for (int j = 0; j < NBLK; ++j) {
                    if (j >= jq && j < nblk) {
                        const double *dv = dinv + j * 16 * Lay::kDinvLd + g;
                        double v[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
                            for (int ck = 0; ck <= ch; ++ck) {
#pragma unroll
                                for (int e = 0; e < 2; ++e) { Dmma884(v[ch], x[2 * j + ck][e], dv[(8 * ck + 2 * t + e) * Lay::kDinvLd + 8 * ch]); }
                            }
                        }
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) { linv[(16 * j + 8 * ch + 2 * t + e) + (q0 + g) * ld_inv] = v[ch][e]; }  // Linv[row][column q0 + g]
                        }
                        const double nv[2][2] = {{-v[0][0], -v[0][1]}, {-v[1][0], -v[1][1]}};
                        const int stride = Lay::Stride(j);
                        const double *base = lp + Lay::Base(j) + g;
#pragma unroll
                        for (int i = j + 1; i < NBLK; ++i) {
                            if (i < nblk) {
#pragma unroll
                                for (int ck = 0; ck < 2; ++ck) {
#pragma unroll
                                    for (int e = 0; e < 2; ++e) {
                                        const double *col = base + (8 * ck + 2 * t + e) * stride + 16 * (i - j);
                                        Dmma884(x[2 * i], nv[ck][e], col[0]);
                                        Dmma884(x[2 * i + 1], nv[ck][e], col[8]);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }

        // a: nk x nk (ld) lower triangle of the diagonal block, factorised in place; linv: 128 x 128 (ld 128), identity-padded inverse
        template<int NBLK>
        __global__ void __launch_bounds__(kThreads, 1)
        DiagFactor64Kernel(double *__restrict__ a, const long ld, const int nk, double *__restrict__ linv, int *__restrict__ info, const int col_offset) {
            static_assert(NBLK == 8, "128 x 128 diagonal blocks");
            using Lay = Layout<NBLK>;
            extern __shared__ __align__(16) unsigned char smem_raw[];
            double *smem = reinterpret_cast<double *>(smem_raw);
            double *lp = smem + Lay::kL;
            double *al = smem + Lay::kAl;
            int *s_fail = reinterpret_cast<int *>(smem + Lay::kMisc);
            const int tid = threadIdx.x;
            const int warp = tid >> 5, lane = tid & 31;
            const int nblk = (nk + 15) >> 4;
            const int npr = 16 * nblk;
            // the tile into the packed layout (column block cb keeps rows >= 16 cb; the diagonal tiles in full: mirror entries)
            for (int cb = 0; cb < nblk; ++cb) {
                double *blk = lp + Lay::Base(cb);
                const int stride = Lay::Stride(cb);
                const int rows = npr - 16 * cb;
                for (int c = warp; c < 16; c += kWarps) {
                    const int col = 16 * cb + c;
                    for (int r = lane; r < rows; r += 32) {
                        const int row = 16 * cb + r;
                        double v = row == col ? 1.0 : 0.0;
                        if (row < nk && col < nk) { v = row >= col ? a[row + static_cast<long>(col) * ld] : a[col + static_cast<long>(row) * ld]; }
                        blk[c * stride + r] = v;
                    }
                }
            }
            for (int e = tid; e < Lay::kNp; e += kThreads) { al[e] = 0.0; }  // (the z that rides along with the pivot tiles is not used here)
            __syncthreads();
            const int fail = Factorize<1, NBLK>(smem, nk, nblk);
            if (tid == 0) { *s_fail = fail; }
            __syncthreads();
            if (*s_fail != 0 && tid == 0 && *info == 0) { *info = col_offset + *s_fail; }
            for (int c = warp; c < nk; c += kWarps) {
                const int cb = c >> 4;
                const double *col = lp + Lay::Base(cb) + (c & 15) * Lay::Stride(cb) - 16 * cb;
                for (int r = lane; r < nk; r += 32) { a[r + static_cast<long>(c) * ld] = r >= 16 * cb ? col[r] : 0.0; }
            }
            for (int e = tid; e < 128 * 128; e += kThreads) { linv[e] = ((e & 127) == (e >> 7) && (e >> 7) >= npr) ? 1.0 : 0.0; }
            __syncthreads();
            InverseFromIdentity<NBLK>(smem, nblk, linv, 128);
        }

        inline size_t
        DiagFactor64SmemBytes() {
            return Layout<8>::kBytes;
        }

        template<int XDIM, int NBLK, int MODE>
        __global__ void __launch_bounds__(kThreads, NBLK <= 8 ? 2 : 1)
        RowGp64Kernel(const BatchParams<double> p) {
            using Lay = Layout<NBLK>;
            extern __shared__ __align__(16) unsigned char smem_raw[];
            double *smem = reinterpret_cast<double *>(smem_raw);
            double *lp = smem + Lay::kL;
            double *pts = smem + Lay::kPts;
            double *al = smem + Lay::kAl;
            double *sv = smem + Lay::kVar;
            int *s_fail = reinterpret_cast<int *>(smem + Lay::kMisc);
            const int g = blockIdx.x;
            const int tid = threadIdx.x;
            const int warp = __shfl_sync(kFull, tid >> 5, 0);
            const int lane = tid & 31;
            const int n = p.n_train[g];
            const long q0 = (MODE & kBatchPredict) ? p.q_offsets[g] : 0;
            const long q1 = (MODE & kBatchPredict) ? p.q_offsets[g + 1] : 0;
            auto invalidate = [&]() {
                if ((MODE & kBatchPredict) && p.valid != nullptr) {
                    for (long q = q0 + tid; q < q1; q += kThreads) { p.valid[p.q_out_index != nullptr ? p.q_out_index[q] : q] = 0; }
                }
            };
            if constexpr ((MODE & kBatchTrain) != 0) {
                if (n <= p.min_train || n <= 0) {  // the reference's `cnt > min_num_samples_per_group` / `cnt > 0` gate
                    if (tid == 0) { p.info[g] = -1; }
                    invalidate();
                    return;
                }
            } else {
                if (p.info[g] != 0 || q1 <= q0) { return; }  // untrained / failed GP: outputs stay untouched
            }
            const int nblk = (n + 15) >> 4;
            const int npr = 16 * nblk;
            const double *gx = p.x + static_cast<long>(g) * p.max_n * XDIM;
            for (int e = tid; e < Lay::kNp; e += kThreads) {
#pragma unroll
                for (int d = 0; d < 3; ++d) { pts[4 * e + d] = (d < XDIM && e < n) ? gx[e * XDIM + (d < XDIM ? d : 0)] : 0.0; }
            }
            if constexpr ((MODE & kBatchTrain) != 0) {
                const double *gy = p.y + static_cast<long>(g) * p.max_n;
                const double *gv = p.var + static_cast<long>(g) * p.max_n;
                for (int e = tid; e < Lay::kNp; e += kThreads) {
                    al[e] = e < n ? gy[e] : 0.0;
                    sv[e] = e < n ? gv[e] : 0.0;
                }
                __syncthreads();
#ifdef ERL_GP_ROWGP64_TIMING
                long long tm_fill = 0, tm_fact = 0, tm_wb = 0, tm_bs = 0, tm_t = clock64();
#endif
                FillGram<XDIM, NBLK>(Cov64(p.cov), smem, n, nblk);
                __syncthreads();
                ERL_GP64_TICK(tm_fill)
                const int fail = Factorize<XDIM, NBLK>(smem, n, nblk);
                ERL_GP64_TICK(tm_fact)
                if (tid == 0) { *s_fail = fail; }  // warp 0 tracked every pivot
                __syncthreads();
                const int failed = *s_fail;
                if (failed != 0) {
                    if (tid == 0) { p.info[g] = failed; }
                    invalidate();
                    return;
                }
                if (p.write_l && p.tma_writeback && ((p.max_n | n) & 1) == 0) {
                    // one bulk asynchronous copy (TMA engine) per column, rows 16 cb .. n: see the FP32 kernel (erl_gp_rowgp.cuh); the
                    // slice is zero above the diagonal blocks from erl_gp_batch_create on
                    double *gl = p.l + static_cast<long>(g) * p.max_n * p.max_n;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    for (int c = tid; c < n; c += kThreads) {
                        const int cb = c >> 4;
                        const uint32_t src = static_cast<uint32_t>(__cvta_generic_to_shared(lp + Lay::Base(cb) + (c & 15) * Lay::Stride(cb)));
                        double *dst = gl + static_cast<long>(c) * p.max_n + 16 * cb;
                        const uint32_t bytes = static_cast<uint32_t>(n - 16 * cb) * 8u;
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                } else if (p.write_l) {  // L write-back: one column per warp and step, rows along the lanes
                    double *gl = p.l + static_cast<long>(g) * p.max_n * p.max_n;
                    for (int c = warp; c < n; c += kWarps) {
                        const int cb = c >> 4;
                        const double *col = lp + Lay::Base(cb) + (c & 15) * Lay::Stride(cb) - 16 * cb;
                        double *gcol = gl + static_cast<long>(c) * p.max_n;
                        for (int r = lane; r < n; r += 32) { gcol[r] = r >= 16 * cb ? col[r] : 0.0; }
                    }
                }
                ERL_GP64_TICK(tm_wb)
                BackSolve<NBLK>(smem, nblk);
                __syncthreads();
                ERL_GP64_TICK(tm_bs)
#ifdef ERL_GP_ROWGP64_TIMING
                if (blockIdx.x == 20000 && tid == 0) { printf("rowgp64 cta %d: fill %lld  factorize %lld  write-back %lld  back-solve %lld\n", blockIdx.x, tm_fill, tm_fact, tm_wb, tm_bs); }
#endif
                double *ga = p.alpha + static_cast<long>(g) * p.max_n;
                for (int e = tid; e < n; e += kThreads) { ga[e] = al[e]; }
                if (tid == 0) { p.info[g] = 0; }
            } else {
                // ---- predict-only: reload L and alpha, rebuild the inverses of the diagonal blocks ----
                const double *gl = p.l + static_cast<long>(g) * p.max_n * p.max_n;
                const double *ga = p.alpha + static_cast<long>(g) * p.max_n;
                for (int e = tid; e < Lay::kNp; e += kThreads) { al[e] = e < n ? ga[e] : 0.0; }
                for (int c = warp; c < npr; c += kWarps) {
                    const int cb = c >> 4;
                    double *col = lp + Lay::Base(cb) + (c & 15) * Lay::Stride(cb) - 16 * cb;
                    const double *gcol = gl + static_cast<long>(c) * p.max_n;
                    for (int r = 16 * cb + lane; r < npr; r += 32) {
                        double v = 0.0;
                        if (r < n && c < n) {
                            v = r >= c ? gcol[r] : 0.0;
                        } else if (r == c) {
                            v = 1.0;  // identity padding
                        }
                        col[r] = v;
                    }
                }
                __syncthreads();
                ComputeDinv<NBLK>(smem, nblk);
            }
            if constexpr ((MODE & kBatchPredict) != 0) {
                __syncthreads();
#ifdef ERL_GP_ROWGP64_TIMING
                const long long tp0 = clock64();
#endif
                for (long qb = q0 + static_cast<long>(blockIdx.y) * kQTile; qb < q1; qb += static_cast<long>(gridDim.y) * kQTile) {
                    const int nq = static_cast<int>(q1 - qb < kQTile ? q1 - qb : kQTile);
                    PredictTile<XDIM, NBLK>(p, smem, n, nblk, qb, nq);
                }
#ifdef ERL_GP_ROWGP64_TIMING
                if (blockIdx.x == 20000 && (tid & 31) == 0 && tid < 96) { printf("rowgp64 cta %d warp %d: predict %lld\n", blockIdx.x, tid >> 5, clock64() - tp0); }
#endif
            }
            if constexpr ((MODE & kBatchTrain) != 0) {
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the bulk copies of L read this CTA's shared memory
            }
        }

        template<int XDIM, int NBLK, int MODE>
        static int
        LaunchInstance(Context *ctx, const BatchParams<double> &params, const int tiles_per_gp) {
            using Lay = Layout<NBLK>;
            auto kernel = RowGp64Kernel<XDIM, NBLK, MODE>;
            if (static_cast<int>(Lay::kBytes) > ctx->max_smem_optin) {
                return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "FP64 row-GP kernel needs %zu B of shared memory, device allows %d", Lay::kBytes, ctx->max_smem_optin);
            }
            ERL_GP_CUDA_OK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(Lay::kBytes)));
            const dim3 grid(static_cast<unsigned>(params.num_gps), static_cast<unsigned>(tiles_per_gp < 1 ? 1 : tiles_per_gp));
            BatchParams<double> launch_params = params;
            {
                static const char *env = std::getenv("ERL_GP_ROWGP_TMA_WB");
                launch_params.tma_writeback = env != nullptr ? std::atoi(env) : 1;
            }
            kernel<<<grid, kThreads, Lay::kBytes, ctx->stream>>>(launch_params);
            ctx->launches += 1;
            ERL_GP_CUDA_OK(ctx, cudaGetLastError());
            return ERL_GP_STATUS_OK;
        }

        template<int XDIM, int NBLK>
        static int
        LaunchMode(Context *ctx, const BatchParams<double> &params, const int mode, const int tiles_per_gp) {
            switch (mode) {
                case kBatchTrain: return LaunchInstance<XDIM, NBLK, kBatchTrain>(ctx, params, 1);
                case kBatchPredict: return LaunchInstance<XDIM, NBLK, kBatchPredict>(ctx, params, tiles_per_gp);
                case kBatchTrainPredict: return LaunchInstance<XDIM, NBLK, kBatchTrainPredict>(ctx, params, 1);
                default: return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "batch: bad mode %d", mode);
            }
        }

        // max_n <= 192 (n in (128, 192]: one CTA per SM, 200 KB of shared memory, two 16-row tiles per warp in the factorisation)
        template<int XDIM>
        int
        Launch(Context *ctx, const BatchParams<double> &params, const int mode, const int tiles_per_gp) {
            if (params.max_n <= 64) { return LaunchMode<XDIM, 4>(ctx, params, mode, tiles_per_gp); }
            if (params.max_n <= 128) { return LaunchMode<XDIM, 8>(ctx, params, mode, tiles_per_gp); }
            return LaunchMode<XDIM, 12>(ctx, params, mode, tiles_per_gp);
        }

    }  // namespace rowgp64
}  // namespace erl_gp
