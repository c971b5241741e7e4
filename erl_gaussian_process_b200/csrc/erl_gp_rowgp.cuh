// FP32 one-CTA-per-GP batched train / predict kernel for n <= 256 ("row GP" kernel), sm_100a.
//
// Same contract as BatchedGpKernel (erl_gp_batched.cuh) and the same reference code replaced:
//   VanillaGaussianProcess::UpdateKtrain + Solve        src/vanilla_gp.cpp:476-505
//   VanillaGaussianProcess::ComputeKtest                src/vanilla_gp.cpp:521-552
//   TestResult::GetMean / GetVariance                   src/vanilla_gp.cpp:61-150
// driven per partition by src/lidar_gp_2d.cpp:366-392 / src/range_sensor_gp_3d.cpp:334-360.
//
// Design (measured history in DESIGN.md 4.1):
//   * 128 threads per GP; n <= 128: 55 KB of shared memory and 128 registers, 4 CTAs per SM, so that the serial parts of one GP
//     (pivot blocks, back-substitution) overlap with the tensor-pipe parts of the others;
//   * the Gram matrix is never stored: the entries of a 16-column panel are generated in the MMA accumulator layout right
//     when the panel is updated (fused distance + covariance + noise diagonal, two entries per packed f32x2 operation);
//   * train (FactorizeMma): left-looking blocked Cholesky, 16-column panels, 16-row tiles, every product a 3xTF32
//     mma.sync.m16n8k8; the 16 x 16 pivot block is factorised by one warp with shuffles while its lanes 16 .. 31 eliminate
//     the unit vectors (= the inverse of the block, for free); the tiles below become L = P Dinv^T straight from the
//     accumulators; z = L^-1 y rides along, alpha = L^-T z is a blocked back-substitution through the block inverses;
//   * L lives in shared memory column-major, packed by 16-column blocks (column block b keeps rows >= 16 b) with a 4-float
//     pad per column, so that "one row per lane", "one column per lane", warp-uniform float4 and both MMA fragment patterns
//     (with the k-slot permutation slot t <-> column 2t, slot t + 4 <-> column 2t + 1) are bank-conflict free;
//   * predict (PredictTileMma): transposed substitution V^T = Kt^T L^-T, one warp per 16 queries, all 16 x n accumulators
//     resident, finished blocks reused as A operands without data movement, no barrier; mean = k*^T alpha and ||v||^2 on the way.
// The FFMA2 versions of both halves (Factorize, PredictTile; n <= 128) are kept for A/B builds
// (-DERL_GP_ROWGP_FFMA_TRAIN / _PREDICT); they were designed from the measurements in tools/fma_lds_rate.cu (a plain FFMA with
// two fresh register operands sustains ~37 TFLOP/s, FFMA2 with a scalar-broadcast operand ~55 when every LDS.128 feeds >= 8 FMAs).
// HBM traffic per GP is the algorithmic minimum (x, y, var in; L, alpha out; queries in; mean / var out).
#pragma once

#include "erl_gp_internal.cuh"

#include <cstdlib>
#include <type_traits>

namespace erl_gp {
    namespace rowgp {

        // compile-time loop: the body sees its index as a constant, so register-resident arrays stay in registers
        // however long the unrolled body gets (#pragma unroll gives up on the 128-column substitution)
        template<int I, int N, typename F>
        __device__ __forceinline__ void
        StaticFor(F &&f) {
            if constexpr (I < N) {
                f(std::integral_constant<int, I>{});
                StaticFor<I + 1, N>(f);
            }
        }

        constexpr int kThreads = 128;  // n <= 192 (and the FFMA2 versions)

        // threads per CTA of an instance: the n <= 256 instances need 171 KB of shared memory (one CTA per SM), so they run 8 warps
        template<int NBLK>
        struct ThreadsFor {
            static constexpr int value = NBLK > 12 ? 256 : 128;
        };
        constexpr unsigned kFull = 0xffffffffu;
        constexpr int kDefaultStaggerCycles = 0;  // per CTA slot, see RowGpKernel
        constexpr int kDefaultTmaWriteback = 1;   // ERL_GP_ROWGP_TMA_WB: L to HBM by bulk asynchronous copies (TMA), see RowGpKernel
        // A/B switches (measured on the B200, C4, fused / train-only ms): z-dot hoisted before the update 5.80 / 2.76, in phase B
        // 5.56 / 2.74; back-substitution through Dinv 5.56 / 2.74, as a 16-step shuffle chain 5.47 / 2.98.
#ifdef ERL_GP_V_ZDOT_EARLY
        constexpr bool kZdotEarly = true;
#else
        constexpr bool kZdotEarly = false;
#endif
#ifdef ERL_GP_V_BACKSOLVE_CHAIN
        constexpr bool kBackSolveDinv = false;
#else
        constexpr bool kBackSolveDinv = true;
#endif
#ifdef ERL_GP_ROWGP_LOOKAHEAD  // A/B: FactorizeMmaLa (pivot tile looked ahead, tiles dealt to the three warps that do not factorise).  Measured on
        constexpr bool kLookAhead = true;   // C4: train alone 2.74 -> 2.62 ms, but fused 4.80 -> 5.38 ms (= train + predict: the idle warps
#else                                       // of the default version are what lets the co-resident CTAs' predicts overlap) - off by default
        constexpr bool kLookAhead = false;
#endif
#ifdef ERL_GP_PIVOT_RSQRT_CHAIN  // A/B: the round-1 pivot chain (refined MUFU.RSQ per column)
        constexpr bool kPivotRcpChain = false;
#else
        constexpr bool kPivotRcpChain = true;
#endif
#ifndef ERL_GP_ROWGP_FFMA_TRAIN
        constexpr bool kMmaTrain = true;
#else
        constexpr bool kMmaTrain = false;  // A/B builds of the FFMA2 factorisation (Factorize)
#endif

        template<int NBLK>
        struct Layout {
            static constexpr int kNp = 16 * NBLK;

            __host__ __device__ static constexpr int
            Stride(const int cb) {  // floats between consecutive columns of column block cb (== 4 mod 16)
                return kNp - 16 * cb + 4;
            }

            __host__ __device__ static constexpr int
            Base(const int cb) {  // first float of column block cb: sum_{b < cb} 16 * Stride(b)
                return 16 * cb * (kNp + 4) - 128 * cb * (cb - 1);
            }

            static constexpr int kL = 0;
            static constexpr int kLFloats = 16 * NBLK * (kNp + 4) - 128 * NBLK * (NBLK - 1);
            static constexpr int kPts = kL + kLFloats;  // float4[kNp]: (x, y, z, alpha)
            static constexpr int kRs = kPts + 4 * kNp;  // 1 / L_jj
            static constexpr int kAl = kRs + kNp;       // y -> z -> alpha
            static constexpr int kVar = kAl + kNp;      // noise variances
            static constexpr int kMisc = kVar + kNp;    // int fail flag
            static constexpr int kDinvLd = 20;          // column stride of a 16 x 16 inverse block (== 4 mod 16, like the L columns)
            static constexpr int kDinv = kMisc + 4;     // inverses of the 16 x 16 diagonal blocks of L, column-major (tensor-path predict)
            static constexpr int kSoa = kDinv + NBLK * 16 * kDinvLd;  // x[kNp], y[kNp], z[kNp], alpha[kNp]: pairs of neighbouring points for the packed (f32x2) covariance code
            static constexpr int kEnd = kSoa + 4 * kNp;
            static constexpr size_t kBytes = static_cast<size_t>(kEnd) * sizeof(float);
        };

        __device__ __forceinline__ float2
        Fma2(const float2 a, const float s, const float2 c) {  // c + a * s, both halves (FFMA2 with broadcast operand)
            return __ffma2_rn(a, make_float2(s, s), c);
        }

        __device__ __forceinline__ float
        RsqrtRefined(const float d) {  // MUFU.RSQ + one Newton step (rsqrtf() wraps the MUFU in a denormal range test: four more
            float r;                   // instructions on the serial chain of every pivot; a pivot that small has failed anyway)
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
            return r * fmaf(-0.5f * d * r, r, 1.5f);
        }

        template<int XDIM>
        __device__ __forceinline__ float
        Dist2(const float4 a, const float (&b)[XDIM]) {
            float r2 = 0;
            const float av[3] = {a.x, a.y, a.z};
#pragma unroll
            for (int d = 0; d < XDIM; ++d) {
                const float diff = av[d] - b[d];
                r2 = fmaf(diff, diff, r2);
            }
            return r2;
        }

        // Branch-free form of the three erl_covariance kernels: with s = sqrt(r2),
        //   k = (1 + a s) exp(-(b s + c r2))      OU: a = 0, b = 1/l     Matern32: a = b = sqrt(3)/l     RBF: c = 1/(2 l^2)
        // One code path instead of three keeps the unrolled Gram / Ktest generators small (instruction cache).
        struct CovCoef {
            float a, b, c;

            // b and c are pre-multiplied by -log2(e): exp(-t) = ex2(-t log2 e) is one FFMA-fed MUFU.EX2 (max relative
            // error 2^-22 + |t| 2^-23, far inside the 1e-4 parity budget; expf() costs ~8 more instructions per entry
            // and the Gram / Ktest entries were 17 % of all instructions of the kernel)
            __device__ __forceinline__ explicit CovCoef(const Covariance<float> &cov) {
                constexpr float kLog2e = 1.4426950408889634f;
                a = cov.type == ERL_GP_KERNEL_MATERN32 ? cov.c0 : 0.f;
                b = -kLog2e * (cov.type == ERL_GP_KERNEL_MATERN32 ? cov.c0 : (cov.type == ERL_GP_KERNEL_OU ? 1.0f / cov.c0 : 0.f));
                c = -kLog2e * (cov.type == ERL_GP_KERNEL_RBF ? 1.0f / cov.c0 : 0.f);
            }

            __device__ __forceinline__ float
            operator()(const float r2) const {
                float sq, e;
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(r2));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(b, sq, c * r2)));
                return fmaf(a * sq, e, e);
            }
        };

        // Two covariance entries at a time on the packed FP32 pipe (add / mul / fma.f32x2): the entries of two neighbouring
        // training points against one query / row.  17 instructions per pair instead of 30 (the Gram / Ktest entries were 17 % of
        // the instructions of the fused kernel, all of it straight-line code the instruction fetch has to stream).
        template<int XDIM>
        __device__ __forceinline__ float2
        Dist2Pair(const float2 (&pc)[XDIM], const float (&negq)[XDIM]) {
            float2 d = __fadd2_rn(pc[0], make_float2(negq[0], negq[0]));
            float2 r2 = __fmul2_rn(d, d);
#pragma unroll
            for (int k = 1; k < XDIM; ++k) {
                d = __fadd2_rn(pc[k], make_float2(negq[k], negq[k]));
                r2 = __ffma2_rn(d, d, r2);
            }
            return r2;
        }

        __device__ __forceinline__ float2
        CovPair(const CovCoef &cov, const float2 r2) {
            float2 sq, e;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq.x) : "f"(r2.x));
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq.y) : "f"(r2.y));
            const float2 arg = __ffma2_rn(make_float2(cov.b, cov.b), sq, __fmul2_rn(make_float2(cov.c, cov.c), r2));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(arg.x));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(arg.y));
            return __ffma2_rn(__fmul2_rn(make_float2(cov.a, cov.a), sq), e, e);
        }

        // 16 x 16 pivot block of a panel, factorised by one warp with shuffles.  LDL^T-style elimination: the update of column
        // cc uses acc[c] / d (reciprocal) and the RAW column entries of the pivot rows, which can be shuffled before the
        // reciprocal is known; the Cholesky entries l[c] = acc[c] / sqrt(d) are formed off the dependency chain.
        // SRC: which lane owns pivot row c: 0 = pair layout, general (last panel); 1 = pair layout, even lanes (2 c); 2 = lane c
        template<int SRC>
        __device__ __forceinline__ void
        PivotBlock(float (&acc)[16], float &zacc, float (&l)[16], const int half, const int c0, const int lane, int &fail, float *__restrict__ rs, float *__restrict__ al) {
            if constexpr (SRC == 2 && kPivotRcpChain) {
                // Tensor-path factorisation, round 2: only MUFU.RCP (1 ulp) sits on the serial chain d_c -> 1 / d_c -> sc -> acc[c + 1] -> d_{c+1};
                // the square roots that turn the eliminated entries into L (l = acc / sqrt(d), z = zc / sqrt(d)) are taken after the
                // loop, one per lane in parallel (lane c keeps d_c) instead of sixteen refined MUFU.RSQ in sequence on every lane:
                // four dependent FP32 instructions less per pivot column on the critical path of every factorisation.
                float dmine = 1.0f, zraw = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float d = __shfl_sync(kFull, acc[c], c);
                    const float zc = __shfl_sync(kFull, zacc, c);
                    float t[16];
#pragma unroll
                    for (int cc = c + 1; cc < 16; ++cc) { t[cc] = __shfl_sync(kFull, acc[c], cc); }
                    if (!(d > 0.f) && fail == 0) { fail = c0 + c + 1; }
                    float invd;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(invd) : "f"(d));
                    const float sc = acc[c] * invd;
#pragma unroll
                    for (int cc = c + 1; cc < 16; ++cc) { acc[cc] = fmaf(-sc, t[cc], acc[cc]); }
                    zacc = fmaf(-sc, zc, zacc);
                    l[c] = acc[c];  // unscaled: acc[c] is final here
                    if ((lane & 15) == c) { dmine = d, zraw = zc; }
                }
                const float rsv = RsqrtRefined(dmine);  // lane c (and c + 16): 1 / sqrt(d_c)
#pragma unroll
                for (int c = 0; c < 16; ++c) { l[c] *= __shfl_sync(kFull, rsv, c); }
                if (lane < 16) {
                    rs[c0 + lane] = rsv;
                    al[c0 + lane] = zraw * rsv;
                }
                return;
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const int src = SRC == 2 ? c : SRC == 1 ? 2 * c : (c < half ? 2 * c : 2 * (c - half) + 1);
                const float d = __shfl_sync(kFull, acc[c], src);
                const float zc = __shfl_sync(kFull, zacc, src);
                float t[16];
#pragma unroll
                for (int cc = c + 1; cc < 16; ++cc) {
                    const int src2 = SRC == 2 ? cc : SRC == 1 ? 2 * cc : (cc < half ? 2 * cc : 2 * (cc - half) + 1);
                    t[cc] = __shfl_sync(kFull, acc[c], src2);
                }
                if (!(d > 0.f) && fail == 0) { fail = c0 + c + 1; }
                float invd, rsv;
                if (SRC == 2) {
                    // tensor-path factorisation: one MUFU.RSQ + one Newton step, 1 / d = rsv^2 (the products around it carry
                    // 2^-20 already; saves a MUFU and 3 dependent instructions per pivot on the serial chain of the kernel)
                    rsv = RsqrtRefined(d);
                    invd = rsv * rsv;
                } else {
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(invd) : "f"(d));
                    invd = invd * fmaf(-d, invd, 2.0f);  // one Newton step on MUFU.RCP
                    rsv = RsqrtRefined(d);
                }
                const float sc = acc[c] * invd;
#pragma unroll
                for (int cc = c + 1; cc < 16; ++cc) { acc[cc] = fmaf(-sc, t[cc], acc[cc]); }
                zacc = fmaf(-sc, zc, zacc);
                l[c] = acc[c] * rsv;
                if (lane == 0) {
                    rs[c0 + c] = rsv;
                    al[c0 + c] = zc * rsv;
                }
            }
        }

        // The tensor core reads only the upper 19 bits of a TF32 operand, so "hi" is the FP32 value itself (truncation is
        // implicit) and lo = x - trunc(x) is exact: one LOP3 + one FADD per split.  (cvt.rna.tf32.f32 is emulated on sm_100a
        // with IADD + FSETP + SEL + LOP3 - it made up 27 % of the instructions of the first version of this kernel.)
        __device__ __forceinline__ uint32_t
        Tf32Lo(const float x) {
            return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u));
        }

        // two values (a packed add.f32x2 for the subtraction was measured slower: the operands have to be moved into pairs)
        __device__ __forceinline__ void
        Tf32LoPair(const float x0, const float x1, uint32_t &lo0, uint32_t &lo1) {
            lo0 = Tf32Lo(x0);
            lo1 = Tf32Lo(x1);
        }

        __device__ __forceinline__ void
        MmaTf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
            asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
        }

        // d += A B with A = ahi + alo (pre-split) and the two B entries of this lane split here
        __device__ __forceinline__ void
        Mma3(float (&d)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], const float bf0, const float bf1) {
            const uint32_t bh0 = __float_as_uint(bf0), bh1 = __float_as_uint(bf1);
            MmaTf32(d, alo, bh0, bh1);
            MmaTf32(d, ahi, Tf32Lo(bf0), Tf32Lo(bf1));
            MmaTf32(d, ahi, bh0, bh1);
        }

        // accumulator tile (columns 2t, 2t+1 of rows g, g+8) -> A fragment (k-slots t, t+4), scaled by sgn, split in hi + lo
        __device__ __forceinline__ void
        AccToA(const float (&c)[4], const float sgn, uint32_t (&hi)[4], uint32_t (&lo)[4]) {
            const float a[4] = {sgn * c[0], sgn * c[2], sgn * c[1], sgn * c[3]};
#pragma unroll
            for (int k = 0; k < 4; ++k) { hi[k] = __float_as_uint(a[k]); }
            Tf32LoPair(a[0], a[1], lo[0], lo[1]);
            Tf32LoPair(a[2], a[3], lo[2], lo[3]);
        }

        // v (two 8-column accumulator tiles) = X Dinv^T for a 16 x 16 tile X held in the accumulator layout (x[0], x[1] = its two
        // 8-column halves) and the inverse Dinv of a 16 x 16 diagonal block of L (lower triangular: the (k-tile 1, n-tile 0)
        // product is zero).  dv = address of Dinv[g][2 t] in the column-major block (stride Layout::kDinvLd).  Three independent
        // accumulators, products issued term by term: consecutive HMMAs never wait for each other.
        template<int LD>
        __device__ __forceinline__ void
        MulDinvT(const float (&x0)[4], const float (&x1)[4], const float *__restrict__ dv, float (&v)[2][4]) {
            uint32_t xhi[2][4], xlo[2][4];
            AccToA(x0, 1.0f, xhi[0], xlo[0]);
            AccToA(x1, 1.0f, xhi[1], xlo[1]);
            float w[4] = {0.f, 0.f, 0.f, 0.f};  // n-tile 1, k-tile 1
#pragma unroll
            for (int c = 0; c < 4; ++c) { v[0][c] = v[1][c] = 0.f; }
            const float d00[2] = {dv[0], dv[LD]};                    // n-tile 0, k-tile 0
            const float d10[2] = {dv[8], dv[LD + 8]};                // n-tile 1, k-tile 0
            const float d11[2] = {dv[8 * LD + 8], dv[9 * LD + 8]};   // n-tile 1, k-tile 1
            uint32_t l00[2], l10[2], l11[2];
            Tf32LoPair(d00[0], d00[1], l00[0], l00[1]);
            Tf32LoPair(d10[0], d10[1], l10[0], l10[1]);
            Tf32LoPair(d11[0], d11[1], l11[0], l11[1]);
            MmaTf32(v[0], xlo[0], __float_as_uint(d00[0]), __float_as_uint(d00[1]));
            MmaTf32(v[1], xlo[0], __float_as_uint(d10[0]), __float_as_uint(d10[1]));
            MmaTf32(w, xlo[1], __float_as_uint(d11[0]), __float_as_uint(d11[1]));
            MmaTf32(v[0], xhi[0], l00[0], l00[1]);
            MmaTf32(v[1], xhi[0], l10[0], l10[1]);
            MmaTf32(w, xhi[1], l11[0], l11[1]);
            MmaTf32(v[0], xhi[0], __float_as_uint(d00[0]), __float_as_uint(d00[1]));
            MmaTf32(v[1], xhi[0], __float_as_uint(d10[0]), __float_as_uint(d10[1]));
            MmaTf32(w, xhi[1], __float_as_uint(d11[0]), __float_as_uint(d11[1]));
#pragma unroll
            for (int c = 0; c < 4; ++c) { v[1][c] += w[c]; }
        }

        // --------------------------------------------------------------------------------------
        // train: blocked left-looking Cholesky with the Gram panel generated on the fly
        // --------------------------------------------------------------------------------------
        template<int XDIM, int NBLK>
        __device__ __forceinline__ int
        Factorize(const CovCoef cov, float *__restrict__ smem, const int n, const int nblk) {
            using Lay = Layout<NBLK>;
            float *lp = smem + Lay::kL;
            const float4 *pts = reinterpret_cast<const float4 *>(smem + Lay::kPts);
            float *rs = smem + Lay::kRs;
            float *al = smem + Lay::kAl;
            const float *sv = smem + Lay::kVar;
            const int tid = threadIdx.x;
            const int warp = __shfl_sync(kFull, tid >> 5, 0);  // warp-uniform by construction: role branches need no reconvergence code
            const int lane = tid & 31;
            const int h = tid & 1;
            const int npr = nblk * 16;
            int fail = 0;

            for (int kb = 0; kb < nblk; ++kb) {
                const int c0 = 16 * kb;
                const int nact = npr - c0;
                const int half = nact >> 1;  // multiple of 8
                const bool warp_active = warp * 16 < half;
                if (warp_active) {
                    int q = tid >> 1;
                    const bool active = q < half;
                    if (!active) { q = half - 1; }  // duplicate the last pair: same code path, no stores
                    const int r_a = c0 + q;
                    const int r_b = r_a + half;
                    const int rme = h ? r_b : r_a;

                    // (a) this pair's update of rows r_a, r_b: thread h takes panel columns [8h, 8h + 8)
                    float2 s_a[4], s_b[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { s_a[k] = s_b[k] = make_float2(0.f, 0.f); }
                    float zsum = 0.f;
                    for (int jb = 0; jb < kb; ++jb) {
                        const int stride = Lay::Stride(jb);
                        const float *blk = lp + Lay::Base(jb) - 16 * jb;
                        const float *colp = blk + c0 + 8 * h;
                        const float *own_a = blk + r_a;
                        const float *own_b = blk + r_b;
                        const float4 *z4 = reinterpret_cast<const float4 *>(al + 16 * jb);
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 zq = z4[j4];
                            const float zv[4] = {zq.x, zq.y, zq.z, zq.w};
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int j = 4 * j4 + jj;
                                const float4 p0 = *reinterpret_cast<const float4 *>(colp + j * stride);
                                const float4 p1 = *reinterpret_cast<const float4 *>(colp + j * stride + 4);
                                const float la = own_a[j * stride];
                                const float lb = own_b[j * stride];
                                s_a[0] = Fma2(make_float2(p0.x, p0.y), la, s_a[0]);
                                s_a[1] = Fma2(make_float2(p0.z, p0.w), la, s_a[1]);
                                s_a[2] = Fma2(make_float2(p1.x, p1.y), la, s_a[2]);
                                s_a[3] = Fma2(make_float2(p1.z, p1.w), la, s_a[3]);
                                s_b[0] = Fma2(make_float2(p0.x, p0.y), lb, s_b[0]);
                                s_b[1] = Fma2(make_float2(p0.z, p0.w), lb, s_b[1]);
                                s_b[2] = Fma2(make_float2(p1.x, p1.y), lb, s_b[2]);
                                s_b[3] = Fma2(make_float2(p1.z, p1.w), lb, s_b[3]);
                                zsum = fmaf(h ? lb : la, zv[jj], zsum);
                            }
                        }
                    }
                    // swap halves inside the pair: I keep row rme, the partner the other row
                    float acc[16];
                    {
                        float mine[8], recv[8];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 keep = h ? s_b[k] : s_a[k];
                            const float2 send = h ? s_a[k] : s_b[k];
                            mine[2 * k] = keep.x;
                            mine[2 * k + 1] = keep.y;
                            recv[2 * k] = __shfl_xor_sync(kFull, send.x, 1);
                            recv[2 * k + 1] = __shfl_xor_sync(kFull, send.y, 1);
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            acc[k] = h ? recv[k] : mine[k];      // columns 0..7 were computed by h = 0
                            acc[8 + k] = h ? mine[k] : recv[k];  // columns 8..15 by h = 1
                        }
                    }
                    // (b) Gram panel entries of my row, fused: acc = K(rme, c0 + c) - sum
                    {
                        const float4 pme = pts[rme];
                        float xme[XDIM];
                        xme[0] = pme.x;
                        if (XDIM > 1) { xme[XDIM > 1 ? 1 : 0] = pme.y; }
                        if (XDIM > 2) { xme[XDIM > 2 ? 2 : 0] = pme.z; }
                        const float diag = 1.0f + sv[rme];
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const int col = c0 + c;
                            float kv = cov(Dist2<XDIM>(pts[col], xme));
                            if (rme >= n || col >= n) { kv = 0.f; }
                            if (rme == col) { kv = rme < n ? diag : 1.0f; }
                            acc[c] = kv - acc[c];
                        }
                    }
                    float zacc = al[rme] - zsum;  // al[rme] still holds y (rme >= c0)
                    float l[16];

                    if (warp == 0) {
                        // (c) pivot block: rows c0 + c live in lanes src(c); every lane's row follows along.  While the panel has
                        // at least 32 rows the pivot rows sit in the even lanes (static shuffle sources); only the last panel
                        // (16 rows: pairs q < 8 hold rows q and q + 8) needs the general mapping.
                        if (half >= 16) {
                            PivotBlock<1>(acc, zacc, l, half, c0, lane, fail, rs, al);
                        } else {
                            PivotBlock<0>(acc, zacc, l, half, c0, lane, fail, rs, al);
                        }
                    }
                    if (warp == 0 && active) {
                        float *dst = lp + Lay::Base(kb) + (rme - c0);
                        const int stride = Lay::Stride(kb);
#pragma unroll
                        for (int c = 0; c < 16; ++c) { dst[c * stride] = (c > rme - c0) ? 0.f : l[c]; }
                    }
                    __syncthreads();  // #1: pivot block, rs, z of this panel are published
                    if (half > 16) {
                        if (warp != 0) {
                            // (d) in-thread elimination of my row against the pivot block
                            const float *dblk = lp + Lay::Base(kb);
                            const int stride = Lay::Stride(kb);
                            float rsb[16], zb[16];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float4 r4 = *reinterpret_cast<const float4 *>(rs + c0 + 4 * k);
                                const float4 z4 = *reinterpret_cast<const float4 *>(al + c0 + 4 * k);
                                rsb[4 * k] = r4.x, rsb[4 * k + 1] = r4.y, rsb[4 * k + 2] = r4.z, rsb[4 * k + 3] = r4.w;
                                zb[4 * k] = z4.x, zb[4 * k + 1] = z4.y, zb[4 * k + 2] = z4.z, zb[4 * k + 3] = z4.w;
                            }
#pragma unroll
                            for (int pcol = 0; pcol < 16; ++pcol) {
                                const float lv = acc[pcol] * rsb[pcol];
                                l[pcol] = lv;
                                zacc = fmaf(-lv, zb[pcol], zacc);
#pragma unroll
                                for (int k4 = (pcol + 1) / 4; k4 < 4; ++k4) {
                                    const float4 dv = *reinterpret_cast<const float4 *>(dblk + pcol * stride + 4 * k4);
                                    // rows <= pcol of the pivot column are zero (or the dead diagonal): harmless
                                    acc[4 * k4] = fmaf(-lv, dv.x, acc[4 * k4]);
                                    acc[4 * k4 + 1] = fmaf(-lv, dv.y, acc[4 * k4 + 1]);
                                    acc[4 * k4 + 2] = fmaf(-lv, dv.z, acc[4 * k4 + 2]);
                                    acc[4 * k4 + 3] = fmaf(-lv, dv.w, acc[4 * k4 + 3]);
                                }
                            }
                            if (active) {
                                float *dst = lp + Lay::Base(kb) + (rme - c0);
#pragma unroll
                                for (int c = 0; c < 16; ++c) { dst[c * stride] = l[c]; }
                            }
                        }
                        __syncthreads();  // #2: the whole panel is published
                    }
                    (void) zacc;
                } else {
                    __syncthreads();  // #1
                    if (half > 16) { __syncthreads(); }  // #2
                }
            }
            return fail;
        }

        // --------------------------------------------------------------------------------------
        // train on the tensor path: the same left-looking blocked Cholesky, but every product is a 3xTF32 mma.sync.
        //
        // Panel kb (16 columns, rows c0 = 16 kb .. npr) is cut in 16-row tiles; tile i goes to warp i % 4.  Per panel:
        //   A. update: P_i = K[tile i, panel] - L[tile i, 0:c0] L[panel rows, 0:c0]^T.  A fragments are rows of L, B fragments the
        //      pivot rows of L, both read from the packed column-major L with the k-slot permutation (slot t <-> column 2 t,
        //      slot t + 4 <-> column 2 t + 1) that makes the LDS.32 bank-conflict free; the Gram entries are generated in the
        //      accumulator layout (never stored).
        //   B. warp 0 owns tile 0 = the 16 x 16 pivot block: it goes through shared memory once to get "lane r owns row r" and
        //      is factorised with shuffles (PivotBlock; z = L^-1 y rides along); lanes 16 .. 31 run the same instructions on the
        //      unit vectors and end up with the columns of its inverse Dinv.
        //   C. the other tiles become L_i = P_i Dinv^T: the accumulators ARE the A operand (same trick as in the predict), so the
        //      in-thread elimination of the FFMA version (136 dependent FMAs + broadcast loads per row) is 9 HMMAs per tile.
        // Two barriers per panel as before, but ~2.5x fewer instructions between them and no idle "finished rows" threads in
        // phase A (tiles are dealt to all four warps).  Cycle counters (-DERL_GP_ROWGP_TIMING, C4, 4 CTAs / SM): a factorisation
        // takes ~62 k cycles, all of it on warp 0's path: 27 k in phase A (two tiles per panel while mt > 4, one of them the pivot
        // tile) and 32 k in phase B (8 pivot blocks, ~4 k each: 16 dependent shuffle + MUFU.RSQ + FMA steps at the latency the
        // loaded SM gives them); the other warps are busy 14 - 19 k cycles.  Tried and measured slower: warp 0 on the pivot tile
        // only (5.95 vs 5.58 ms fused), and splitting the pivot tile's update over the four warps by column block with the
        // shares summed through shared memory behind a third barrier (5.43 vs 4.97 ms); keeping the 16 pivot diagonals up to date
        // in every lane so that the MUFU.RSQ chain does not run through the shuffles (+240 independent instructions per pivot
        // block: train alone 2.65 -> 2.71 ms, fused 4.97 -> 5.30 ms).  Every change that ADDED instructions made the fused kernel
        // slower and every one that removed some made it faster: with 4 warps per scheduler the fused kernel behaves as
        // issue-throughput bound, not as bound by the latency of warp 0's chain.  Accuracy: tests/test_gpu_batch.py; a numpy emulation of the split
        // (tools/emulate_3xtf32.py) gives mean / variance errors of 4e-6 / 1e-6 against 1e-6 / 6e-7 for plain FP32.
        // --------------------------------------------------------------------------------------
        template<int XDIM, int NBLK>
        __device__ __forceinline__ int
        FactorizeMma(const CovCoef cov, float *__restrict__ smem, const int n, const int nblk) {
            using Lay = Layout<NBLK>;
            constexpr int kWarps = ThreadsFor<NBLK>::value / 32;
            constexpr int kSlots = (NBLK + kWarps - 1) / kWarps;  // tiles per warp and panel
            float *lp = smem + Lay::kL;
            const float4 *pts = reinterpret_cast<const float4 *>(smem + Lay::kPts);
            float *rs = smem + Lay::kRs;
            float *al = smem + Lay::kAl;
            const float *sv = smem + Lay::kVar;
            float *dinv = smem + Lay::kDinv;
            const int tid = threadIdx.x;
            const int warp = __shfl_sync(kFull, tid >> 5, 0);
            const int lane = tid & 31;
            const int g = lane >> 2, t = lane & 3;
            int fail = 0;

            // -DERL_GP_ROWGP_TIMING: cycle counters per phase and warp, printed by two CTAs (kernel experiments only)
#ifdef ERL_GP_ROWGP_TIMING
            long long tm_a = 0, tm_b = 0, tm_w1 = 0, tm_c = 0, tm_w2 = 0, tm_t = clock64();
#define ERL_GP_TICK(acc_) { const long long now_ = clock64(); acc_ += now_ - tm_t; tm_t = now_; }
#else
#define ERL_GP_TICK(acc_)
#endif
            for (int kb = 0; kb < nblk; ++kb) {
                const int c0 = 16 * kb;
                const int mt = nblk - kb;  // tiles of this panel
                const int stride_k = Lay::kNp - 16 * kb + 4;
                float *panel = lp + (16 * kb * (Lay::kNp + 4) - 128 * kb * (kb - 1));  // element (row c0, column c0)

                // z: sum_{j < c0} L[r][j] z_j for the pivot rows r (warp 0; two lanes per row: even / odd column blocks)
                auto pivot_row_dot_z = [&]() {
                    const int r = lane & 15, hh = lane >> 4;
                    float zp[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int jb = hh; jb < kb; jb += 2) {
                        const int stride = Lay::kNp - 16 * jb + 4;
                        const float *rowp = lp + (16 * jb * (Lay::kNp + 4) - 128 * jb * (jb - 1)) + (c0 + r - 16 * jb);
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 z4 = *reinterpret_cast<const float4 *>(al + 16 * jb + 4 * j4);
                            zp[0] = fmaf(rowp[(4 * j4) * stride], z4.x, zp[0]);
                            zp[1] = fmaf(rowp[(4 * j4 + 1) * stride], z4.y, zp[1]);
                            zp[2] = fmaf(rowp[(4 * j4 + 2) * stride], z4.z, zp[2]);
                            zp[3] = fmaf(rowp[(4 * j4 + 3) * stride], z4.w, zp[3]);
                        }
                    }
                    const float zsum = (zp[0] + zp[1]) + (zp[2] + zp[3]);
                    return zsum + __shfl_xor_sync(kFull, zsum, 16);
                };
                float zs = 0.f;
                if (kZdotEarly && warp == 0) { zs = pivot_row_dot_z(); }  // before the update: the FMA chain overlaps with the HMMAs

                // ---- A: update of my tiles -------------------------------------------------------------------------
                float acc[kSlots][2][4];
#pragma unroll
                for (int sl = 0; sl < kSlots; ++sl) {
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) { acc[sl][nt][0] = acc[sl][nt][1] = acc[sl][nt][2] = acc[sl][nt][3] = 0.f; }
                }
                if (warp < mt) {
                    for (int jb = 0; jb < kb; ++jb) {
                        const int stride = Lay::kNp - 16 * jb + 4;
                        // element (row 16 jb + g, column 16 jb + 2 t) of column block jb; rows are added relative to 16 jb
                        const float *cb = lp + (16 * jb * (Lay::kNp + 4) - 128 * jb * (jb - 1)) + 2 * t * stride + g;
                        const float *brow = cb + (c0 - 16 * jb);
#pragma unroll
                        for (int kt = 0; kt < 2; ++kt) {
                            float b[2][2];
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt) {
                                b[nt][0] = brow[8 * kt * stride + 8 * nt];           // L[c0 + 8 nt + g][16 jb + 8 kt + 2 t]
                                b[nt][1] = brow[(8 * kt + 1) * stride + 8 * nt];     // ... + 1
                            }
                            uint32_t bl[2][2];
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt) { Tf32LoPair(b[nt][0], b[nt][1], bl[nt][0], bl[nt][1]); }
#pragma unroll
                            for (int sl = 0; sl < kSlots; ++sl) {
                                const int ti = warp + kWarps * sl;
                                if (ti < mt) {
                                    const float *arow = brow + 16 * ti + 8 * kt * stride;
                                    // slots (row g, k t), (row g + 8, k t), (row g, k t + 4), (row g + 8, k t + 4)
                                    const float a[4] = {arow[0], arow[8], arow[stride], arow[stride + 8]};
                                    uint32_t ahi[4], alo[4];
#pragma unroll
                                    for (int k = 0; k < 4; ++k) { ahi[k] = __float_as_uint(a[k]); }
                                    Tf32LoPair(a[0], a[1], alo[0], alo[1]);
                                    Tf32LoPair(a[2], a[3], alo[2], alo[3]);
#pragma unroll
                                    for (int nt = 0; nt < 2; ++nt) { MmaTf32(acc[sl][nt], alo, __float_as_uint(b[nt][0]), __float_as_uint(b[nt][1])); }
#pragma unroll
                                    for (int nt = 0; nt < 2; ++nt) { MmaTf32(acc[sl][nt], ahi, bl[nt][0], bl[nt][1]); }
#pragma unroll
                                    for (int nt = 0; nt < 2; ++nt) { MmaTf32(acc[sl][nt], ahi, __float_as_uint(b[nt][0]), __float_as_uint(b[nt][1])); }
                                }
                            }
                        }
                    }
                    // P = Gram tile - update (Gram entries in the accumulator layout, two neighbouring columns per packed
                    // operation, fused noise diagonal, identity padding)
                    const float2 *soa = reinterpret_cast<const float2 *>(smem + Lay::kSoa);
                    float2 pcx[2][XDIM];
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                        for (int d = 0; d < XDIM; ++d) { pcx[nt][d] = soa[d * (Lay::kNp / 2) + (c0 + 8 * nt) / 2 + t]; }
                    }
                    const bool ragged = c0 + 16 > n;  // padding columns in this panel (warp-uniform)
#pragma unroll
                    for (int sl = 0; sl < kSlots; ++sl) {
                        const int ti = warp + kWarps * sl;
                        if (ti < mt) {
#pragma unroll
                            for (int hr = 0; hr < 2; ++hr) {
                                const int row = c0 + 16 * ti + g + 8 * hr;
                                const float4 pr = pts[row];
                                float negr[XDIM];
                                negr[0] = -pr.x;
                                if (XDIM > 1) { negr[XDIM > 1 ? 1 : 0] = -pr.y; }
                                if (XDIM > 2) { negr[XDIM > 2 ? 2 : 0] = -pr.z; }
                                const float diag = row < n ? 1.0f + sv[row] : 1.0f;
#pragma unroll
                                for (int nt = 0; nt < 2; ++nt) {
                                    const int col = c0 + 8 * nt + 2 * t;
                                    float2 kv = CovPair(cov, Dist2Pair<XDIM>(pcx[nt], negr));
                                    if (row >= n) { kv = make_float2(0.f, 0.f); }
                                    if (ragged) {
                                        if (col >= n) { kv.x = 0.f; }
                                        if (col + 1 >= n) { kv.y = 0.f; }
                                    }
                                    if (row == col) { kv.x = diag; }
                                    if (row == col + 1) { kv.y = diag; }
                                    acc[sl][nt][2 * hr] = kv.x - acc[sl][nt][2 * hr];
                                    acc[sl][nt][2 * hr + 1] = kv.y - acc[sl][nt][2 * hr + 1];
                                }
                            }
                        }
                    }
                }

                ERL_GP_TICK(tm_a)
                // ---- B: pivot block (warp 0 holds tile 0 in slot 0) --------------------------------------------------
                if (warp == 0) {
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            panel[(8 * nt + 2 * t + e) * stride_k + g] = acc[0][nt][e];
                            panel[(8 * nt + 2 * t + e) * stride_k + g + 8] = acc[0][nt][2 + e];
                        }
                    }
                    const int r = lane & 15;
                    if (!kZdotEarly) { zs = pivot_row_dot_z(); }
                    float zacc = al[c0 + r] - zs;  // al[c0 + r] still holds y
                    __syncwarp();
                    // Lanes 0 .. 15 own the rows of the pivot tile.  Lanes 16 + j start from the unit vector e_j instead and run
                    // the very same elimination: with t[i] = L[i][c] L[c][c] and sc = acc[c] / d the update acc[i] -= sc t[i] is the
                    // forward substitution of L x = e_j, and the scaled entries l[c] = acc[c] / L[c][c] that lanes 0 .. 15 read as row
                    // r of L are, in lane 16 + j, column j of Dinv = L^-1.  The inverse costs no instruction at all.
                    float prow[16], l[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const float pv = panel[c * stride_k + r];
                        prow[c] = lane < 16 ? pv : (c == r ? 1.0f : 0.f);
                    }
                    PivotBlock<2>(prow, zacc, l, 0, c0, lane, fail, rs, al);
                    __syncwarp();  // everybody has read the raw tile
                    if (lane < 16) {
#pragma unroll
                        for (int c = 0; c < 16; ++c) { panel[c * stride_k + r] = c > r ? 0.f : l[c]; }
                    } else {
                        float *dst = dinv + kb * 16 * Lay::kDinvLd + r * Lay::kDinvLd;  // column r of Dinv (zero above the diagonal)
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) { *reinterpret_cast<float4 *>(dst + 4 * k4) = make_float4(l[4 * k4], l[4 * k4 + 1], l[4 * k4 + 2], l[4 * k4 + 3]); }
                    }
                }
                ERL_GP_TICK(tm_b)
                __syncthreads();  // #1: pivot block, Dinv, rs, z of this panel are published
                ERL_GP_TICK(tm_w1)
                if (mt > 1) {
                    // ---- C: L_i = P_i Dinv^T for the tiles below the pivot block ----------------------------------------
                    const float *dv = dinv + kb * 16 * Lay::kDinvLd + 2 * t * Lay::kDinvLd + g;
#pragma unroll
                    for (int sl = 0; sl < kSlots; ++sl) {
                        const int ti = warp + kWarps * sl;
                        if (ti > 0 && ti < mt) {
                            float v[2][4];
                            MulDinvT<Lay::kDinvLd>(acc[sl][0], acc[sl][1], dv, v);
                            float *dst = panel + 16 * ti + g;
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    dst[(8 * nt + 2 * t + e) * stride_k] = v[nt][e];
                                    dst[(8 * nt + 2 * t + e) * stride_k + 8] = v[nt][2 + e];
                                }
                            }
                        }
                    }
                    ERL_GP_TICK(tm_c)
                    __syncthreads();  // #2: the whole panel is published
                    ERL_GP_TICK(tm_w2)
                }
            }
#ifdef ERL_GP_ROWGP_TIMING
            if ((blockIdx.x == 20000 || blockIdx.x == 31111) && lane == 0) {
                printf("cta %d warp %d: A %lld  B %lld  wait1 %lld  C %lld  wait2 %lld\n", blockIdx.x, warp, tm_a, tm_b, tm_w1, tm_c, tm_w2);
            }
#endif
            return fail;
        }

        // --------------------------------------------------------------------------------------
        // FactorizeMma with the pivot tile LOOKED AHEAD (round 2; 128-thread instances, n <= 128).
        //
        // In FactorizeMma the serial chain of a panel is  update of the pivot tile (all earlier blocks, warp 0) -> pivot tile (warp 0, the
        // other warps wait) -> L_i = P_i Dinv^T: 62 k cycles per factorisation, all of it on warp 0 (27 k of updates + 32 k of pivot
        // tiles).  Here the pivot tile of panel kb + 1 is finished by the warp that will factorise it while panel kb is still in
        // flight, and the other tiles are dealt to the three warps that are NOT factorising:
        //   * panel kb is factorised by warp p = kb % 4; the tiles below (kb + 1 ...) are dealt to the helper warps p + 1, p + 2, p + 3,
        //     tile kb + 1 always to p + 1 = the warp that factorises panel kb + 1;
        //   * while p runs the pivot tile (shuffles, one warp), the helpers run the left-looking update of their tiles; the owner of
        //     tile kb + 1 also accumulates that tile's product with ITSELF over the earlier blocks (its A fragments are, re-ordered,
        //     the B fragments: no extra loads);
        //   * after the barrier (Dinv of panel kb published) the helpers form L_i = P_i Dinv^T; the owner of tile kb + 1 adds the
        //     product of the rows it has just computed with themselves (both operands are its own accumulator registers), subtracts
        //     from the Gram tile and holds P_{kb+1,kb+1}: after the second barrier it factorises it at once.
        // Serial chain per panel: pivot tile || update of <= 3 tiles, then 9 + 12 HMMA products.  Same instruction count as FactorizeMma.
        // --------------------------------------------------------------------------------------
        template<int XDIM, int NBLK>
        __device__ __forceinline__ int
        FactorizeMmaLa(const CovCoef cov, float *__restrict__ smem, const int n, const int nblk) {
            using Lay = Layout<NBLK>;
            static_assert(ThreadsFor<NBLK>::value == 128 && NBLK <= 8, "four warps, at most 7 tiles below a pivot tile");
            float *lp = smem + Lay::kL;
            const float4 *pts = reinterpret_cast<const float4 *>(smem + Lay::kPts);
            float *rs = smem + Lay::kRs;
            float *al = smem + Lay::kAl;
            const float *sv = smem + Lay::kVar;
            float *dinv = smem + Lay::kDinv;
            const float2 *soa = reinterpret_cast<const float2 *>(smem + Lay::kSoa);
            int *s_fail4 = reinterpret_cast<int *>(smem + Lay::kMisc);  // one slot per warp
            const int tid = threadIdx.x;
            const int warp = __shfl_sync(kFull, tid >> 5, 0);
            const int lane = tid & 31;
            const int g = lane >> 2, t = lane & 3;
            int fail = 0;
            float pt[2][4];  // the pivot tile this warp factorises next (accumulator layout)

            // Gram tile K[row0 + (g, g + 8)][col0 + 8 nt + 2 t + (0, 1)] in the accumulator layout (fused noise diagonal, identity padding)
            auto gram = [&](const int row0, const int col0, float (&kt)[2][4]) {
                float2 pcx[2][XDIM];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                    for (int d = 0; d < XDIM; ++d) { pcx[nt][d] = soa[d * (Lay::kNp / 2) + (col0 + 8 * nt) / 2 + t]; }
                }
                const bool ragged = col0 + 16 > n;
#pragma unroll
                for (int hr = 0; hr < 2; ++hr) {
                    const int row = row0 + g + 8 * hr;
                    const float4 pr = pts[row];
                    float negr[XDIM];
                    negr[0] = -pr.x;
                    if (XDIM > 1) { negr[XDIM > 1 ? 1 : 0] = -pr.y; }
                    if (XDIM > 2) { negr[XDIM > 2 ? 2 : 0] = -pr.z; }
                    const float diag = row < n ? 1.0f + sv[row] : 1.0f;
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        const int col = col0 + 8 * nt + 2 * t;
                        float2 kv = CovPair(cov, Dist2Pair<XDIM>(pcx[nt], negr));
                        if (row >= n) { kv = make_float2(0.f, 0.f); }
                        if (ragged) {
                            if (col >= n) { kv.x = 0.f; }
                            if (col + 1 >= n) { kv.y = 0.f; }
                        }
                        if (row == col) { kv.x = diag; }
                        if (row == col + 1) { kv.y = diag; }
                        kt[nt][2 * hr] = kv.x;
                        kt[nt][2 * hr + 1] = kv.y;
                    }
                }
            };

            if (warp == 0) { gram(0, 0, pt); }
            for (int kb = 0; kb < nblk; ++kb) {
                const int c0 = 16 * kb;
                const int p = kb & 3;         // the warp that factorises this panel's pivot tile
                const int ntl = nblk - kb - 1;  // tiles below the pivot tile
                const int stride_k = Lay::kNp - 16 * kb + 4;
                float *panel = lp + (16 * kb * (Lay::kNp + 4) - 128 * kb * (kb - 1));  // element (row c0, column c0)
                const int h = (warp - p - 1) & 3;  // helper index 0 .. 2 (3: the pivot warp)
                // tile u (rows c0 + 16 (1 + u)) of helper h in slot s: h = 0: {0, 5}, h = 1: {1, 3, 6}, h = 2: {2, 4}
                const int u1 = h == 0 ? 5 : (h == 1 ? 3 : 4);
                const int tile_u[3] = {h, u1, h == 1 ? 6 : 99};
                float acc[3][2][4], pn[2][4];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) { acc[0][nt][e] = acc[1][nt][e] = acc[2][nt][e] = pn[nt][e] = 0.f; }
                }
                if (warp != p) {
                    // ---- helpers: left-looking update of my tiles (+ the self product of tile kb + 1 on helper 0) ----
                    for (int jb = 0; jb < kb; ++jb) {
                        const int stride = Lay::kNp - 16 * jb + 4;
                        const float *cb = lp + (16 * jb * (Lay::kNp + 4) - 128 * jb * (jb - 1)) + 2 * t * stride + g;
                        const float *brow = cb + (c0 - 16 * jb);
#pragma unroll
                        for (int kt = 0; kt < 2; ++kt) {
                            float b[2][2];
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt) {
                                b[nt][0] = brow[8 * kt * stride + 8 * nt];
                                b[nt][1] = brow[(8 * kt + 1) * stride + 8 * nt];
                            }
                            uint32_t bl[2][2];
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt) { Tf32LoPair(b[nt][0], b[nt][1], bl[nt][0], bl[nt][1]); }
#pragma unroll
                            for (int sl = 0; sl < 3; ++sl) {
                                if (tile_u[sl] < ntl) {
                                    const float *arow = brow + 16 * (1 + tile_u[sl]) + 8 * kt * stride;
                                    const float a[4] = {arow[0], arow[8], arow[stride], arow[stride + 8]};
                                    uint32_t ahi[4], alo[4];
#pragma unroll
                                    for (int k = 0; k < 4; ++k) { ahi[k] = __float_as_uint(a[k]); }
                                    Tf32LoPair(a[0], a[1], alo[0], alo[1]);
                                    Tf32LoPair(a[2], a[3], alo[2], alo[3]);
#pragma unroll
                                    for (int nt = 0; nt < 2; ++nt) { MmaTf32(acc[sl][nt], alo, __float_as_uint(b[nt][0]), __float_as_uint(b[nt][1])); }
#pragma unroll
                                    for (int nt = 0; nt < 2; ++nt) { MmaTf32(acc[sl][nt], ahi, bl[nt][0], bl[nt][1]); }
#pragma unroll
                                    for (int nt = 0; nt < 2; ++nt) { MmaTf32(acc[sl][nt], ahi, __float_as_uint(b[nt][0]), __float_as_uint(b[nt][1])); }
                                    if (sl == 0 && h == 0) {
                                        // tile kb + 1 times itself: B[k][n] = L[row n][col k] are my own A entries (rows g -> n-tile 0, g + 8 -> n-tile 1)
                                        MmaTf32(pn[0], alo, ahi[0], ahi[2]);
                                        MmaTf32(pn[1], alo, ahi[1], ahi[3]);
                                        MmaTf32(pn[0], ahi, alo[0], alo[2]);
                                        MmaTf32(pn[1], ahi, alo[1], alo[3]);
                                        MmaTf32(pn[0], ahi, ahi[0], ahi[2]);
                                        MmaTf32(pn[1], ahi, ahi[1], ahi[3]);
                                    }
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int sl = 0; sl < 3; ++sl) {
                        if (tile_u[sl] < ntl) {
                            float kt[2][4];
                            gram(c0 + 16 * (1 + tile_u[sl]), c0, kt);
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) { acc[sl][nt][e] = kt[nt][e] - acc[sl][nt][e]; }
                            }
                        }
                    }
                } else {
                    // ---- the pivot warp: P_kb,kb is in pt ----
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            panel[(8 * nt + 2 * t + e) * stride_k + g] = pt[nt][e];
                            panel[(8 * nt + 2 * t + e) * stride_k + g + 8] = pt[nt][2 + e];
                        }
                    }
                    const int r = lane & 15, hh = lane >> 4;
                    float zp[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int jb = hh; jb < kb; jb += 2) {
                        const int stride = Lay::kNp - 16 * jb + 4;
                        const float *rowp = lp + (16 * jb * (Lay::kNp + 4) - 128 * jb * (jb - 1)) + (c0 + r - 16 * jb);
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 z4 = *reinterpret_cast<const float4 *>(al + 16 * jb + 4 * j4);
                            zp[0] = fmaf(rowp[(4 * j4) * stride], z4.x, zp[0]);
                            zp[1] = fmaf(rowp[(4 * j4 + 1) * stride], z4.y, zp[1]);
                            zp[2] = fmaf(rowp[(4 * j4 + 2) * stride], z4.z, zp[2]);
                            zp[3] = fmaf(rowp[(4 * j4 + 3) * stride], z4.w, zp[3]);
                        }
                    }
                    float zs = (zp[0] + zp[1]) + (zp[2] + zp[3]);
                    zs += __shfl_xor_sync(kFull, zs, 16);
                    float zacc = al[c0 + r] - zs;  // al[c0 + r] still holds y
                    __syncwarp();
                    float prow[16], l[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const float pv = panel[c * stride_k + r];
                        prow[c] = lane < 16 ? pv : (c == r ? 1.0f : 0.f);
                    }
                    PivotBlock<2>(prow, zacc, l, 0, c0, lane, fail, rs, al);
                    __syncwarp();  // everybody has read the raw tile
                    if (lane < 16) {
#pragma unroll
                        for (int c = 0; c < 16; ++c) { panel[c * stride_k + r] = c > r ? 0.f : l[c]; }
                    } else {
                        float *dst = dinv + kb * 16 * Lay::kDinvLd + r * Lay::kDinvLd;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) { *reinterpret_cast<float4 *>(dst + 4 * k4) = make_float4(l[4 * k4], l[4 * k4 + 1], l[4 * k4 + 2], l[4 * k4 + 3]); }
                    }
                }
                __syncthreads();  // #1: pivot tile, Dinv, rs, z of this panel are published
                if (ntl > 0) {
                    if (warp != p) {
                        const float *dv = dinv + kb * 16 * Lay::kDinvLd + 2 * t * Lay::kDinvLd + g;
#pragma unroll
                        for (int sl = 0; sl < 3; ++sl) {
                            if (tile_u[sl] < ntl) {
                                float v[2][4];
                                MulDinvT<Lay::kDinvLd>(acc[sl][0], acc[sl][1], dv, v);
                                float *dst = panel + 16 * (1 + tile_u[sl]) + g;
#pragma unroll
                                for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                                    for (int e = 0; e < 2; ++e) {
                                        dst[(8 * nt + 2 * t + e) * stride_k] = v[nt][e];
                                        dst[(8 * nt + 2 * t + e) * stride_k + 8] = v[nt][2 + e];
                                    }
                                }
                                if (sl == 0 && h == 0) {
                                    // look-ahead: P_{kb+1,kb+1} = K - (earlier blocks) - L_{kb+1,kb} L_{kb+1,kb}^T, both operands from v
#pragma unroll
                                    for (int kt = 0; kt < 2; ++kt) {
                                        uint32_t ahi[4], alo[4];
                                        AccToA(v[kt], 1.0f, ahi, alo);
                                        uint32_t bl0[2], bl1[2];
                                        Tf32LoPair(v[kt][0], v[kt][1], bl0[0], bl0[1]);
                                        Tf32LoPair(v[kt][2], v[kt][3], bl1[0], bl1[1]);
                                        MmaTf32(pn[0], alo, __float_as_uint(v[kt][0]), __float_as_uint(v[kt][1]));
                                        MmaTf32(pn[1], alo, __float_as_uint(v[kt][2]), __float_as_uint(v[kt][3]));
                                        MmaTf32(pn[0], ahi, bl0[0], bl0[1]);
                                        MmaTf32(pn[1], ahi, bl1[0], bl1[1]);
                                        MmaTf32(pn[0], ahi, __float_as_uint(v[kt][0]), __float_as_uint(v[kt][1]));
                                        MmaTf32(pn[1], ahi, __float_as_uint(v[kt][2]), __float_as_uint(v[kt][3]));
                                    }
                                    float kt2[2][4];
                                    gram(c0 + 16, c0 + 16, kt2);
#pragma unroll
                                    for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                                        for (int e = 0; e < 4; ++e) { pt[nt][e] = kt2[nt][e] - pn[nt][e]; }
                                    }
                                }
                            }
                        }
                    }
                    __syncthreads();  // #2: the whole panel is published
                }
            }
            // every warp factorised some of the pivot tiles: the first failing column over the four of them
            if (lane == 0) { s_fail4[warp] = fail; }
            __syncthreads();
            int first = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int f = s_fail4[w];
                if (f != 0 && (first == 0 || f < first)) { first = f; }
            }
            __syncthreads();  // (slot 0 is rewritten by the caller)
            return first;
        }

        // alpha = L^-T z (al holds z on entry, alpha on exit); thread = column, blocked from the bottom
        // USE_DINV: the inverses of the diagonal blocks are in shared memory (FactorizeMma)
        template<int NBLK, bool USE_DINV>
        __device__ __forceinline__ void
        BackSolve(float *__restrict__ smem, const int nblk) {
            using Lay = Layout<NBLK>;
            constexpr int kThr = ThreadsFor<NBLK>::value;
            const float *lp = smem + Lay::kL;
            const float *rs = smem + Lay::kRs;
            float *al = smem + Lay::kAl;
            const int tid = threadIdx.x;
            const int warp = __shfl_sync(kFull, tid >> 5, 0);  // warp-uniform by construction: role branches need no reconvergence code
            const int lane = tid & 31;
            constexpr int kColSlots = (Lay::kNp + kThr - 1) / kThr;  // columns per thread: column = tid + kThr * slot
            float s[kColSlots];
#pragma unroll
            for (int sl = 0; sl < kColSlots; ++sl) { s[sl] = 0.f; }
            for (int kb = nblk - 1; kb >= 0; --kb) {
                const int c0 = 16 * kb;
                if (warp == ((c0 & (kThr - 1)) >> 5)) {
                    const int lb = c0 & 31;
                    const bool mine = lane >= lb && lane < lb + 16;
                    const int jj = mine ? lane - lb : 0;
                    float sj = s[0];
#pragma unroll
                    for (int sl = 1; sl < kColSlots; ++sl) {
                        if (c0 >= kThr * sl) { sj = s[sl]; }
                    }
                    float amine;
                    if constexpr (USE_DINV) {
                        // alpha_blk = Dinv^T (z_blk - s_blk): 16 independent shuffles + FMAs instead of a 16-step substitution chain
                        const float vj = al[c0 + jj] - sj;
                        const float *dcol = smem + Lay::kDinv + kb * 16 * Lay::kDinvLd + jj * Lay::kDinvLd;  // column jj of Dinv: Dinv[r][jj]
                        float dr[16];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 v = *reinterpret_cast<const float4 *>(dcol + 4 * k);
                            dr[4 * k] = v.x, dr[4 * k + 1] = v.y, dr[4 * k + 2] = v.z, dr[4 * k + 3] = v.w;
                        }
                        float a0 = 0.f, a1 = 0.f;
#pragma unroll
                        for (int r = 0; r < 16; r += 2) {
                            a0 = fmaf(dr[r], __shfl_sync(kFull, vj, lb + r), a0);  // rows r < jj of the column are zero
                            a1 = fmaf(dr[r + 1], __shfl_sync(kFull, vj, lb + r + 1), a1);
                        }
                        amine = a0 + a1;
                    } else {
                        const float *colp = lp + Lay::Base(kb) + jj * Lay::Stride(kb);  // rows c0.. of column c0 + jj
                        float lblk[16];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 v = *reinterpret_cast<const float4 *>(colp + 4 * k);
                            lblk[4 * k] = v.x, lblk[4 * k + 1] = v.y, lblk[4 * k + 2] = v.z, lblk[4 * k + 3] = v.w;
                        }
                        const float zj = al[c0 + jj];
                        const float rsj = rs[c0 + jj];
                        amine = 0.f;
#pragma unroll
                        for (int c = 15; c >= 0; --c) {
                            const float a = (zj - sj) * rsj;
                            const float ac = __shfl_sync(kFull, a, lb + c);
                            if (jj == c) { amine = a; }
                            if (mine && jj < c) { sj = fmaf(lblk[c], ac, sj); }  // L(c0 + c, c0 + jj) * alpha_c
                        }
                    }
                    if (mine) { al[c0 + jj] = amine; }
                }
                __syncthreads();
#pragma unroll
                for (int sl = 0; sl < kColSlots; ++sl) {
                    const int col = tid + kThr * sl;
                    if (col < c0) {
                        const int cb = col >> 4;
                        const float *colp = lp + Lay::Base(cb) + (col & 15) * Lay::Stride(cb) + (c0 - 16 * cb);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 lv = *reinterpret_cast<const float4 *>(colp + 4 * k);
                            const float4 av = *reinterpret_cast<const float4 *>(al + c0 + 4 * k);
                            s[sl] = fmaf(lv.x, av.x, s[sl]);
                            s[sl] = fmaf(lv.y, av.y, s[sl]);
                            s[sl] = fmaf(lv.z, av.z, s[sl]);
                            s[sl] = fmaf(lv.w, av.w, s[sl]);
                        }
                    }
                }
            }
        }

        // --------------------------------------------------------------------------------------
        // predict one tile of up to 64 queries.  A thread OCTET owns 4 queries; thread h of the octet keeps rows
        // 32 M + 4 h + {0..3} (M = 0 .. n/32 - 1, "superblocks") of the four V = L^-1 k* columns in registers: 64
        // registers at n = 128, so the kernel fits the 128-register / 4-CTAs-per-SM budget, and every LDS.128 of L (4
        // rows of one column, 4 shared-memory wavefronts per warp) feeds 16 FMAs per thread = 128 FMAs per wavefront
        // and cycle, which is what the FFMA2 pipe can absorb (with 2 queries per thread the loop was shared-memory
        // bound at half that: ncu l1tex wavefronts 75 % of the cycles).
        // The substitution is right-looking: column j finalises v_j = r_j / L_jj in its owner thread, one shuffle
        // per query hands it to the octet, every thread then updates its rows below j.  The loop over the
        // 16-column blocks is a runtime loop; the code of a block is specialised on its superblock index so that all
        // register indices are static (a fully unrolled substitution was 110 KB of SASS and starved the instruction
        // cache).  Rows on / above the diagonal of the current superblock are updated with stored zeros or, once
        // they are final, with whatever the packed column holds there (dead registers).
        // --------------------------------------------------------------------------------------
        constexpr int kTileQ = kThreads / 2;  // queries per tile

        struct Slot {          // 4 rows x 4 queries
            float2 q[4][2];    // q[query][0] = rows (0,1), q[query][1] = rows (2,3)
        };

        __device__ __forceinline__ void
        SlotUpdate(Slot &v, const float4 lv, const float (&nv)[4]) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v.q[k][0] = Fma2(make_float2(lv.x, lv.y), nv[k], v.q[k][0]);
                v.q[k][1] = Fma2(make_float2(lv.z, lv.w), nv[k], v.q[k][1]);
            }
        }

        __device__ __forceinline__ float4
        Lds4(const float *__restrict__ ptr) {
            return *reinterpret_cast<const float4 *>(ptr);
        }

        // ld.shared.v4.f32 [addr + OFF]: one base register per column and immediate offsets.  (Left to itself the
        // compiler re-derived every address of the column from the block index: ~8 integer instructions per LDS.)
        template<int OFF>
        __device__ __forceinline__ float4
        LdsOff(const uint32_t addr) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
            return v;
        }

        template<int S, int CNT, int N>
        __device__ __forceinline__ void
        LoadBelow(float4 (&l)[N], const uint32_t addr) {  // l[S] = my rows of superblock (current + 1 + S): + 128 B per superblock
            if constexpr (S < CNT) {
                l[S] = LdsOff<128 * (S + 1)>(addr);
                LoadBelow<S + 1, CNT, N>(l, addr);
            }
        }

        // One 16-column block (PART = 0 / 1: first / second half of superblock M) of the substitution.
        //   cur_a: shared address of element (row 32 M + 4 h, column j0) of L, j0 = first column of the block
        template<int M, int NSB>
        __device__ __forceinline__ void
        SolveBlock(Slot (&v)[NSB], float (&ss)[4], uint32_t cur_a, const uint32_t stride_b, const float *__restrict__ rsp, const int lane, const int part) {
            constexpr int kBelow = NSB - 1 - M;
            for (int hq = 0; hq < 4; ++hq) {
                const int src = (lane & ~7) | (4 * part + hq);  // owner of columns 4 hq .. 4 hq + 3 of this block
                const float4 rs4 = *reinterpret_cast<const float4 *>(rsp + 4 * hq);
                const float rsv[4] = {rs4.x, rs4.y, rs4.z, rs4.w};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 lcur = LdsOff<0>(cur_a);
                    float4 lb[kBelow > 0 ? kBelow : 1];
                    LoadBelow<0, kBelow, (kBelow > 0 ? kBelow : 1)>(lb, cur_a);
                    float nv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 pr = v[M].q[k][r >> 1];
                        const float c = ((r & 1) ? pr.y : pr.x) * rsv[r];
                        const float vj = __shfl_sync(kFull, c, src);
                        ss[k] = fmaf(vj, vj, ss[k]);
                        nv[k] = -vj;
                    }
                    SlotUpdate(v[M], lcur, nv);
#pragma unroll
                    for (int s = 0; s < kBelow; ++s) { SlotUpdate(v[M + 1 + s], lb[s], nv); }
                    cur_a += stride_b;
                }
            }
        }

        template<int LO, int HI, int NSB>
        __device__ __forceinline__ void
        SolveBlockDispatch(const int m, Slot (&v)[NSB], float (&ss)[4], const uint32_t cur_a, const uint32_t stride_b, const float *__restrict__ rsp, const int lane, const int part) {
            if constexpr (HI - LO == 1) {
                SolveBlock<LO, NSB>(v, ss, cur_a, stride_b, rsp, lane, part);
            } else {
                constexpr int kMid = (LO + HI) / 2;
                if (m < kMid) {
                    SolveBlockDispatch<LO, kMid, NSB>(m, v, ss, cur_a, stride_b, rsp, lane, part);
                } else {
                    SolveBlockDispatch<kMid, HI, NSB>(m, v, ss, cur_a, stride_b, rsp, lane, part);
                }
            }
        }

        template<int XDIM, int NBLK>
        __device__ __forceinline__ void
        PredictTile(const BatchParams<float> &p, const CovCoef cov, const float *__restrict__ smem, const int n, const int nblk, const long q_begin, const int nq) {
            using Lay = Layout<NBLK>;
            static_assert(NBLK % 2 == 0, "superblocks are 32 rows");
            constexpr int kNsb = NBLK / 2;
            const float *lp = smem + Lay::kL;
            const float4 *pts = reinterpret_cast<const float4 *>(smem + Lay::kPts);
            const float *rs = smem + Lay::kRs;
            const int tid = threadIdx.x;
            const int lane = tid & 31;
            const int h = tid & 7;
            const int octet = tid >> 3;

            float xq[4][XDIM];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int d = 0; d < XDIM; ++d) { xq[k][d] = 4 * octet + k < nq ? p.q_x[(q_begin + 4 * octet + k) * XDIM + d] : 0.f; }
            }

            // Ktest entries of my rows for the four queries (never stored anywhere but registers); mean on the way
            Slot v[kNsb];
            float mean[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int m = 0; m < kNsb; ++m) {
                const int row0 = 32 * m + 4 * h;
                float kv[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 pt = pts[row0 + r];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float a = cov(Dist2<XDIM>(pt, xq[k]));
                        if (row0 + r >= n) { a = 0.f; }
                        mean[k] = fmaf(a, pt.w, mean[k]);
                        kv[k][r] = a;
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    v[m].q[k][0] = make_float2(kv[k][0], kv[k][1]);
                    v[m].q[k][1] = make_float2(kv[k][2], kv[k][3]);
                }
            }

            float ss[4] = {0.f, 0.f, 0.f, 0.f};
            for (int jb = 0; jb < nblk; ++jb) {
                const int m = jb >> 1;
                const int part = jb & 1;
                const uint32_t stride_b = 4u * static_cast<uint32_t>(Lay::Stride(jb));
                // element (row 32 m + 4 h, column 16 jb) of L: column block jb starts at row 16 jb
                const uint32_t cur_a = static_cast<uint32_t>(__cvta_generic_to_shared(lp + Lay::Base(jb) + (32 * m + 4 * h - 16 * jb)));
                SolveBlockDispatch<0, kNsb, kNsb>(m, v, ss, cur_a, stride_b, rs + 16 * jb, lane, part);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                mean[k] += __shfl_xor_sync(kFull, mean[k], 1);
                mean[k] += __shfl_xor_sync(kFull, mean[k], 2);
                mean[k] += __shfl_xor_sync(kFull, mean[k], 4);
            }

            if (h < 4) {
                const int myq = 4 * octet + h;
                if (myq < nq) {
                    const long src = q_begin + myq;
                    const long dst = p.q_out_index != nullptr ? p.q_out_index[src] : src;
                    const float mk = h == 0 ? mean[0] : h == 1 ? mean[1] : h == 2 ? mean[2] : mean[3];
                    const float sk = h == 0 ? ss[0] : h == 1 ? ss[1] : h == 2 ? ss[2] : ss[3];
                    if (p.mean != nullptr) {
                        float f = mk;
                        if (p.mapping != ERL_GP_MAPPING_NONE) { f = MappingInv<float>(p.mapping, p.mapping_scale, f); }
                        p.mean[dst] = f;
                    }
                    if (p.variance != nullptr) { p.variance[dst] = 1.0f - sk; }  // literal prior 1.0f, src/vanilla_gp.cpp:121
                    if (p.valid != nullptr) { p.valid[dst] = 1; }
                }
            }
        }

        // --------------------------------------------------------------------------------------
        // Tensor-path predict (3xTF32 on mma.sync.m16n8k8): V^T = Kt^T L^-T, one warp per 16 queries.
        //
        // The substitution is written for the TRANSPOSED system so that a finished 16-column block of V^T, which the MMA
        // leaves in the accumulator ("C") layout, is directly the A operand of the next products: C holds (row g, columns
        // 2t, 2t+1), A wants (row g, k-slots t, t+4), and since the order of the reduction index is free the two columns of
        // C simply become the slots t and t+4 - the B fragment is loaded with the matching permutation, b0 = B[2t][g],
        // b1 = B[2t+1][g].  No shuffle, no shared-memory round trip, no barrier: a warp walks the 16-column blocks of L
        //     V^T_j = X_j Dinv_j^T          (X_j = the accumulated block, Dinv_j = inverse of the 16 x 16 diagonal block)
        //     X_i  -= V^T_j L_ij^T          for every block row i below j
        // with all 16 x 128 accumulators (64 registers) resident.  L stays in the packed column-major layout of the
        // factorisation: b0 / b1 are two LDS.32 whose 32 lanes hit 32 different banks (column stride == 4 mod 16).
        // FP32 accuracy comes from the 3xTF32 split a = hi + lo (hi = a truncated to TF32, lo = a - hi exact): a b ~ hi_a hi_b +
        // lo_a hi_b + hi_a lo_b, error ~2^-20 relative per product (measured in tests/test_gpu_batch.py).
        // One m16n8k8 is 1024 FMAs per issue slot instead of 64 for a warp-wide FFMA2: the FMA / issue pipes that bound the
        // previous predict are left to the factorisations of the other resident CTAs.
        // --------------------------------------------------------------------------------------
        // inverses of the 16 x 16 diagonal blocks of L (lane c < 16 of a warp: column c by forward substitution on e_c)
        template<int NBLK>
        __device__ __forceinline__ void
        ComputeDinv(float *__restrict__ smem, const int nblk) {
            using Lay = Layout<NBLK>;
            const float *lp = smem + Lay::kL;
            const float *rs = smem + Lay::kRs;
            float *dinv = smem + Lay::kDinv;
            const int warp = threadIdx.x >> 5;
            const int lane = threadIdx.x & 31;
            if (lane >= 16) { return; }
            for (int kb = warp; kb < nblk; kb += ThreadsFor<NBLK>::value / 32) {
                const float *blk = lp + Lay::Base(kb);
                const int stride = Lay::Stride(kb);
                float sres[16], x[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { sres[i] = i == lane ? 1.0f : 0.f; }
#pragma unroll
                for (int pc = 0; pc < 16; ++pc) {
                    x[pc] = sres[pc] * rs[16 * kb + pc];
                    float col[16];
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const float4 v4 = *reinterpret_cast<const float4 *>(blk + pc * stride + 4 * k4);  // warp-uniform address
                        col[4 * k4] = v4.x, col[4 * k4 + 1] = v4.y, col[4 * k4 + 2] = v4.z, col[4 * k4 + 3] = v4.w;
                    }
#pragma unroll
                    for (int i = pc + 1; i < 16; ++i) { sres[i] = fmaf(-col[i], x[pc], sres[i]); }
                }
                float *dst = dinv + kb * 16 * Lay::kDinvLd + lane * Lay::kDinvLd;  // column `lane` of the inverse
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) { *reinterpret_cast<float4 *>(dst + 4 * k4) = make_float4(x[4 * k4], x[4 * k4 + 1], x[4 * k4 + 2], x[4 * k4 + 3]); }
            }
        }

        // FULL: nblk == NBLK is known at compile time (every "is this block row active" test folds away: the common case of a
        // full GP then runs branch-free straight-line code; with the runtime tests every group of HMMAs ends in a taken branch
        // over its short-row variant, and the fused kernel spent 18 % of its warp samples waiting for instructions)
        template<int XDIM, int NBLK, bool FULL>
        __device__ __forceinline__ void
        PredictTileMma(const BatchParams<float> &p, const CovCoef cov, const float *__restrict__ smem, const int n, const int nblk_rt, const long q_begin, const int nq) {
            using Lay = Layout<NBLK>;
            const int nblk = FULL ? NBLK : nblk_rt;
            const float *lp = smem + Lay::kL;
            const float4 *pts = reinterpret_cast<const float4 *>(smem + Lay::kPts);
            const float *dinv = smem + Lay::kDinv;
            const int tid = threadIdx.x;
            const int warp = __shfl_sync(kFull, tid >> 5, 0);
            const int lane = tid & 31;
            const int g = lane >> 2, t = lane & 3;
            if (16 * warp >= nq) { return; }  // warp-uniform; PredictTileMma has no barrier
            const int qrow[2] = {16 * warp + g, 16 * warp + g + 8};

            float xq[2][XDIM];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
#pragma unroll
                for (int d = 0; d < XDIM; ++d) { xq[k][d] = qrow[k] < nq ? p.q_x[(q_begin + qrow[k]) * XDIM + d] : 0.f; }
            }

            // Ktest^T tile in the accumulator layout: acc[j] = columns 8 j + 2 t, + 1 of query rows g (slots 0, 1) and g + 8 (slots 2, 3);
            // the two neighbouring columns of a query are one packed (f32x2) evaluation
            float acc[2 * NBLK][4];
            float mean[2];
            {
                const float2 *soa = reinterpret_cast<const float2 *>(smem + Lay::kSoa);
                float negq[2][XDIM];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
#pragma unroll
                    for (int d = 0; d < XDIM; ++d) { negq[k][d] = -xq[k][d]; }
                }
                float2 mean2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                for (int j = 0; j < 2 * NBLK; ++j) {
                    if (FULL || 8 * j < 16 * nblk) {
                        float2 pc[XDIM];
#pragma unroll
                        for (int d = 0; d < XDIM; ++d) { pc[d] = soa[d * (Lay::kNp / 2) + 4 * j + t]; }
                        const float2 av = soa[3 * (Lay::kNp / 2) + 4 * j + t];
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            float2 kv = CovPair(cov, Dist2Pair<XDIM>(pc, negq[k]));
                            if ((!FULL || j >= 2 * NBLK - 2) && 8 * j + 8 > n) {  // padding columns (warp-uniform test; FULL: only the last block can have any)
                                if (8 * j + 2 * t >= n) { kv.x = 0.f; }
                                if (8 * j + 2 * t + 1 >= n) { kv.y = 0.f; }
                            }
                            mean2[k] = __ffma2_rn(kv, av, mean2[k]);
                            acc[j][2 * k] = kv.x;
                            acc[j][2 * k + 1] = kv.y;
                        }
                    } else {
                        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
                    }
                }
                mean[0] = mean2[0].x + mean2[0].y;
                mean[1] = mean2[1].x + mean2[1].y;
            }

            // The loop over the blocks is fully unrolled (static register indices).  (A runtime loop with the accumulators
            // rotated down by one block per step halves the code size but ran 25 % slower: no overlap across the back edge.)
            // Issue order inside a step: the three products of the 3xTF32 split go to the same accumulator and a dependent
            // HMMA waits ~20 cycles for its predecessor, so the products are issued term by term over a GROUP of accumulator
            // tiles (two block rows = 4 tiles) - consecutive HMMAs never touch the same accumulator.
            float ss[2] = {0.f, 0.f};
            StaticFor<0, NBLK>([&](auto jb_c) {
                constexpr int jb = decltype(jb_c)::value;
                if (jb < nblk) {
                    // V^T_jb = X_jb Dinv_jb^T
                    float v[2][4];
                    MulDinvT<Lay::kDinvLd>(acc[2 * jb], acc[2 * jb + 1], dinv + jb * 16 * Lay::kDinvLd + 2 * t * Lay::kDinvLd + g, v);
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        ss[0] = fmaf(v[nt][0], v[nt][0], ss[0]);
                        ss[0] = fmaf(v[nt][1], v[nt][1], ss[0]);
                        ss[1] = fmaf(v[nt][2], v[nt][2], ss[1]);
                        ss[1] = fmaf(v[nt][3], v[nt][3], ss[1]);
                    }
                    if (jb + 1 < nblk) {
                        // X_i -= V^T_jb L_{i,jb}^T for the block rows below, two block rows (4 accumulator tiles) per group
                        uint32_t ahi[2][4], alo[2][4];
                        AccToA(v[0], -1.0f, ahi[0], alo[0]);
                        AccToA(v[1], -1.0f, ahi[1], alo[1]);
                        constexpr int stride = Lay::Stride(jb);
                        const float *lb = lp + Lay::Base(jb) + 2 * t * stride + g - 16 * jb;  // + row: L[row + g][16 jb + 2 t]
                        auto update = [&](auto i_c, auto cnt_c) {  // block rows i .. i + cnt - 1
                            constexpr int i = decltype(i_c)::value;
                            constexpr int kTiles = 2 * decltype(cnt_c)::value;
#pragma unroll
                            for (int kt = 0; kt < 2; ++kt) {
                                float b[kTiles][2];
                                uint32_t bl[kTiles][2];
#pragma unroll
                                for (int u = 0; u < kTiles; ++u) {
                                    const float *bp = lb + 8 * kt * stride + 16 * i + 8 * u;  // L[16 i + 8 u + g][16 jb + 8 kt + 2 t (+ 1)]
                                    b[u][0] = bp[0];
                                    b[u][1] = bp[stride];
                                }
#pragma unroll
                                for (int u = 0; u < kTiles; ++u) { Tf32LoPair(b[u][0], b[u][1], bl[u][0], bl[u][1]); }
#pragma unroll
                                for (int u = 0; u < kTiles; ++u) { MmaTf32(acc[2 * i + u], alo[kt], __float_as_uint(b[u][0]), __float_as_uint(b[u][1])); }
#pragma unroll
                                for (int u = 0; u < kTiles; ++u) { MmaTf32(acc[2 * i + u], ahi[kt], bl[u][0], bl[u][1]); }
#pragma unroll
                                for (int u = 0; u < kTiles; ++u) { MmaTf32(acc[2 * i + u], ahi[kt], __float_as_uint(b[u][0]), __float_as_uint(b[u][1])); }
                            }
                        };
                        StaticFor<0, (NBLK - jb) / 2>([&](auto grp_c) {
                            constexpr int i = jb + 1 + 2 * decltype(grp_c)::value;
                            if constexpr (i + 1 < NBLK) {
                                if (i + 1 < nblk) {
                                    update(std::integral_constant<int, i>{}, std::integral_constant<int, 2>{});
                                } else if (i < nblk) {
                                    update(std::integral_constant<int, i>{}, std::integral_constant<int, 1>{});
                                }
                            } else if constexpr (i < NBLK) {
                                if (i < nblk) { update(std::integral_constant<int, i>{}, std::integral_constant<int, 1>{}); }
                            }
                        });
                    }
                }
            });
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                mean[k] += __shfl_xor_sync(kFull, mean[k], 1);
                mean[k] += __shfl_xor_sync(kFull, mean[k], 2);
                ss[k] += __shfl_xor_sync(kFull, ss[k], 1);
                ss[k] += __shfl_xor_sync(kFull, ss[k], 2);
            }
            if (t < 2 && qrow[t] < nq) {
                const long src = q_begin + qrow[t];
                const long dst = p.q_out_index != nullptr ? p.q_out_index[src] : src;
                const float mk = t == 0 ? mean[0] : mean[1];
                const float sk = t == 0 ? ss[0] : ss[1];
                if (p.mean != nullptr) {
                    float f = mk;
                    if (p.mapping != ERL_GP_MAPPING_NONE) { f = MappingInv<float>(p.mapping, p.mapping_scale, f); }
                    p.mean[dst] = f;
                }
                if (p.variance != nullptr) { p.variance[dst] = 1.0f - sk; }  // literal prior 1.0f, src/vanilla_gp.cpp:121
                if (p.valid != nullptr) { p.valid[dst] = 1; }
            }
        }

#ifndef ERL_GP_ROWGP_FFMA_PREDICT
        constexpr bool kMmaPredict = true;
#else
        constexpr bool kMmaPredict = false;  // A/B builds of the FFMA2 predict (PredictTile)
#endif

        template<int XDIM, int NBLK, int MODE>
        __global__ void __launch_bounds__(ThreadsFor<NBLK>::value, NBLK <= 8 ? 4 : (NBLK <= 12 ? 2 : 1))
        RowGpKernel(const BatchParams<float> p) {
            using Lay = Layout<NBLK>;
            constexpr int kThr = ThreadsFor<NBLK>::value;  // threads of this instance
            constexpr int kQTile = kThr / 2;                 // queries per predict pass: 16 per warp
            extern __shared__ __align__(16) unsigned char smem_raw[];
            float *smem = reinterpret_cast<float *>(smem_raw);
            float *lp = smem + Lay::kL;
            float4 *pts = reinterpret_cast<float4 *>(smem + Lay::kPts);
            float *rs = smem + Lay::kRs;
            float *al = smem + Lay::kAl;
            float *sv = smem + Lay::kVar;

            const int g = blockIdx.x;
            const int tid = threadIdx.x;
            const int warp = __shfl_sync(kFull, tid >> 5, 0);  // warp-uniform by construction: role branches need no reconvergence code
            const int lane = tid & 31;
            if constexpr (MODE == kBatchTrainPredict) {
                // The CTAs of one SM start together and do identical work, so they would stay in lock-step: all four in the
                // barrier-bound factorisation, then all four on the tensor pipe.  Delaying the slots of the first wave by a
                // fraction of the per-GP time keeps them out of phase for the rest of the launch (the hardware refills a
                // slot when its CTA retires), so that factorisations overlap with tensor-path predicts.
                if (p.stagger_cycles > 0 && g < 4 * p.sm_count) {
                    const long long wait = static_cast<long long>(g / p.sm_count) * p.stagger_cycles;
                    const long long t0 = clock64();
                    while (clock64() - t0 < wait) { __nanosleep(500); }
                }
            }
            const int n = p.n_train[g];
            const long q0 = (MODE & kBatchPredict) ? p.q_offsets[g] : 0;
            const long q1 = (MODE & kBatchPredict) ? p.q_offsets[g + 1] : 0;

            if constexpr ((MODE & kBatchTrain) != 0) {
                if (n <= p.min_train || n <= 0) {  // the reference's `cnt > min_num_samples_per_group` / `cnt > 0` gate
                    if (tid == 0) { p.info[g] = -1; }
                    if ((MODE & kBatchPredict) && p.valid != nullptr) {
                        for (long q = q0 + tid; q < q1; q += kThr) { p.valid[p.q_out_index != nullptr ? p.q_out_index[q] : q] = 0; }
                    }
                    return;
                }
            } else {
                if (p.info[g] != 0 || q1 <= q0) { return; }  // untrained / failed GP: outputs stay untouched
            }
            const int nblk = (n + 15) >> 4;
            const int npr = nblk * 16;
            const CovCoef cov(p.cov);

            // ---- stage the training inputs (everything below works on the first 16 * nblk rows / columns only) ----
            const float *gx = p.x + static_cast<long>(g) * p.max_n * XDIM;
            for (int e = tid; e < Lay::kNp; e += kThr) {
                float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e < n) {
                    pt.x = gx[e * XDIM];
                    if (XDIM > 1) { pt.y = gx[e * XDIM + (XDIM > 1 ? 1 : 0)]; }
                    if (XDIM > 2) { pt.z = gx[e * XDIM + (XDIM > 2 ? 2 : 0)]; }
                    if (!(MODE & kBatchTrain)) { pt.w = p.alpha[static_cast<long>(g) * p.max_n + e]; }
                }
                pts[e] = pt;
                rs[e] = 1.0f;
                smem[Lay::kSoa + e] = pt.x;
                smem[Lay::kSoa + Lay::kNp + e] = pt.y;
                smem[Lay::kSoa + 2 * Lay::kNp + e] = pt.z;
                smem[Lay::kSoa + 3 * Lay::kNp + e] = pt.w;  // alpha (predict-only mode; the train modes fill it after the back-substitution)
            }

            if constexpr ((MODE & kBatchTrain) != 0) {
                const float *gy = p.y + static_cast<long>(g) * p.max_n;
                const float *gv = p.var + static_cast<long>(g) * p.max_n;
                for (int e = tid; e < Lay::kNp; e += kThr) {
                    al[e] = e < n ? gy[e] : 0.f;
                    sv[e] = e < n ? gv[e] : 0.f;
                }
                __syncthreads();
                int fail;
                if constexpr (kMmaTrain && kLookAhead && NBLK <= 8) {
                    fail = FactorizeMmaLa<XDIM, NBLK>(cov, smem, n, nblk);
                } else if constexpr (kMmaTrain || NBLK > 8) {  // the FFMA version is thread-per-row: n <= 128 only
                    fail = FactorizeMma<XDIM, NBLK>(cov, smem, n, nblk);
                } else {
                    fail = Factorize<XDIM, NBLK>(cov, smem, n, nblk);
                }
                int *s_fail = reinterpret_cast<int *>(smem + Lay::kMisc);
                if (tid == 0) { *s_fail = fail; }  // warp 0 tracked every pivot
                __syncthreads();
                const int failed = *s_fail;
                if (failed != 0) {
                    if (tid == 0) { p.info[g] = failed; }
                    if ((MODE & kBatchPredict) && p.valid != nullptr) {
                        for (long q = q0 + tid; q < q1; q += kThr) { p.valid[p.q_out_index != nullptr ? p.q_out_index[q] : q] = 0; }
                    }
                    return;
                }
                // ---- L write-back (coalesced, one column per warp and step), issued before the back-substitution ----
                if (p.write_l && p.tma_writeback && (p.max_n & 3) == 0 && (n & 3) == 0) {
                    // One bulk asynchronous copy (TMA engine) per column: the stored part of column c (rows 16 cb .. n, 16-byte aligned in
                    // the packed layout and in the caller's column-major slice) goes to HBM without passing through registers, and
                    // the copies run under the back-substitution and the predict.  Rows above the column's diagonal block are never
                    // written: the slice starts from zeros (erl_gp_batch_create) and only lower-triangular entries are ever stored.
                    float *gl = p.l + static_cast<long>(g) * p.max_n * p.max_n;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the generic-proxy stores of L, ordered before the async-proxy reads
                    for (int c = tid; c < n; c += kThr) {
                        const int cb = c >> 4;
                        const uint32_t src = static_cast<uint32_t>(__cvta_generic_to_shared(lp + Lay::Base(cb) + (c & 15) * Lay::Stride(cb)));
                        float *dst = gl + static_cast<long>(c) * p.max_n + 16 * cb;
                        const uint32_t bytes = static_cast<uint32_t>(n - 16 * cb) * 4u;
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                } else if (p.write_l) {
                    float *gl = p.l + static_cast<long>(g) * p.max_n * p.max_n;
                    if ((p.max_n & 3) == 0) {
                        constexpr int kRowChunks = (Lay::kNp + 127) / 128;  // 128 rows per warp and step
                        if ((n & 3) == 0) {
                            // fast path (no ragged float4): one LDS.128 + one STG.128 per lane and column, nothing else
                            for (int cb = 0; cb < nblk; ++cb) {
                                const float *blk = lp + Lay::Base(cb) - 16 * cb + warp * Lay::Stride(cb);
                                const int stride4 = (kThr / 32) * Lay::Stride(cb);
                                float *gcol = gl + static_cast<long>(16 * cb + warp) * p.max_n;
#pragma unroll
                                for (int k = 0; k < 16 / (kThr / 32); ++k) {
                                    if (16 * cb + warp + (kThr / 32) * k < n) {
#pragma unroll
                                        for (int ch = 0; ch < kRowChunks; ++ch) {
                                            const int r4 = 4 * lane + 128 * ch;
                                            if (r4 < n) {
                                                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                                                if (r4 >= 16 * cb) { v = *reinterpret_cast<const float4 *>(blk + k * stride4 + r4); }
                                                *reinterpret_cast<float4 *>(gcol + static_cast<long>((kThr / 32) * k) * p.max_n + r4) = v;
                                            }
                                        }
                                    }
                                }
                            }
                        } else
                        for (int cb = 0; cb < nblk; ++cb) {  // per column block: base / stride of the packed layout once
                            const float *blk = lp + Lay::Base(cb) - 16 * cb;
                            const int stride = Lay::Stride(cb);
#pragma unroll
                            for (int k = 0; k < 16 / (kThr / 32); ++k) {
                                const int c = 16 * cb + warp + (kThr / 32) * k;
                                if (c < n) {
                                    float *gcol = gl + static_cast<long>(c) * p.max_n;
#pragma unroll
                                    for (int ch = 0; ch < kRowChunks; ++ch) {
                                        const int r4 = 4 * lane + 128 * ch;
                                        if (r4 < n) {
                                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                                            if (r4 >= 16 * cb) { v = *reinterpret_cast<const float4 *>(blk + (c & 15) * stride + r4); }
                                            if (r4 + 3 < n) {
                                                *reinterpret_cast<float4 *>(gcol + r4) = v;
                                            } else {
                                                gcol[r4] = v.x;
                                                if (r4 + 1 < n) { gcol[r4 + 1] = v.y; }
                                                if (r4 + 2 < n) { gcol[r4 + 2] = v.z; }
                                            }
                                        }
                                    }
                                }
                            }
                        }
                    } else {
                        for (int c = warp; c < n; c += kThr / 32) {
                            const int cb = c >> 4;
                            for (int r = lane; r < n; r += 32) {
                                gl[r + static_cast<long>(c) * p.max_n] = r >= 16 * cb ? lp[Lay::Base(cb) + (c & 15) * Lay::Stride(cb) + (r - 16 * cb)] : 0.f;
                            }
                        }
                    }
                }
                BackSolve<NBLK, (kBackSolveDinv && (kMmaTrain || NBLK > 8))>(smem, nblk);
                __syncthreads();
                float *ga = p.alpha + static_cast<long>(g) * p.max_n;
                for (int e = tid; e < n; e += kThr) {
                    const float a = al[e];
                    ga[e] = a;
                    smem[Lay::kPts + 4 * e + 3] = a;
                    smem[Lay::kSoa + 3 * Lay::kNp + e] = a;
                }
                if (tid == 0) { p.info[g] = 0; }
                if constexpr ((kMmaPredict || NBLK > 8) && !(kMmaTrain || NBLK > 8) && (MODE & kBatchPredict) != 0) { ComputeDinv<NBLK>(smem, nblk); }  // FactorizeMma leaves Dinv behind
            } else {
                // ---- predict-only: reload L (float4 along the rows when the layout allows), rebuild 1 / L_jj ----
                __syncthreads();
                const float *gl = p.l + static_cast<long>(g) * p.max_n * p.max_n;
                const bool vec_ok = (p.max_n & 3) == 0;
                constexpr int kBatch = 8;  // columns in flight per warp: the loads of a batch are all issued before the first store
                constexpr int kRowChunks = (Lay::kNp + 127) / 128;
                for (int cc0 = warp * kBatch; cc0 < npr * kRowChunks; cc0 += (kThr / 32) * kBatch) {
                    const int c0 = kRowChunks == 1 ? cc0 : cc0 % npr;             // npr is a multiple of 16, kBatch divides 16: a batch never straddles the wrap
                    const int rchunk = kRowChunks == 1 ? 0 : 128 * (cc0 / npr);  // rows [16 cb + rchunk, + 128) of the column
                    float val[kBatch][4];
#pragma unroll
                    for (int u = 0; u < kBatch; ++u) {
                        const int c = c0 + u;
                        const int r4 = 16 * (c >> 4) + 4 * lane + rchunk;
                        const float *gcol = gl + static_cast<long>(c) * p.max_n;
                        if (vec_ok && c < n && r4 + 3 < n) {
                            const float4 t = *reinterpret_cast<const float4 *>(gcol + r4);
                            val[u][0] = t.x, val[u][1] = t.y, val[u][2] = t.z, val[u][3] = t.w;
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k) { val[u][k] = (c < n && r4 + k < n) ? gcol[r4 + k] : 0.f; }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kBatch; ++u) {
                        const int c = c0 + u;
                        const int cb = c >> 4;
                        const int r4 = 16 * cb + 4 * lane + rchunk;
                        if (c < npr && r4 < npr) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int r = r4 + k;
                                if (r < c) { val[u][k] = 0.f; }
                                if (r == c) {
                                    if (c >= n) { val[u][k] = 1.0f; }
                                    rs[c] = 1.0f / val[u][k];
                                }
                            }
                            *reinterpret_cast<float4 *>(lp + Lay::Base(cb) + (c & 15) * Lay::Stride(cb) + (r4 - 16 * cb)) = make_float4(val[u][0], val[u][1], val[u][2], val[u][3]);
                        }
                    }
                }
                if constexpr (kMmaPredict || NBLK > 8) {
                    __syncthreads();
                    ComputeDinv<NBLK>(smem, nblk);
                }
            }

            if constexpr ((MODE & kBatchPredict) != 0) {
                __syncthreads();
                for (long qb = q0 + static_cast<long>(blockIdx.y) * kQTile; qb < q1; qb += static_cast<long>(gridDim.y) * kQTile) {
                    const int nq = static_cast<int>(q1 - qb < kQTile ? q1 - qb : kQTile);
                    if constexpr (kMmaPredict || NBLK > 8) {
                        if (NBLK <= 8 && nblk == NBLK) {  // (the larger instances keep one variant: build time)
                            PredictTileMma<XDIM, (NBLK <= 8 ? NBLK : 2), (NBLK <= 8)>(p, cov, smem, n, nblk, qb, nq);
                        } else {
                            PredictTileMma<XDIM, NBLK, false>(p, cov, smem, n, nblk, qb, nq);
                        }
                    } else {
                        PredictTile<XDIM, NBLK>(p, cov, smem, n, nblk, qb, nq);
                    }
                }
            }
            if constexpr ((MODE & kBatchTrain) != 0) {
                // the bulk copies of L read this CTA's shared memory: they must have completed before it is released
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            }
        }

        template<int XDIM, int NBLK, int MODE>
        static int
        LaunchInstance(Context *ctx, const BatchParams<float> &params, const int tiles_per_gp) {
            using Lay = Layout<NBLK>;
            auto kernel = RowGpKernel<XDIM, NBLK, MODE>;
            if (static_cast<int>(Lay::kBytes) > ctx->max_smem_optin) {
                return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "row-GP kernel needs %zu B of shared memory, device allows %d", Lay::kBytes, ctx->max_smem_optin);
            }
            ERL_GP_CUDA_OK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(Lay::kBytes)));
            const dim3 grid(static_cast<unsigned>(params.num_gps), static_cast<unsigned>(tiles_per_gp < 1 ? 1 : tiles_per_gp));
            BatchParams<float> launch_params = params;
            if (MODE == kBatchTrainPredict) {
                static const char *env = std::getenv("ERL_GP_ROWGP_STAGGER");
                launch_params.stagger_cycles = env != nullptr ? std::atoi(env) : kDefaultStaggerCycles;
                launch_params.sm_count = ctx->sm_count;
            }
            {
                static const char *env = std::getenv("ERL_GP_ROWGP_TMA_WB");
                launch_params.tma_writeback = env != nullptr ? std::atoi(env) : kDefaultTmaWriteback;
            }
            kernel<<<grid, ThreadsFor<NBLK>::value, Lay::kBytes, ctx->stream>>>(launch_params);
            ctx->launches += 1;
            ERL_GP_CUDA_OK(ctx, cudaGetLastError());
            return ERL_GP_STATUS_OK;
        }

        template<int XDIM, int NBLK>
        int
        LaunchMode(Context *ctx, const BatchParams<float> &params, const int mode, const int tiles_per_gp) {
            switch (mode) {
                case kBatchTrain: return LaunchInstance<XDIM, NBLK, kBatchTrain>(ctx, params, 1);
                case kBatchPredict: return LaunchInstance<XDIM, NBLK, kBatchPredict>(ctx, params, tiles_per_gp);
                case kBatchTrainPredict: return LaunchInstance<XDIM, NBLK, kBatchTrainPredict>(ctx, params, 1);
                default: return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "batch: bad mode %d", mode);
            }
        }

        // The kernels are instantiated in their own translation units (erl_gp_rowgp_x<dim>_<a|b|c>.cu: n <= 128 / 192 / 256), so that
        // a clean build compiles them in parallel; everybody else only sees these declarations.
#ifdef ERL_GP_ROWGP_EXTERN_INSTANCES
#define ERL_GP_ROWGP_EXTERN(XD)                                                                          \
    extern template int LaunchMode<XD, 2>(Context *, const BatchParams<float> &, int, int);           \
    extern template int LaunchMode<XD, 4>(Context *, const BatchParams<float> &, int, int);           \
    extern template int LaunchMode<XD, 6>(Context *, const BatchParams<float> &, int, int);           \
    extern template int LaunchMode<XD, 8>(Context *, const BatchParams<float> &, int, int);           \
    extern template int LaunchMode<XD, 12>(Context *, const BatchParams<float> &, int, int);          \
    extern template int LaunchMode<XD, 16>(Context *, const BatchParams<float> &, int, int);
        ERL_GP_ROWGP_EXTERN(1)
        ERL_GP_ROWGP_EXTERN(2)
        ERL_GP_ROWGP_EXTERN(3)
#undef ERL_GP_ROWGP_EXTERN
#endif

        // max_n <= 256
        template<int XDIM>
        int
        Launch(Context *ctx, const BatchParams<float> &params, const int mode, const int tiles_per_gp) {
            const int max_n = params.max_n;
#ifdef ERL_GP_ROWGP_FAST_BUILD  // kernel experiments only (make EXTRA=-DERL_GP_ROWGP_FAST_BUILD): one instance, n <= 128
            if (max_n <= 128) { return LaunchMode<XDIM, 8>(ctx, params, mode, tiles_per_gp); }
            return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "fast build: max_n=%d > 128", max_n);
#else
            if (max_n <= 32) { return LaunchMode<XDIM, 2>(ctx, params, mode, tiles_per_gp); }
            if (max_n <= 64) { return LaunchMode<XDIM, 4>(ctx, params, mode, tiles_per_gp); }
            if (max_n <= 96) { return LaunchMode<XDIM, 6>(ctx, params, mode, tiles_per_gp); }
            if (max_n <= 128) { return LaunchMode<XDIM, 8>(ctx, params, mode, tiles_per_gp); }
            if (max_n <= 192) { return LaunchMode<XDIM, 12>(ctx, params, mode, tiles_per_gp); }  // 104 KB of shared memory, 2 CTAs / SM
            return LaunchMode<XDIM, 16>(ctx, params, mode, tiles_per_gp);                        // 171 KB, 1 CTA / SM
#endif
        }

    }  // namespace rowgp
}  // namespace erl_gp
