// FP32 fused train + predict for n <= 128 on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
//
// Same contract and the same reference code replaced as rowgp::RowGpKernel (erl_gp_rowgp.cuh):
//   VanillaGaussianProcess::UpdateKtrain + Solve        src/vanilla_gp.cpp:476-505
//   VanillaGaussianProcess::ComputeKtest                src/vanilla_gp.cpp:521-552
//   TestResult::GetMean / GetVariance                   src/vanilla_gp.cpp:61-150
// driven per partition by src/lidar_gp_2d.cpp:366-392 / src/range_sensor_gp_3d.cpp:334-360, and the batched solve
// BatchGaussianProcessUpdateTorch::Solve intended (src/batch_gp_update_torch.cpp:74-82).
//
// Formulation.  The first 128 queries of a GP ride along with the factorisation: with Kt = K(X, X*) the augmented matrix
//     [ K   ]            [ L   ]
//     [ Kt^T]  L^-T   =  [ V^T ]        (V = L^-1 Kt,  var = 1 - colsumsq(V),  mean = V^T z,  z = L^-1 y)
// is a right-looking blocked elimination with 16-column panels in which the query rows are just 128 more rows below the
// training rows.  Per panel j:  X_j = (panel columns of the running matrix),  pivot tile -> L_jj and Dinv_j = L_jj^-1,
// rows below: V_j = X_j Dinv_j^T,  trailing columns -= V_j L_j^T.  The trailing update carries all the flops and is a GEMM
// with M = 128 rows per group (T = training rows, Q = query rows), N = remaining columns, K = 16:
//   * accumulators (the running matrix) live in TMEM: 128 lanes x 128 FP32 columns per group = 256 columns per CTA, two
//     persistent CTAs per SM; thread = row, tcgen05.ld.32x32b.x16 hands a thread the 16 panel entries of its own row;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 with M = 128, N = 16 .. 112, K = 8; FP32 accuracy by the
//     3xTF32 split a b ~ lo_a hi_b + hi_a lo_b + hi_a hi_b (hi = the FP32 pattern itself - the tensor core drops the low 13
//     bits - lo = a - trunc(a), exact); the A operands V_hi / V_lo are written back into dead TMEM columns (tcgen05.st) and
//     read by the MMA from there, the B operands (rows of L_j, hi and lo) sit in shared memory in the K-major no-swizzle
//     canonical layout, double-buffered by panel parity; the products are subtracted by the a_negate bit of the descriptor.
// Warp roles (320 threads): warps 0-3 training rows, 4-7 query rows, 8 pivot tiles, 9 MMA issue + back-substitution.
//   * LOOK-AHEAD: the serial chain of a blocked Cholesky is pivot tile -> Dinv -> V -> update -> next pivot tile.  Here the
//     tensor-core round trip is taken off that chain: the 16 rows of tile j+1 hand the pivot warp their entries of panel j
//     (S_j) AND of panel j+1 (D_{j+1}), both with the updates of panels < j applied (from TMEM, after update j-1); the pivot
//     warp applies panel j itself on the FP32 pipe (L_{j+1,j} = S_j Dinv_j^T, D_{j+1} -= L_{j+1,j} L_{j+1,j}^T, 16^3
//     multiply-adds) and factorises tile j+1 while the row warps compute V_j and the tensor core runs update j.
//   * the query group never blocks the training group: separate operand-ready / MMA-done barriers per group.
//   * everything is synchronised with mbarriers; the only CTA-wide barriers frame the kernel, one 256-thread named barrier
//     per GP publishes the staged points.
//   * L goes to HBM straight from the registers of the row threads (one coalesced 128-byte line per warp and column) and
//     alpha = L^-T z is computed by the MMA warp from the shared-memory copy of L WHILE the next GP is factorised
//     (it sweeps the column blocks 7 -> 0, the next factorisation fills them 0 -> 7 and waits for it once, before its first
//     store).
// The Gram matrix and Ktest are never stored: their entries are generated in the thread that owns the row, right when the
// panel is consumed (fused distance + covariance + noise diagonal, rowgp::CovPair).  Further query tiles (beyond 128 queries
// per GP) reuse the shared-memory layout and the mma.sync predict of erl_gp_rowgp.cuh.
#pragma once

#include "erl_gp_rowgp.cuh"

namespace erl_gp {
    namespace rowgp_tc {

        using Lay = rowgp::Layout<8>;
        using rowgp::kFull;

        constexpr int kRowThreads = 256;  // warps 0-3: training rows (group T), warps 4-7: query rows (group Q)
        constexpr int kThreads = 320;     // + warp 8: pivot tiles, warp 9: MMA issue and back-substitution
        constexpr int kTmemCols = 256;    // T region: columns [0, 128), Q region: [128, 256)
        constexpr uint32_t kChunkBytes = 128 * 16;   // distance of the 16-byte K chunks of an operand buffer (LBO)
        constexpr int kOperandBytes = 128 * 16 * 4;  // 128 rows x 16 k x FP32
        constexpr int kBandLd = 36;  // floats per row of a band hand-over buffer: S (16), D (16), y (1), pad
        constexpr int kLbLd = 20;
#ifndef ERL_GP_TC_SLEEP_NS
#define ERL_GP_TC_SLEEP_NS 0  // back-off of the waits that are not on the critical path of a GP (A/B builds)
#endif

        // shared memory (bytes): the row-GP layout, then the MMA operand buffers and the small hand-over buffers
        constexpr size_t kOffB = (Lay::kBytes + 127) / 128 * 128;      // rows of L_j: [panel parity][hi, lo] (B operands; T's A_lo at panel 0)
        constexpr size_t kOffAqLo = kOffB + 4 * kOperandBytes;         // V_0 lo of the query rows (A operand of panel 0 only)
        constexpr size_t kOffSoa = kOffAqLo + kOperandBytes;           // staged training points, [GP parity][x, y, z, -][128]
        constexpr size_t kOffBand = kOffSoa + 2 * 4 * 128 * 4;         // band hand-over: [tile parity][16 rows][kBandLd]
        constexpr size_t kOffLb = kOffBand + 2 * 16 * kBandLd * 4;     // band warp scratch: L_{j+1,j}, k-major
        constexpr size_t kOffLt = kOffLb + 16 * kLbLd * 4;             // rows of L_jj as the pivot warp leaves them, [panel parity][16][kLbLd]
        constexpr size_t kOffBars = kOffLt + 2 * 16 * kLbLd * 4;       // 9 mbarriers (16 slots)
        constexpr size_t kOffSlot = kOffBars + 16 * 8;                 // TMEM base address, fail flags [GP parity]
        constexpr size_t kOffDbg = kOffSlot + 16;                      // -DERL_GP_TC_WATCHDOG_PRINT: one progress word per warp
        constexpr size_t kSmemBytes = kOffDbg + 64;
        static_assert(2 * (kSmemBytes + 1024) <= 228 * 1024, "two CTAs per SM");

        // mbarriers.  A waiter names a phase by its parity only, so a producer must never complete TWO phases of a barrier before
        // every consumer has looked at the first one (the wait would then block on the phase after).  Where the protocol lets the
        // producer run further ahead, the barrier is replicated and used round-robin:
        //   * band hand-over: tiles 0 and 1 are both handed over at panel 0 with nothing in between -> 2 barriers (tile t:
        //     barrier t & 1, phase t >> 1);
        //   * Dinv: with the look-ahead, pivot tile j + 1 only needs the BAND warp to have consumed Dinv_j; the other training
        //     warps are only bound by update j (Dinv_{j+2} needs it), the query warps by update j + 1 -> 4 barriers (panel p:
        //     barrier p & 3, phase p >> 2).
        //   The other barriers are safe with one instance: operands-ready / MMA-done alternate strictly per group, "factorisation
        //   done" and "back-substitution done" are chained through the pivot warp's wait at the start of a GP (see the kernel).
        enum Bar : int { kBarBand = 0 /* and 1 */, kBarDinv = 2 /* .. 5 */, kBarOpT = 6, kBarOpQ = 7, kBarMmaT = 8, kBarMmaQ = 9, kBarFact = 10, kBarEpi = 11 };

        // ---- PTX wrappers (syntax as in CUTLASS's cute/arch/mma_sm100_umma.hpp, copy_sm100.hpp, tmem_allocator_sm100.hpp) ----
        __device__ __forceinline__ uint32_t
        SmemAddr(const void *p) {
            return static_cast<uint32_t>(__cvta_generic_to_shared(p));
        }

        // K-major, no swizzle: core matrix = 8 rows x 16 bytes; LBO = distance of the two K chunks, SBO = distance of 8-row groups
        __device__ __forceinline__ uint64_t
        SmemDesc(const uint32_t saddr) {
            return static_cast<uint64_t>((saddr >> 4) & 0x3fff) | (static_cast<uint64_t>(kChunkBytes >> 4) << 16) | (static_cast<uint64_t>(128 >> 4) << 32) | (1ull << 46);
        }

        // kind::tf32 (a/b format 2), FP32 accumulator (c format 1), A negated, K-major A and B, M = 128
        __device__ __forceinline__ uint32_t
        InstrDesc(const int n) {
            return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 13) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
        }

        // The MMA warp executes the issue code convergently (all operands are warp-uniform, so they can live in uniform
        // registers); `elected` (one lane, from elect.sync) predicates the instruction itself.  The probe measured ~52 cycles
        // per MMA when a single divergent thread issues (tools/tcgen05_probe.cu).
        __device__ __forceinline__ uint32_t
        ElectOne() {
            uint32_t pred;
            asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}\n" : "=r"(pred));
            return pred;
        }

        __device__ __forceinline__ void
        MmaSS(const uint32_t elected, const uint32_t d_tmem, const uint64_t a_desc, const uint64_t b_desc, const uint32_t idesc, const uint32_t accumulate) {
            asm volatile(
                "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 e, %5, 0;\n\t"
                "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(elected)
                : "memory");
        }

        __device__ __forceinline__ void
        MmaTS(const uint32_t elected, const uint32_t d_tmem, const uint32_t a_tmem, const uint64_t b_desc, const uint32_t idesc, const uint32_t accumulate) {
            asm volatile(
                "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 e, %5, 0;\n\t"
                "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(elected)
                : "memory");
        }

        __device__ __forceinline__ void
        Commit(const uint32_t elected, const uint32_t bar) {
            asm volatile("{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %1, 0;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(bar), "r"(elected) : "memory");
        }

        __device__ __forceinline__ void
        MbarInit(const uint32_t bar, const uint32_t count) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
        }

        __device__ __forceinline__ void
        MbarArrive(const uint32_t bar) {
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(bar) : "memory");
        }

        __device__ __forceinline__ void
        MbarArriveCount(const uint32_t bar, const uint32_t count) {  // this thread arrives for `count` pending arrivals
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(bar), "r"(count) : "memory");
        }

        // try_wait suspends the thread until the phase completes or a system-dependent time limit passes.  A protocol error would
        // hang the GPU until the watchdog of the box fires, so the number of attempts is bounded and the kernel traps instead
        // (-DERL_GP_TC_WATCHDOG_PRINT names the barrier first: it costs ~25 instructions per wait, and the instruction cache
        // is what this kernel runs out of; -DERL_GP_TC_NO_WATCHDOG removes the counter).
        template<int SLEEP_NS = 0>
        __device__ __forceinline__ void
        MbarWait(const uint32_t bar, const uint32_t parity) {
            uint32_t done = 0;
#ifdef ERL_GP_TC_WATCHDOG_PRINT
            const long long t_start = clock64();
#elif !defined(ERL_GP_TC_NO_WATCHDOG)
            uint32_t tries = 0;
#endif
            while (true) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                if (done) { break; }
                if (SLEEP_NS > 0) { __nanosleep(SLEEP_NS); }  // waits off the critical path: keep the polling out of the issue slots
#ifdef ERL_GP_TC_WATCHDOG_PRINT
                if (clock64() - t_start > 1500000000ll) {  // debugging build: every waiter names its barrier, then the kernel traps
                    if ((threadIdx.x & 31) == 0) {
                        extern __shared__ __align__(1024) unsigned char smem_dbg[];
                        const volatile int *dbg = reinterpret_cast<const volatile int *>(smem_dbg + kOffDbg);
                        printf("mbarrier time-out: CTA %d warp %d barrier %u parity %u | progress (gp:panel:stage) %d:%d:%d %d:%d:%d %d:%d:%d %d:%d:%d | %d:%d:%d %d:%d:%d %d:%d:%d %d:%d:%d | %d:%d:%d %d:%d:%d\n",
                               blockIdx.x, threadIdx.x >> 5, (bar >> 3) & 15u, parity, dbg[0] >> 8, (dbg[0] >> 4) & 15, dbg[0] & 15, dbg[1] >> 8, (dbg[1] >> 4) & 15, dbg[1] & 15, dbg[2] >> 8,
                               (dbg[2] >> 4) & 15, dbg[2] & 15, dbg[3] >> 8, (dbg[3] >> 4) & 15, dbg[3] & 15, dbg[4] >> 8, (dbg[4] >> 4) & 15, dbg[4] & 15, dbg[5] >> 8, (dbg[5] >> 4) & 15, dbg[5] & 15,
                               dbg[6] >> 8, (dbg[6] >> 4) & 15, dbg[6] & 15, dbg[7] >> 8, (dbg[7] >> 4) & 15, dbg[7] & 15, dbg[8] >> 8, (dbg[8] >> 4) & 15, dbg[8] & 15, dbg[9] >> 8, (dbg[9] >> 4) & 15,
                               dbg[9] & 15);
                    }
                    for (int i = 0; i < 1000; ++i) { __nanosleep(1000000); }
                    asm volatile("trap;");
                }
#elif !defined(ERL_GP_TC_NO_WATCHDOG)
                if (++tries > (1u << 22)) { asm volatile("trap;"); }
#endif
            }
        }

        __device__ __forceinline__ void
        FenceBefore() {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }

        __device__ __forceinline__ void
        FenceAfter() {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }

        __device__ __forceinline__ void
        FenceProxyAsync() {  // generic-proxy shared-memory writes -> visible to the tensor core (async proxy)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }

        __device__ __forceinline__ void
        TmemLd16(const uint32_t taddr, float (&v)[16]) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
                "tcgen05.wait::ld.sync.aligned;\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
                  "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr)
                : "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) { v[i] = __uint_as_float(r[i]); }
        }

        __device__ __forceinline__ void
        TmemSt16(const uint32_t taddr, const uint32_t (&v)[16]) {  // no wait: see TmemStWait
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr), "r"(v[0]),
                         "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
                         "r"(v[15])
                         : "memory");
        }

        __device__ __forceinline__ void
        TmemStWait() {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }

        __device__ __forceinline__ void
        NamedBarrier(const int id, const int threads) {
            asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
        }

        // byte offset of the 16-byte chunk (row, k chunk kc) in an operand buffer
        __device__ __forceinline__ uint32_t
        OperandChunk(const int row, const int kc) {
            return static_cast<uint32_t>(kc) * kChunkBytes + static_cast<uint32_t>(row >> 3) * 128u + static_cast<uint32_t>(row & 7) * 16u;
        }

        // 16 covariance entries k(point, training point cb + c), c = 0 .. 15, on the packed FP32 pipe, as a rolled loop of four
        // steps (four entries each) in a rotating frame - static register indices, a quarter of the unrolled code.  Training
        // rows (TRAIN): zero padding beyond n, 1 + noise variance on the diagonal (1 for padding rows); query rows: zero
        // beyond n.
        template<int XDIM>
        __device__ __forceinline__ void
        PanelEntries(const rowgp::CovCoef &cov, const float2 *__restrict__ soa, const int cb, const float (&negp)[XDIM], const bool train, const int row, const int n, const float diag,
                     const bool need_mask, float (&x)[16]) {
#pragma unroll 2
            for (int it = 0; it < 4; ++it) {
                const int cq = cb + 4 * it;
                float e[4];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    float2 pc[XDIM];
#pragma unroll
                    for (int d = 0; d < XDIM; ++d) { pc[d] = soa[d * 64 + cq / 2 + m]; }
                    const float2 kv = rowgp::CovPair(cov, rowgp::Dist2Pair<XDIM>(pc, negp));
                    e[2 * m] = kv.x;
                    e[2 * m + 1] = kv.y;
                }
                if (need_mask) {  // warp-uniform: padding in this GP, or the diagonal crosses this warp's rows in these columns
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int col = cq + i;
                        if (col >= n || (train && row >= n)) { e[i] = 0.f; }
                        if (train && row == col) { e[i] = diag; }
                    }
                }
#pragma unroll
                for (int i = 0; i < 12; ++i) { x[i] = x[i + 4]; }
#pragma unroll
                for (int i = 0; i < 4; ++i) { x[12 + i] = e[i]; }
            }
        }

        // v = x Dinv^T for one row (Dinv column-major with leading dimension Lay::kDinvLd, zero above the diagonal), two
        // outputs per instruction on the packed FP32 pipe.  The row threads (V_j) and the pivot warp (L_{j+1,j}) run the same
        // instruction sequence, so both get the same bits.
        __device__ __forceinline__ void
        RowTimesDinvT(const float (&x)[16], const float *__restrict__ dj, float (&v)[16]) {
            float2 v2[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) { v2[c] = make_float2(0.f, 0.f); }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
#pragma unroll
                for (int c4 = k / 4; c4 < 4; ++c4) {
                    const float4 w = *reinterpret_cast<const float4 *>(dj + k * Lay::kDinvLd + 4 * c4);  // Dinv[4 c4 .. 4 c4 + 3][k]
                    v2[2 * c4] = rowgp::Fma2(make_float2(w.x, w.y), x[k], v2[2 * c4]);
                    v2[2 * c4 + 1] = rowgp::Fma2(make_float2(w.z, w.w), x[k], v2[2 * c4 + 1]);
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) { v[2 * c] = v2[c].x, v[2 * c + 1] = v2[c].y; }
        }

        // 16 x 16 pivot tile by one warp, as a ROLLED loop over the pivot columns (the instruction cache is 32 KB per SM and
        // five warp roles share it: the fully unrolled rowgp::PivotBlock alone is 9 KB).  Lanes 0-15 own the rows of the
        // tile, lanes 16-31 carry the unit vectors through the same elimination (= the columns of Dinv = L_jj^-1 for free).
        // The row is kept in a ROTATING frame: a[k] is the entry of column c + k at step c, so every register index is static.
        // LDL^T-style elimination (the reciprocal is off the shuffle chain), z = L^-1 y rides along.  Outputs go straight to
        // shared memory (lanes 0-15: L row, stride `ostride`; lanes 16-31: Dinv column, stride 1) and, for the L rows, to HBM.
        __device__ __forceinline__ void
        PivotTile(float (&a)[16], float zacc, const int c0, const int lane, int &fail, float *__restrict__ rs, float *__restrict__ al, float *__restrict__ out, const int ostride,
                  float *__restrict__ gout, const long gstride, const int gcols) {
            const int r = lane & 15;
#pragma unroll 1
            for (int c = 0; c < 16; ++c) {
                const float d = __shfl_sync(kFull, a[0], c);
                const float zc = __shfl_sync(kFull, zacc, c);
                float t[16];
#pragma unroll
                for (int k = 1; k < 16; ++k) { t[k] = __shfl_sync(kFull, a[0], c + k); }  // (lanes >= 16 - c: dead columns)
                if (!(d > 0.f) && fail == 0) { fail = c0 + c + 1; }
                float r0;
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d));
                const float rsv = r0 * fmaf(-0.5f * d * r0, r0, 1.5f);  // one Newton step
                const float invd = rsv * rsv;
                const float sc = a[0] * invd;
                const float lv = a[0] * rsv;
#pragma unroll
                for (int k = 1; k < 16; ++k) { a[k - 1] = fmaf(-sc, t[k], a[k]); }
                a[15] = 0.f;
                zacc = fmaf(-sc, zc, zacc);
                if (lane == 0) {
                    rs[c0 + c] = rsv;
                    al[c0 + c] = zc * rsv;
                }
                if (lane < 16) {
                    const float o = c > r ? 0.f : lv;
                    out[c * ostride] = o;
                    if (gout != nullptr && c < gcols) { gout[c * gstride] = o; }
                } else {
                    out[c] = lv;
                }
            }
        }

        __device__ __forceinline__ float
        Dot16(const float (&v)[16], const float *__restrict__ z) {  // z: 16 floats in shared memory, 16-byte aligned
            float dz = 0.f;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                const float4 zz = *reinterpret_cast<const float4 *>(z + 4 * k4);
                dz = fmaf(v[4 * k4], zz.x, dz);
                dz = fmaf(v[4 * k4 + 1], zz.y, dz);
                dz = fmaf(v[4 * k4 + 2], zz.z, dz);
                dz = fmaf(v[4 * k4 + 3], zz.w, dz);
            }
            return dz;
        }

        // 3xTF32 operands of one row: hi (the FP32 pattern) and lo into TMEM columns, optionally into operand buffers
        __device__ __forceinline__ void
        StoreOperandRow(const uint32_t sbuf, const int row, const uint32_t (&w)[16]) {
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + OperandChunk(row, kc)), "r"(w[4 * kc]), "r"(w[4 * kc + 1]), "r"(w[4 * kc + 2]), "r"(w[4 * kc + 3]) : "memory");
            }
        }

        // alpha = L^-T z by ONE warp, dot-product form: column blocks from the last to the first; z is read from Lay::kAl and
        // left alone (the query rows of this GP may still be reading it), alpha goes to Lay::kVar.  Step kb: s_c = sum_{r >= c0 + 16} L[r][c] alpha_r for the 16 columns of the block (two lanes per column,
        // rows dealt in chunks of four), alpha_blk = Dinv_kb^T (z_blk - s).
        __device__ __forceinline__ void
        BackSolveWarp(float *__restrict__ smem, const int nblk, const int lane) {
            const float *lp = smem + Lay::kL;
            const float *zs = smem + Lay::kAl;
            float *al = smem + Lay::kVar;
            const int c = lane & 15, h = lane >> 4;
            for (int kb = nblk - 1; kb >= 0; --kb) {
                const int c0 = 16 * kb;
                const float *colp = lp + Lay::Base(kb) + c * Lay::Stride(kb);  // element (c0, c0 + c)
                float s0 = 0.f, s1 = 0.f;
                const int chunks = 4 * (nblk - 1 - kb);  // 4-row chunks below the diagonal block
                for (int m = h; m < chunks; m += 2) {
                    const float4 lv = *reinterpret_cast<const float4 *>(colp + 16 + 4 * m);
                    const float4 av = *reinterpret_cast<const float4 *>(al + c0 + 16 + 4 * m);
                    s0 = fmaf(lv.x, av.x, s0);
                    s1 = fmaf(lv.y, av.y, s1);
                    s0 = fmaf(lv.z, av.z, s0);
                    s1 = fmaf(lv.w, av.w, s1);
                }
                float s = s0 + s1;
                s += __shfl_xor_sync(kFull, s, 16);
                const float w = zs[c0 + c] - s;  // both half-warps hold w_c in lane c / c + 16
                const float *dcol = smem + Lay::kDinv + kb * 16 * Lay::kDinvLd + c * Lay::kDinvLd;  // column c of Dinv_kb
                float dr[16];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 v4 = *reinterpret_cast<const float4 *>(dcol + 4 * k);
                    dr[4 * k] = v4.x, dr[4 * k + 1] = v4.y, dr[4 * k + 2] = v4.z, dr[4 * k + 3] = v4.w;
                }
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int r = 0; r < 16; r += 2) {
                    a0 = fmaf(dr[r], __shfl_sync(kFull, w, r), a0);  // rows r < c of the column are zero
                    a1 = fmaf(dr[r + 1], __shfl_sync(kFull, w, r + 1), a1);
                }
                __syncwarp();
                if (lane < 16) { al[c0 + c] = a0 + a1; }
                __syncwarp();
            }
        }

        // -DERL_GP_TC_TIMING: per-role cycle counters (threads 96 / 128 / 256 / 288 / 320: last training warp, first query warp,
        // pivot, MMA, back-substitution warp), summed over the GPs of a CTA and printed by CTA 7 (kernel experiments only)
#ifdef ERL_GP_TC_WATCHDOG_PRINT
#define ERL_GP_TC_AT(j_, stage_) { if (lane == 0) { reinterpret_cast<volatile int *>(smem_raw + kOffDbg)[warp] = (g << 8) | ((j_) << 4) | (stage_); } }
#else
#define ERL_GP_TC_AT(j_, stage_)
#endif
#ifdef ERL_GP_TC_TIMING
#define ERL_GP_TC_TICK(k_) { const long long now_ = clock64(); tm_acc[k_] += now_ - tm_t; tm_t = now_; }
#else
#define ERL_GP_TC_TICK(k_)
#endif

        template<int XDIM>
        __global__ void __launch_bounds__(kThreads, 2)
        RowGpTcKernel(const BatchParams<float> p) {
            extern __shared__ __align__(1024) unsigned char smem_raw[];
            float *smem = reinterpret_cast<float *>(smem_raw);
            float *lp = smem + Lay::kL;
            float *rs = smem + Lay::kRs;
            float *al = smem + Lay::kAl;
            float *dinv = smem + Lay::kDinv;
            float *band = reinterpret_cast<float *>(smem_raw + kOffBand);
            float *lbs = reinterpret_cast<float *>(smem_raw + kOffLb);
            float *ltile = reinterpret_cast<float *>(smem_raw + kOffLt);
            uint32_t *slot = reinterpret_cast<uint32_t *>(smem_raw + kOffSlot);
            int *s_fail = reinterpret_cast<int *>(smem_raw + kOffSlot + 4);
            const uint32_t bars = SmemAddr(smem_raw + kOffBars);
            const uint32_t s_b = SmemAddr(smem_raw + kOffB), s_aqlo = SmemAddr(smem_raw + kOffAqLo);

            const int tid = threadIdx.x;
            const int warp = __shfl_sync(kFull, tid >> 5, 0);  // warp-uniform by construction
            const int lane = tid & 31;
            const bool is_t = warp < 4;
            const bool is_q = warp >= 4 && warp < 8;

            if (tid == 0) {
                // One arrival per WARP (__syncwarp, then one lane arrives): an mbarrier arrival is a shared-memory atomic on one
                // word, and 300 serialised arrivals per panel (one per thread) kept the shared-memory pipe busy - the pipe the
                // shuffles of the pivot chain go through.
                MbarInit(bars + 8 * kBarBand, 1);
                MbarInit(bars + 8 * (kBarBand + 1), 1);
                for (int i = 0; i < 4; ++i) { MbarInit(bars + 8 * (kBarDinv + i), 1); }
                MbarInit(bars + 8 * kBarOpT, 4);
                MbarInit(bars + 8 * kBarOpQ, 4);
                MbarInit(bars + 8 * kBarMmaT, 1);
                MbarInit(bars + 8 * kBarMmaQ, 1);
                MbarInit(bars + 8 * kBarFact, 5);
                MbarInit(bars + 8 * kBarEpi, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            if (warp == 8) {
                asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemAddr(slot)), "n"(kTmemCols) : "memory");
                asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
            }
            FenceBefore();
            __syncthreads();
            FenceAfter();
            const uint32_t tmem = *slot;
            const int grp = is_q ? 1 : 0;
            const uint32_t my_tmem = tmem + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + 128u * grp;  // this warp's lanes, this group's region
            const int row = tid & 127;  // training row (group T) or query slot (group Q)
            const rowgp::CovCoef cov(p.cov);
            // sequence numbers (identical in every thread): trained GPs, panels and updates so far -> mbarrier phase parities
            uint32_t tg = 0, pan_base = 0, upd_base = 0, updq_base = 0;  // (updq: updates of GPs that have queries - the query group's barriers)
            if (warp == 9 && lane == 0) { MbarArrive(bars + 8 * kBarEpi); }  // phase 0: "the epilogue before the first GP" is done
#ifdef ERL_GP_TC_TIMING
            long long tm_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            long long tm_t = clock64();
            const long long tm_start = tm_t;
#endif

            // Inputs are fetched ahead of their use (the loads fly during the previous factorisation): the sizes and query
            // offsets of a GP two GPs ahead, its rows one GP ahead.  One copy of each fetch in the loop (instruction cache).
            int n_cur = 0, n_nx = 0;
            long q0_cur = 0, q1_cur = 0, q0_nx = 0, q1_nx = 0;
            float pf[XDIM + 2];  // training row: point, y, noise variance; query row: point
#pragma unroll
            for (int d = 0; d < XDIM + 2; ++d) { pf[d] = 0.f; }
            const int g0 = blockIdx.x;
            const int stride = static_cast<int>(gridDim.x);
            if (g0 < p.num_gps) {
                n_nx = p.n_train[g0];
                q0_nx = p.q_offsets[g0];
                q1_nx = p.q_offsets[g0 + 1];
            }

            // iteration -1 only fetches (the rows of the first GP); iteration i >= 0 works on GP g0 + i * stride
            for (int g = g0 - stride; g < p.num_gps; g += stride) {
                const int n = n_cur;
                const long q0 = q0_cur, q1 = q1_cur;
                const int g_nx = g + stride;
                // rotate the fetched sizes, fetch the sizes of GP g + 2 * stride
                n_cur = n_nx, q0_cur = q0_nx, q1_cur = q1_nx;
                if (g_nx + stride < p.num_gps) {
                    n_nx = p.n_train[g_nx + stride];
                    q0_nx = p.q_offsets[g_nx + stride];
                    q1_nx = p.q_offsets[g_nx + stride + 1];
                }
                // the rows of this GP (fetched during the previous iteration), then fetch the rows of GP g_nx
                float row_in[XDIM + 2];
#pragma unroll
                for (int d = 0; d < XDIM + 2; ++d) { row_in[d] = pf[d]; }
                if (g_nx < p.num_gps) {
                    if (is_t && row < p.max_n) {
                        const float *gx = p.x + (static_cast<long>(g_nx) * p.max_n + row) * XDIM;
#pragma unroll
                        for (int d = 0; d < XDIM; ++d) { pf[d] = gx[d]; }
                        pf[XDIM] = p.y[static_cast<long>(g_nx) * p.max_n + row];
                        pf[XDIM + 1] = p.var[static_cast<long>(g_nx) * p.max_n + row];
                    } else if (is_q) {
                        const long nqn = q1_cur - q0_cur < 128 ? q1_cur - q0_cur : 128;
                        if (nqn > 0) {
                            const long qi = row < nqn ? row : nqn - 1;  // spare query rows recompute the last query (no stores)
                            const float *gq = p.q_x + (q0_cur + qi) * XDIM;
#pragma unroll
                            for (int d = 0; d < XDIM; ++d) { pf[d] = gq[d]; }
                        }
                    }
                }
                if (g < 0) { continue; }
                if (n <= p.min_train || n <= 0) {  // the reference's `cnt > min_num_samples_per_group` / `cnt > 0` gate
                    if (tid == 0) { p.info[g] = -1; }
                    if (p.valid != nullptr) {
                        for (long q = q0 + tid; q < q1; q += kThreads) { p.valid[q] = 0; }
                    }
                    continue;
                }
#ifdef ERL_GP_TC_TRACE
                if (blockIdx.x == ERL_GP_TC_TRACE && lane == 0) { printf("trace: warp %d GP %d n %d q %ld..%ld tg %u\n", warp, g, n, q0, q1, tg); }
#endif
                const int nblk = (n + 15) >> 4;
                const int npr = nblk * 16;
                const int nq = static_cast<int>(q1 - q0 < 128 ? q1 - q0 : 128);  // queries that ride along with the factorisation
                const bool has_q = nq > 0;
                const uint32_t par_g = tg & 1;
                float *soa_w = reinterpret_cast<float *>(smem_raw + kOffSoa) + par_g * 512;
                const float2 *soa = reinterpret_cast<const float2 *>(soa_w);

                if (warp < 8) {
                    // ================= row threads =================
                    float negp[XDIM];
                    float yacc = 0.f, diag = 1.0f;
                    if (is_t) {
                        const bool in_n = row < n;
#pragma unroll
                        for (int d = 0; d < XDIM; ++d) {
                            const float xv = in_n ? row_in[d] : 0.f;
                            soa_w[d * 128 + row] = xv;
                            negp[d] = -xv;
                        }
                        yacc = in_n ? row_in[XDIM] : 0.f;
                        diag = in_n ? 1.0f + row_in[XDIM + 1] : 1.0f;
                    } else {
#pragma unroll
                        for (int d = 0; d < XDIM; ++d) { negp[d] = -row_in[d]; }
                    }
                    ERL_GP_TC_AT(0, 1)
                    NamedBarrier(2, kRowThreads);  // the staged points of this GP are visible to all row threads
                    ERL_GP_TC_AT(0, 2)
                    ERL_GP_TC_TICK(0)              // staging
                    float mean = 0.f, ss = 0.f;
                    float *gl = p.l + static_cast<long>(g) * p.max_n * p.max_n;
                    const bool store_l = is_t && p.write_l != 0 && row < n;
                    const int w_last = (npr - 1) >> 5;  // the training warp that is live in every panel

                    if (is_t || has_q) {
                        for (int j = 0; j < nblk; ++j) {
                            const int c0 = 16 * j;
                            const bool last = j == nblk - 1;
                            // a training warp is live while it has rows in or below the tile (and inside the padded size); the others
                            // only write their zeros of the strict upper triangle of L
                            const bool warp_live = is_t ? (32 * warp + 31 >= c0 && 32 * warp < npr) : true;
                            const bool below = is_t ? (row >= c0 + 16 && row < npr) : true;
                            const bool band_warp = is_t && !last && warp == ((c0 + 16) >> 5);
                            const bool band_row = band_warp && row >= c0 + 16 && row < c0 + 32;
                            float v[16];
#pragma unroll
                            for (int c = 0; c < 16; ++c) { v[c] = 0.f; }
                            // A wait names its phase by parity only, so every training warp has to see each phase of the "previous
                            // back-substitution done" barrier: the live warps wait for it before their first store to L below, a warp
                            // whose rows lie beyond the padded size does it here (it would otherwise reach the wait for THIS GP's
                            // back-substitution, further query tiles, one phase early - or two phases late).
                            if (j == 0 && is_t && !warp_live) { MbarWait(bars + 8 * kBarEpi, tg & 1); }
                            if (warp_live) {
                                // (a) the entries of this panel: Gram (training rows) / Ktest (query rows), minus the accumulated
                                //     updates (the tensor core subtracted them from zero).  Look-ahead: the warp that holds the rows of
                                //     tile j + 1 first does the same for the NEXT panel's columns (D_{j+1}) - one rolled loop.
                                float x[16];
#pragma unroll
                                for (int c = 0; c < 16; ++c) { x[c] = 0.f; }
                                bool waited = false;
#pragma unroll 1
                                for (int pass = band_warp ? 1 : 0; pass >= 0; --pass) {
                                    const int cb = c0 + 16 * pass;
                                    const bool need_mask = n != npr || (is_t && 32 * warp <= cb + 15 && 32 * warp + 31 >= cb);
                                    PanelEntries<XDIM>(cov, soa, cb, negp, is_t, row, n, diag, need_mask, x);  // (before the wait: it fills it)
                                    if (j > 0) {
                                        if (!waited) {
                                            ERL_GP_TC_TICK(1)  // entries
                                            ERL_GP_TC_AT(j, 3)
                                            if (is_t) {
                                                MbarWait(bars + 8 * kBarMmaT, (upd_base + j - 1) & 1);
                                            } else {
                                                MbarWait<ERL_GP_TC_SLEEP_NS>(bars + 8 * kBarMmaQ, (updq_base + j - 1) & 1);
                                            }
                                            FenceAfter();
                                            waited = true;
                                            ERL_GP_TC_TICK(8)  // wait for the MMAs
                                        }
                                        float d[16];
                                        TmemLd16(my_tmem + cb, d);
#pragma unroll
                                        for (int c = 0; c < 16; ++c) { x[c] += d[c]; }
                                    }
                                    // the tile's own columns go to the hand-over buffer: tile 0 is complete (Gram entries), tile j + 1 still
                                    // lacks the update of panel j, which its warp applies below once Dinv_j is there
                                    if (pass == 1 ? band_row : (j == 0 && is_t && row < 16)) {
                                        float *br = band + ((j + pass) & 1) * 16 * kBandLd + (row & 15) * kBandLd + 16;
#pragma unroll
                                        for (int k4 = 0; k4 < 4; ++k4) { *reinterpret_cast<float4 *>(br + 4 * k4) = make_float4(x[4 * k4], x[4 * k4 + 1], x[4 * k4 + 2], x[4 * k4 + 3]); }
                                        if (pass == 0) { br[16] = yacc; }
                                    }
                                }
                                if (j == 0 && is_t && warp == 0) {
                                    __syncwarp();
                                    if (lane == 0) { MbarArrive(bars + 8 * (kBarBand + (pan_base & 1))); }  // tile 0
                                }
                                ERL_GP_TC_TICK(2)  // entries, tcgen05.ld, band hand-over
                                // (b) Dinv_j and z_j
#ifdef ERL_GP_TC_DEFER_Q
                                if (!is_t && !last) {  // the query rows stay out of the way until the look-ahead hand-over of this panel is done
                                    MbarWait(bars + 8 * (kBarBand + ((pan_base + j + 1) & 1)), ((pan_base + j + 1) >> 1) & 1);
                                }
#endif
                                ERL_GP_TC_AT(j, 4)
                                MbarWait(bars + 8 * (kBarDinv + ((pan_base + j) & 3)), ((pan_base + j) >> 2) & 1);
                                ERL_GP_TC_TICK(3)  // wait for the pivot warp
                                // (c) v = x Dinv_j^T, then the z products; the rows of the tile take their row of L_jj from the pivot warp
                                const bool in_tile = is_t && row >= c0 && row < c0 + 16;
                                if (below) {
#ifdef ERL_GP_TC_EXPERIMENT_NO_QV  // timing experiment only (wrong results): the query rows skip the Dinv product
                                    if (is_t) {
                                        RowTimesDinvT(x, dinv + j * 16 * Lay::kDinvLd, v);
                                    } else {
#pragma unroll
                                        for (int c = 0; c < 16; ++c) { v[c] = x[c]; }
                                    }
#else
                                    RowTimesDinvT(x, dinv + j * 16 * Lay::kDinvLd, v);
#endif
                                    const float dz = Dot16(v, al + c0);
                                    if (is_t) {
                                        yacc -= dz;
                                    } else {
                                        float sq = 0.f;
#pragma unroll
                                        for (int c = 0; c < 16; ++c) { sq = fmaf(v[c], v[c], sq); }
                                        mean += dz;
                                        ss += sq;
                                    }
                                } else if (in_tile) {
                                    const float *lt = ltile + (j & 1) * 16 * kLbLd + (row - c0) * kLbLd;
#pragma unroll
                                    for (int k4 = 0; k4 < 4; ++k4) {
                                        const float4 t4 = *reinterpret_cast<const float4 *>(lt + 4 * k4);
                                        v[4 * k4] = t4.x, v[4 * k4 + 1] = t4.y, v[4 * k4 + 2] = t4.z, v[4 * k4 + 3] = t4.w;
                                    }
#pragma unroll
                                    for (int c = 0; c < 16; ++c) {
                                        if (c > row - c0) { v[c] = 0.f; }
                                    }
                                }
                                // (c') LOOK-AHEAD, on the critical path of the GP: the warp that holds the rows of tile j + 1 applies the
                                //      update of panel j to that tile itself (D_{j+1} -= L L^T, L = L_{j+1,j} = the V rows just computed:
                                //      16^3 multiply-adds over the 32 lanes, L through shared memory k-major) and hands the tile to the pivot
                                //      warp, while the tensor core applies the update to everything else
                                if (band_warp) {
                                    float *bb = band + ((j + 1) & 1) * 16 * kBandLd;
                                    if (band_row) {
#pragma unroll
                                        for (int k = 0; k < 16; ++k) { lbs[k * kLbLd + (row & 15)] = v[k]; }
                                        bb[(row & 15) * kBandLd + 32] = yacc;
                                    }
                                    __syncwarp();
                                    {
                                        const int rr = lane & 15, cb = 8 * (lane >> 4);
                                        float *dp = bb + rr * kBandLd + 16 + cb;
                                        const float4 da = *reinterpret_cast<const float4 *>(dp), db = *reinterpret_cast<const float4 *>(dp + 4);
                                        float2 d2[4] = {make_float2(da.x, da.y), make_float2(da.z, da.w), make_float2(db.x, db.y), make_float2(db.z, db.w)};
#pragma unroll 4
                                        for (int k = 0; k < 16; ++k) {
                                            const float mine = -lbs[k * kLbLd + rr];
                                            const float4 ca = *reinterpret_cast<const float4 *>(lbs + k * kLbLd + cb), cc = *reinterpret_cast<const float4 *>(lbs + k * kLbLd + cb + 4);
                                            d2[0] = rowgp::Fma2(make_float2(ca.x, ca.y), mine, d2[0]);
                                            d2[1] = rowgp::Fma2(make_float2(ca.z, ca.w), mine, d2[1]);
                                            d2[2] = rowgp::Fma2(make_float2(cc.x, cc.y), mine, d2[2]);
                                            d2[3] = rowgp::Fma2(make_float2(cc.z, cc.w), mine, d2[3]);
                                        }
                                        *reinterpret_cast<float4 *>(dp) = make_float4(d2[0].x, d2[0].y, d2[1].x, d2[1].y);
                                        *reinterpret_cast<float4 *>(dp + 4) = make_float4(d2[2].x, d2[2].y, d2[3].x, d2[3].y);
                                    }
                                    __syncwarp();
                                    if (lane == 0) { MbarArrive(bars + 8 * (kBarBand + ((pan_base + j + 1) & 1))); }
                                    ERL_GP_TC_TICK(7)  // look-ahead update (band warp)
                                }
                                // (d) operands of the trailing update - before the stores of L: the tensor core waits for them (TMEM columns
                                //     and operand buffers are free: this group's update j - 1 was awaited above, and its commit followed the
                                //     other group's update j - 2)
                                if (!last) {
                                    uint32_t hi[16], lo[16];
#pragma unroll
                                    for (int c = 0; c < 16; ++c) {
                                        hi[c] = __float_as_uint(v[c]);
                                        lo[c] = rowgp::Tf32Lo(v[c]);
                                    }
                                    TmemSt16(my_tmem + c0, hi);                     // X_j's own columns are dead: V_hi
                                    if (j > 0) { TmemSt16(my_tmem + c0 - 16, lo); }  // V_{j-1} hi is dead as well: V_lo
                                    if (is_t) {
                                        const uint32_t sb = s_b + static_cast<uint32_t>(j & 1) * 2u * kOperandBytes;
                                        StoreOperandRow(sb, row, hi);
                                        StoreOperandRow(sb + kOperandBytes, row, lo);
                                    } else if (j == 0) {
                                        StoreOperandRow(s_aqlo, row, lo);
                                    }
                                    TmemStWait();
                                    FenceProxyAsync();
                                    FenceBefore();
                                    if (is_t) {
                                        // dead warps (rows above the panel, or beyond the padded size) do not come here: the warp
                                        // that is live in every panel arrives for them
                                        int n_live = 0;
#pragma unroll
                                        for (int w = 0; w < 4; ++w) { n_live += (32 * w + 31 >= c0 && 32 * w < npr) ? 1 : 0; }
                                        __syncwarp();
                                        if (lane == 0) { MbarArriveCount(bars + 8 * kBarOpT, warp == w_last ? static_cast<uint32_t>(5 - n_live) : 1u); }
                                    } else {
                                        __syncwarp();
                                        if (lane == 0) { MbarArrive(bars + 8 * kBarOpQ); }
                                    }
                                    ERL_GP_TC_TICK(5)  // operand split and stores
                                }
                                // (d') the shared-memory copy of L (back-substitution, further query tiles)
                                if (is_t && (below || in_tile)) {
                                    ERL_GP_TC_AT(j, 5)
                                    if (j == 0) { MbarWait(bars + 8 * kBarEpi, tg & 1); }  // the back-substitution of the previous GP has read L
                                    ERL_GP_TC_AT(j, 6)
                                    float *lcol = lp + Lay::Base(j) + (row - c0);
#pragma unroll
                                    for (int c = 0; c < 16; ++c) { lcol[c * Lay::Stride(j)] = v[c]; }
                                }
                            }
                            // (e) L to HBM straight from the registers: one 128-byte line per warp and column; the rows above the tile
                            //     write the zeros of the strict upper triangle (the rows of the tile are written by the pivot warp)
                            if (store_l) {  // (every row of the matrix: below the tile V, in the tile L_jj, above it zeros)
                                float *gp = gl + row + c0 * p.max_n;
                                if (p.max_n == 128) {  // immediate offsets: one instruction per column
#pragma unroll
                                    for (int c = 0; c < 16; ++c) {
                                        if (c0 + c < n) { gp[c * 128] = v[c]; }
                                    }
                                } else {  // any other capacity: a rolled loop over a rotating copy (code size)
#pragma unroll 1
                                    for (int c = 0; c < 16; ++c) {
                                        if (c0 + c < n) { *gp = v[0]; }
                                        gp += p.max_n;
                                        const float v0 = v[0];
#pragma unroll
                                        for (int i = 0; i < 15; ++i) { v[i] = v[i + 1]; }
                                        v[15] = v0;
                                    }
                                }
                            }
                            ERL_GP_TC_TICK(4)  // V = X Dinv^T, z products, L stores
                        }
                    }
                    ERL_GP_TC_AT(15, 7)
                    if (is_t) {
                        __syncwarp();
                        if (lane == 0) { MbarArrive(bars + 8 * kBarFact); }
                    }
                    // ---- outputs of the first query tile (the fail flag was published before the last Dinv) ----
                    if (is_q) {
                        if (has_q) {
                            const int failed = s_fail[par_g];
                            if (failed != 0) {
                                if (p.valid != nullptr) {
                                    for (long q = q0 + row; q < q1; q += 128) { p.valid[q] = 0; }
                                }
                            } else if (row < nq) {
                                const long dst = q0 + row;
                                if (p.mean != nullptr) { p.mean[dst] = mean; }
                                if (p.variance != nullptr) { p.variance[dst] = 1.0f - ss; }  // literal prior 1.0f, src/vanilla_gp.cpp:121
                                if (p.valid != nullptr) { p.valid[dst] = 1; }
                            }
                        }
                    }
                    ERL_GP_TC_TICK(6)  // outputs
                    // ---- query tiles beyond the first 128: the mma.sync predict of erl_gp_rowgp.cuh on its shared-memory layout ----
                    if (is_t && q1 - q0 > 128) {
                        ERL_GP_TC_AT(15, 8)
                        // alpha of THIS GP is in shared memory (every training warp has seen the previous back-substitution complete
                        // during panel 0, so the parity names the right phase)
                        MbarWait(bars + 8 * kBarEpi, (tg + 1) & 1);
                        ERL_GP_TC_AT(15, 9)
                        if (s_fail[par_g] == 0) {
                            float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
                            pt.x = -negp[0];
                            if (XDIM > 1) { pt.y = -negp[XDIM > 1 ? 1 : 0]; }
                            if (XDIM > 2) { pt.z = -negp[XDIM > 2 ? 2 : 0]; }
                            pt.w = row < n ? smem[Lay::kVar + row] : 0.f;
                            reinterpret_cast<float4 *>(smem + Lay::kPts)[row] = pt;
                            smem[Lay::kSoa + row] = pt.x;
                            smem[Lay::kSoa + Lay::kNp + row] = pt.y;
                            smem[Lay::kSoa + 2 * Lay::kNp + row] = pt.z;
                            smem[Lay::kSoa + 3 * Lay::kNp + row] = pt.w;
                            NamedBarrier(1, 128);
                            for (long qb = q0 + 128; qb < q1; qb += 64) {
                                const int nq2 = static_cast<int>(q1 - qb < 64 ? q1 - qb : 64);
                                rowgp::PredictTileMma<XDIM, 8, false>(p, cov, smem, n, nblk, qb, nq2);
                            }
                        }
                        
                    }
                } else if (warp == 8) {
                    // ================= pivot warp =================
                    // Nothing but the serial chain: fetch the 16 rows of the tile, eliminate, leave L_jj / Dinv_j / z_j in shared
                    // memory.  (A lone warp retires about one instruction every five cycles here: whatever else used to be done by
                    // this warp - the strided and global stores of L_jj, the look-ahead update - sat on the critical path of the GP.)
                    const int r = lane & 15;
                    int fail = 0;
                    MbarWait(bars + 8 * kBarEpi, tg & 1);  // the back-substitution of the previous GP is done with z, Dinv and L
                    for (int j = 0; j < nblk; ++j) {
                        const int c0 = 16 * j;
                        MbarWait(bars + 8 * (kBarBand + ((pan_base + j) & 1)), ((pan_base + j) >> 1) & 1);
                        ERL_GP_TC_TICK(1)  // pivot warp: wait for the tile
                        float prow[16], l[16];
                        const float *br = band + (j & 1) * 16 * kBandLd + r * kBandLd;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const float4 t4 = *reinterpret_cast<const float4 *>(br + 16 + 4 * k4);
                            prow[4 * k4] = t4.x, prow[4 * k4 + 1] = t4.y, prow[4 * k4 + 2] = t4.z, prow[4 * k4 + 3] = t4.w;
                        }
                        float zacc = br[32];
                        if (lane >= 16) {
#pragma unroll
                            for (int c = 0; c < 16; ++c) { prow[c] = c == r ? 1.0f : 0.f; }  // the unit vectors: Dinv_j for free
                        }
                        rowgp::PivotBlock<2>(prow, zacc, l, 0, c0, lane, fail, rs, al);
                        // lanes 0-15: row r of L_jj (the tile's row threads mask the upper triangle and store it); lanes 16-31: column r of Dinv_j
                        float *dst = lane < 16 ? ltile + (j & 1) * 16 * kLbLd + r * kLbLd : dinv + j * 16 * Lay::kDinvLd + r * Lay::kDinvLd;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) { *reinterpret_cast<float4 *>(dst + 4 * k4) = make_float4(l[4 * k4], l[4 * k4 + 1], l[4 * k4 + 2], l[4 * k4 + 3]); }
                        if (j == nblk - 1 && lane == 0) { s_fail[par_g] = fail; }
                        __syncwarp();
                        // The pivot warp's "factorisation done" arrival comes BEFORE the last Dinv is published: no training warp can
                        // leave this GP (and, with a short next GP, arrive for the NEXT phase of that barrier) ahead of it.  The phase
                        // still completes only after the warp that waits for this Dinv has stored its rows and arrived.
                        if (j == nblk - 1 && lane == 0) { MbarArrive(bars + 8 * kBarFact); }
                        if (lane == 0) { MbarArrive(bars + 8 * (kBarDinv + ((pan_base + j) & 3))); }
                        ERL_GP_TC_TICK(2)  // pivot warp: pivot tile
                    }
                } else if (warp == 9) {
                    // ================= MMA warp =================
                    const uint32_t elected = ElectOne();
                    for (int j = 0; j + 1 < nblk; ++j) {
                        const int c0 = 16 * j;
                        const int col0 = c0 + 16;
                        const uint32_t idesc = InstrDesc(npr - col0);
                        const uint32_t brow_off = static_cast<uint32_t>(col0 >> 3) * 128u;  // B = rows col0 .. of L_j
                        const uint32_t sb = s_b + static_cast<uint32_t>(j & 1) * 2u * kOperandBytes;
                        for (int gq = 0; gq < (has_q ? 2 : 1); ++gq) {
                            MbarWait(bars + 8 * (gq == 0 ? kBarOpT : kBarOpQ), ((gq == 0 ? upd_base : updq_base) + j) & 1);
                            FenceAfter();
                            ERL_GP_TC_TICK(1 + 2 * gq)  // MMA warp: wait for the operands
                            const uint32_t dcol = tmem + 128u * gq + col0;
                            const uint32_t a_hi = tmem + 128u * gq + c0;
                            const uint32_t a_lo = a_hi - 16;
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                const uint64_t bhi = SmemDesc(sb + brow_off + ks * 2 * kChunkBytes);
                                const uint64_t blo = SmemDesc(sb + kOperandBytes + brow_off + ks * 2 * kChunkBytes);
                                const uint32_t acc0 = (j > 0 || ks > 0) ? 1u : 0u;
                                if (j == 0) {
                                    MmaSS(elected, dcol, SmemDesc((gq == 0 ? sb + kOperandBytes : s_aqlo) + ks * 2 * kChunkBytes), bhi, idesc, acc0);
                                } else {
                                    MmaTS(elected, dcol, a_lo + 8 * ks, bhi, idesc, acc0);
                                }
                                MmaTS(elected, dcol, a_hi + 8 * ks, blo, idesc, 1u);
                                MmaTS(elected, dcol, a_hi + 8 * ks, bhi, idesc, 1u);
                            }
                            Commit(elected, bars + 8 * (gq == 0 ? kBarMmaT : kBarMmaQ));
                            __syncwarp();
                            ERL_GP_TC_TICK(2 + 2 * gq)  // MMA warp: issue
                        }
                    }
                    // ---- alpha = L^-T z of this GP: the MMA warp has nothing to issue until the next GP's first V is ready (about one
                    //      pivot tile + one V after its start), which is the time the back-substitution takes ----
                    MbarWait<ERL_GP_TC_SLEEP_NS>(bars + 8 * kBarFact, tg & 1);
                    ERL_GP_TC_TICK(5)  // MMA warp: wait for the end of the factorisation
                    const int failed = s_fail[par_g];
                    if (failed == 0) {
                        BackSolveWarp(smem, nblk, lane);
                        for (int i = lane; i < n; i += 32) { p.alpha[static_cast<long>(g) * p.max_n + i] = smem[Lay::kVar + i]; }
                    }
                    if (lane == 0) { p.info[g] = failed; }
                    __syncwarp();
                    if (lane == 0) { MbarArrive(bars + 8 * kBarEpi); }
                    ERL_GP_TC_TICK(6)  // MMA warp: back-substitution
                }
                ERL_GP_TC_AT(15, 10)
                tg += 1;
                pan_base += static_cast<uint32_t>(nblk);
                upd_base += static_cast<uint32_t>(nblk - 1);
                if (has_q) { updq_base += static_cast<uint32_t>(nblk - 1); }
            }
#ifdef ERL_GP_TC_TIMING
            if (blockIdx.x == 7 && (tid == 96 || tid == 128 || tid == 256 || tid == 288)) {
                printf("tc timing tid %3d total %lld: %lld %lld %lld %lld %lld %lld %lld %lld %lld\n", tid, clock64() - tm_start, tm_acc[0], tm_acc[1], tm_acc[2], tm_acc[3], tm_acc[4], tm_acc[5],
                       tm_acc[6], tm_acc[7], tm_acc[8]);
            }
#endif
#ifdef ERL_GP_TC_TRACE
            if (blockIdx.x == ERL_GP_TC_TRACE && lane == 0) { printf("trace: warp %d leaves the GP loop\n", warp); }
#endif
            FenceBefore();
            __syncthreads();
            if (warp == 8) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory"); }
        }

        template<int XDIM>
        int
        Launch(Context *ctx, const BatchParams<float> &params) {
            auto kernel = RowGpTcKernel<XDIM>;
            if (static_cast<int>(kSmemBytes) > ctx->max_smem_optin) {
                return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "row-GP tensor kernel needs %zu B of shared memory, device allows %d", kSmemBytes, ctx->max_smem_optin);
            }
            ERL_GP_CUDA_OK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBytes)));
            int ctas = 2 * ctx->sm_count;  // persistent: two CTAs (2 x 256 TMEM columns) per SM
            if (ctas > params.num_gps) { ctas = params.num_gps; }
            kernel<<<static_cast<unsigned>(ctas), kThreads, kSmemBytes, ctx->stream>>>(params);
            ctx->launches += 1;
            ERL_GP_CUDA_OK(ctx, cudaGetLastError());
            return ERL_GP_STATUS_OK;
        }

    }  // namespace rowgp_tc
}  // namespace erl_gp
