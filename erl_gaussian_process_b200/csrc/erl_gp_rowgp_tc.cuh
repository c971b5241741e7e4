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
//     read by the MMA from there (no shared-memory round trip: with N <= 112 an A operand in shared memory would cost more
//     shared-memory bandwidth than the whole rest of the kernel), the B operands (rows of L_j, hi and lo) sit in shared
//     memory in the K-major no-swizzle canonical layout, once per GP and panel;
//   * the products are subtracted by the a_negate bit of the instruction descriptor;
//   * the update of the NEXT panel's 16 columns is issued and committed first (mbarrier 0), the rest second (mbarrier 1): the
//     serial chain pivot -> V -> update -> next pivot only ever waits for 12 MMAs with N = 16;
//   * a dedicated ninth warp factorises the 16 x 16 pivot tiles (rowgp::PivotBlock: lanes 0-15 own the rows, lanes 16-31
//     eliminate the unit vectors = Dinv for free, z = L^-1 y rides along) and issues the MMAs; everything is synchronised
//     with mbarriers (tile ready / Dinv ready / operands ready / MMA done), the CTA-wide barrier only frames a GP.
// The Gram matrix and Ktest are never stored: their entries are generated in the thread that owns the row, right when the
// panel is consumed (fused distance + covariance + noise diagonal, rowgp::CovPair).  L (packed FP32), alpha by blocked
// back-substitution through the Dinv blocks, the L write-back and further query tiles (beyond 128 queries per GP) reuse the
// shared-memory layout and the routines of erl_gp_rowgp.cuh.
#pragma once

#include "erl_gp_rowgp.cuh"

namespace erl_gp {
    namespace rowgp_tc {

        using Lay = rowgp::Layout<8>;
        using rowgp::kFull;

        constexpr int kRowThreads = 256;  // warps 0-3: training rows (group T), warps 4-7: query rows (group Q)
        constexpr int kThreads = 288;     // + warp 8: pivot tiles and MMA issue
        constexpr int kTmemCols = 256;    // T region: columns [0, 128), Q region: [128, 256)
        constexpr uint32_t kChunkBytes = 128 * 16;  // distance of the 16-byte K chunks of an operand buffer (LBO)
        constexpr int kOperandBytes = 128 * 16 * 4;  // 128 rows x 16 k x FP32

        // shared memory (bytes): the row-GP layout, then the MMA operand buffers and the small hand-over buffers
        constexpr size_t kOffBhi = (Lay::kBytes + 127) / 128 * 128;  // rows of L_j, hi (B operand; at panel 0 also nobody's A)
        constexpr size_t kOffBlo = kOffBhi + kOperandBytes;          // rows of L_j, lo (B operand; at panel 0 the T group's A_lo)
        constexpr size_t kOffAqLo = kOffBlo + kOperandBytes;         // V_0 lo of the query rows (A operand of panel 0 only)
        constexpr size_t kOffTile = kOffAqLo + kOperandBytes;        // pivot tile, column-major 16 x 16, + 16 running y
        constexpr size_t kOffBars = kOffTile + (16 * 16 + 16) * 4;   // 5 mbarriers
        constexpr size_t kOffSlot = kOffBars + 8 * 8;                // TMEM base address, fail flag
        constexpr size_t kSmemBytes = kOffSlot + 16;
        static_assert(kSmemBytes > 227 * 1024 / 3, "three resident CTAs would need 3 x 256 TMEM columns");

        enum Bar : int { kBarTile = 0, kBarDinv = 1, kBarFull = 2, kBarMma0 = 3, kBarMma1 = 4 };

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

        // try_wait suspends the thread for a hardware-defined time per attempt; a protocol error would otherwise hang the GPU
        // until the watchdog of the box fires, so the number of attempts is bounded and the kernel traps instead
        // (-DERL_GP_TC_NO_WATCHDOG removes the counter).
        __device__ __forceinline__ void
        MbarWait(const uint32_t bar, const uint32_t parity) {
#ifdef ERL_GP_TC_NO_WATCHDOG
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "WAIT_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra DONE_%=;\n\t"
                "bra WAIT_%=;\n\t"
                "DONE_%=:\n\t}\n" ::"r"(bar),
                "r"(parity)
                : "memory");
#else
            uint32_t done = 0;
            for (uint32_t tries = 0; !done; ++tries) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                if (!done && tries > (1u << 22)) { __trap(); }
            }
#endif
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

        // 16 covariance entries k(point, training point c0 + c), c = 0 .. 15, on the packed FP32 pipe (two at a time)
        template<int XDIM>
        __device__ __forceinline__ void
        PanelEntries(const rowgp::CovCoef &cov, const float *__restrict__ smem, const int c0, const float (&negp)[XDIM], float (&e)[16]) {
            const float2 *soa = reinterpret_cast<const float2 *>(smem + Lay::kSoa);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                float2 pc[XDIM];
#pragma unroll
                for (int d = 0; d < XDIM; ++d) { pc[d] = soa[d * (Lay::kNp / 2) + c0 / 2 + m]; }
                const float2 kv = rowgp::CovPair(cov, rowgp::Dist2Pair<XDIM>(pc, negp));
                e[2 * m] = kv.x;
                e[2 * m + 1] = kv.y;
            }
        }

        // alpha = L^-T z for the 128 training-row threads (al: z on entry, alpha on exit); thread = column, blocked from the
        // bottom through the inverses of the diagonal blocks (the algorithm of rowgp::BackSolve<8, true>, with a named barrier)
        __device__ __forceinline__ void
        BackSolveT(float *__restrict__ smem, const int nblk, const int tid, const int warp, const int lane) {
            const float *lp = smem + Lay::kL;
            float *al = smem + Lay::kAl;
            float s = 0.f;
            for (int kb = nblk - 1; kb >= 0; --kb) {
                const int c0 = 16 * kb;
                if (warp == (c0 >> 5)) {
                    const int lb = c0 & 31;
                    const bool mine = lane >= lb && lane < lb + 16;
                    const int jj = mine ? lane - lb : 0;
                    const float vj = al[c0 + jj] - s;
                    const float *dcol = smem + Lay::kDinv + kb * 16 * Lay::kDinvLd + jj * Lay::kDinvLd;  // column jj of Dinv
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
                    if (mine) { al[c0 + jj] = a0 + a1; }
                }
                NamedBarrier(1, 128);
                if (tid < c0) {
                    const int cb = tid >> 4;
                    const float *colp = lp + Lay::Base(cb) + (tid & 15) * Lay::Stride(cb) + (c0 - 16 * cb);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 lv = *reinterpret_cast<const float4 *>(colp + 4 * k);
                        const float4 av = *reinterpret_cast<const float4 *>(al + c0 + 4 * k);
                        s = fmaf(lv.x, av.x, s);
                        s = fmaf(lv.y, av.y, s);
                        s = fmaf(lv.z, av.z, s);
                        s = fmaf(lv.w, av.w, s);
                    }
                }
            }
        }

        template<int XDIM>
        __global__ void __launch_bounds__(kThreads, 2)
        RowGpTcKernel(const BatchParams<float> p) {
            extern __shared__ __align__(1024) unsigned char smem_raw[];
            float *smem = reinterpret_cast<float *>(smem_raw);
            float *lp = smem + Lay::kL;
            float4 *pts = reinterpret_cast<float4 *>(smem + Lay::kPts);
            float *rs = smem + Lay::kRs;
            float *al = smem + Lay::kAl;
            float *sv = smem + Lay::kVar;
            float *dinv = smem + Lay::kDinv;
            float *tile = reinterpret_cast<float *>(smem_raw + kOffTile);
            float *tile_y = tile + 256;
            uint32_t *slot = reinterpret_cast<uint32_t *>(smem_raw + kOffSlot);
            int *s_fail = reinterpret_cast<int *>(smem_raw + kOffSlot + 4);
            const uint32_t bars = SmemAddr(smem_raw + kOffBars);
            const uint32_t s_bhi = SmemAddr(smem_raw + kOffBhi), s_blo = SmemAddr(smem_raw + kOffBlo), s_aqlo = SmemAddr(smem_raw + kOffAqLo);

            const int tid = threadIdx.x;
            const int warp = __shfl_sync(kFull, tid >> 5, 0);  // warp-uniform by construction
            const int lane = tid & 31;
            const bool is_t = warp < 4;
            const bool is_q = warp >= 4 && warp < 8;

            if (tid == 0) {
                MbarInit(bars + 8 * kBarTile, 32);
                MbarInit(bars + 8 * kBarDinv, 32);
                MbarInit(bars + 8 * kBarFull, kRowThreads);
                MbarInit(bars + 8 * kBarMma0, 1);
                MbarInit(bars + 8 * kBarMma1, 1);
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
            uint32_t n_tile = 0, n_dinv = 0, n_full = 0, n_m0 = 0, n_m1 = 0;  // completed phases of each mbarrier, as seen by this thread
            const rowgp::CovCoef cov(p.cov);

            for (int g = blockIdx.x; g < p.num_gps; g += gridDim.x) {
                const int n = p.n_train[g];
                const long q0 = p.q_offsets[g];
                const long q1 = p.q_offsets[g + 1];
                if (n <= p.min_train || n <= 0) {  // the reference's `cnt > min_num_samples_per_group` / `cnt > 0` gate
                    if (tid == 0) { p.info[g] = -1; }
                    if (p.valid != nullptr) {
                        for (long q = q0 + tid; q < q1; q += kThreads) { p.valid[q] = 0; }
                    }
                    continue;
                }
                const int nblk = (n + 15) >> 4;
                const int npr = nblk * 16;
                const int nq = static_cast<int>(q1 - q0 < 128 ? q1 - q0 : 128);  // queries that ride along with the factorisation
                const bool has_q = nq > 0;

                // ---- stage the training inputs (as rowgp::RowGpKernel) ----
                float yacc = 0.f, diag = 1.0f;
                float negp[XDIM];  // minus the point of this row (training point or query)
#pragma unroll
                for (int d = 0; d < XDIM; ++d) { negp[d] = 0.f; }
                if (is_t) {
                    const float *gx = p.x + (static_cast<long>(g) * p.max_n + row) * XDIM;
                    float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
                    float yv = 0.f, vv = 0.f;
                    if (row < n) {
                        pt.x = gx[0];
                        if (XDIM > 1) { pt.y = gx[XDIM > 1 ? 1 : 0]; }
                        if (XDIM > 2) { pt.z = gx[XDIM > 2 ? 2 : 0]; }
                        yv = p.y[static_cast<long>(g) * p.max_n + row];
                        vv = p.var[static_cast<long>(g) * p.max_n + row];
                    }
                    pts[row] = pt;
                    rs[row] = 1.0f;
                    smem[Lay::kSoa + row] = pt.x;
                    smem[Lay::kSoa + Lay::kNp + row] = pt.y;
                    smem[Lay::kSoa + 2 * Lay::kNp + row] = pt.z;
                    smem[Lay::kSoa + 3 * Lay::kNp + row] = 0.f;
                    al[row] = yv;
                    sv[row] = vv;
                    yacc = yv;
                    diag = row < n ? 1.0f + vv : 1.0f;
                    negp[0] = -pt.x;
                    if (XDIM > 1) { negp[XDIM > 1 ? 1 : 0] = -pt.y; }
                    if (XDIM > 2) { negp[XDIM > 2 ? 2 : 0] = -pt.z; }
                } else if (is_q && has_q) {
                    const int qi = row < nq ? row : nq - 1;  // spare query rows recompute the last query (no stores)
                    const float *gq = p.q_x + (q0 + qi) * XDIM;
#pragma unroll
                    for (int d = 0; d < XDIM; ++d) { negp[d] = -gq[d]; }
                }
                float mean = 0.f, ss = 0.f;
                int fail = 0;
                __syncthreads();

                for (int j = 0; j < nblk; ++j) {
                    const int c0 = 16 * j;
                    const bool last = j == nblk - 1;
                    if (warp < 8) {
                        // ================= row threads =================
                        const bool warp_live = is_t ? (32 * warp + 31 >= c0 && 32 * warp < npr) : has_q;
                        const bool in_tile = is_t && row >= c0 && row < c0 + 16;
                        const bool below = is_t ? (row >= c0 + 16 && row < npr) : true;
                        float x[16];
                        if (warp_live) {
                            // (a) the entries of this panel: Gram (training rows) / Ktest (query rows)
                            PanelEntries<XDIM>(cov, smem, c0, negp, x);
                            if (is_t) {
#pragma unroll
                                for (int c = 0; c < 16; ++c) {
                                    const int col = c0 + c;
                                    if (row >= n || col >= n) { x[c] = 0.f; }
                                    if (row == col) { x[c] = diag; }
                                }
                            } else if (c0 + 16 > n) {
#pragma unroll
                                for (int c = 0; c < 16; ++c) {
                                    if (c0 + c >= n) { x[c] = 0.f; }
                                }
                            }
                            // (b) minus the accumulated updates (the tensor core subtracted them from zero)
                            if (j > 0) {
                                if (is_t) {
                                    MbarWait(bars + 8 * kBarMma0, n_m0 & 1);
                                } else {
                                    MbarWait(bars + 8 * kBarMma1, n_m1 & 1);
                                }
                                FenceAfter();
                                float d[16];
                                TmemLd16(my_tmem + c0, d);
#pragma unroll
                                for (int c = 0; c < 16; ++c) { x[c] += d[c]; }
                            }
                            // (c) pivot rows -> the pivot warp
                            if (in_tile) {
#pragma unroll
                                for (int c = 0; c < 16; ++c) { tile[c * 16 + (row - c0)] = x[c]; }
                                tile_y[row - c0] = yacc;
                            }
                        }
                        if (j > 0) {
                            n_m0 += 1;
                            if (!is_t) { n_m1 += 1; }  // the query rows are done with this phase (waited above, or not needed: no queries)
                        }
                        if (is_t && warp == (c0 >> 5)) { MbarArrive(bars + 8 * kBarTile); }
                        // (d) Dinv_j and z_j
                        MbarWait(bars + 8 * kBarDinv, n_dinv & 1);
                        n_dinv += 1;
                        float v[16];
#pragma unroll
                        for (int c = 0; c < 16; ++c) { v[c] = 0.f; }
                        if (warp_live) {
                            // (e) v = x Dinv_j^T (Dinv column-major, zero above the diagonal), then the z products
                            const float *dj = dinv + j * 16 * Lay::kDinvLd;
#pragma unroll
                            for (int k = 0; k < 16; ++k) {
#pragma unroll
                                for (int c4 = k / 4; c4 < 4; ++c4) {
                                    const float4 w = *reinterpret_cast<const float4 *>(dj + k * Lay::kDinvLd + 4 * c4);  // Dinv[4 c4 .. 4 c4 + 3][k]
                                    v[4 * c4] = fmaf(x[k], w.x, v[4 * c4]);
                                    v[4 * c4 + 1] = fmaf(x[k], w.y, v[4 * c4 + 1]);
                                    v[4 * c4 + 2] = fmaf(x[k], w.z, v[4 * c4 + 2]);
                                    v[4 * c4 + 3] = fmaf(x[k], w.w, v[4 * c4 + 3]);
                                }
                            }
                            float dz = 0.f, sq = 0.f;
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                const float4 z = *reinterpret_cast<const float4 *>(al + c0 + 4 * k4);
                                dz = fmaf(v[4 * k4], z.x, dz);
                                dz = fmaf(v[4 * k4 + 1], z.y, dz);
                                dz = fmaf(v[4 * k4 + 2], z.z, dz);
                                dz = fmaf(v[4 * k4 + 3], z.w, dz);
                            }
#pragma unroll
                            for (int c = 0; c < 16; ++c) { sq = fmaf(v[c], v[c], sq); }
                            if (is_t) {
                                if (below) {
                                    yacc -= dz;
                                    float *lcol = lp + Lay::Base(j) + (row - c0);
#pragma unroll
                                    for (int c = 0; c < 16; ++c) { lcol[c * Lay::Stride(j)] = v[c]; }
                                }
                            } else {
                                mean += dz;
                                ss += sq;
                            }
                        }
                        // (f) operands of the trailing update
                        if (!last) {
                            if (j > 0 && is_t) {
                                MbarWait(bars + 8 * kBarMma1, n_m1 & 1);  // the previous update no longer reads the buffers / TMEM columns
                                n_m1 += 1;
                                FenceAfter();
                            }
                            if (warp_live) {
                                uint32_t hi[16], lo[16];
#pragma unroll
                                for (int c = 0; c < 16; ++c) {
                                    hi[c] = __float_as_uint(v[c]);
                                    lo[c] = rowgp::Tf32Lo(v[c]);
                                }
                                TmemSt16(my_tmem + c0, hi);                     // X_j's own columns are dead: V_hi
                                if (j > 0) { TmemSt16(my_tmem + c0 - 16, lo); }  // V_{j-1} hi is dead as well: V_lo
                                if (is_t) {
                                    if (below) {
#pragma unroll
                                        for (int kc = 0; kc < 4; ++kc) {
                                            const uint32_t off = OperandChunk(row, kc);
                                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s_bhi + off), "r"(hi[4 * kc]), "r"(hi[4 * kc + 1]), "r"(hi[4 * kc + 2]), "r"(hi[4 * kc + 3]) : "memory");
                                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s_blo + off), "r"(lo[4 * kc]), "r"(lo[4 * kc + 1]), "r"(lo[4 * kc + 2]), "r"(lo[4 * kc + 3]) : "memory");
                                        }
                                    }
                                } else if (j == 0) {
#pragma unroll
                                    for (int kc = 0; kc < 4; ++kc) {
                                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s_aqlo + OperandChunk(row, kc)), "r"(lo[4 * kc]), "r"(lo[4 * kc + 1]), "r"(lo[4 * kc + 2]), "r"(lo[4 * kc + 3]) : "memory");
                                    }
                                }
                                TmemStWait();
                                FenceProxyAsync();
                            }
                            FenceBefore();
                            MbarArrive(bars + 8 * kBarFull);
                        }
                    } else {
                        // ================= pivot / MMA warp =================
                        MbarWait(bars + 8 * kBarTile, n_tile & 1);
                        n_tile += 1;
                        {
                            const int r = lane & 15;
                            float prow[16], l[16];
#pragma unroll
                            for (int c = 0; c < 16; ++c) {
                                const float pv = tile[c * 16 + r];
                                prow[c] = lane < 16 ? pv : (c == r ? 1.0f : 0.f);
                            }
                            float zacc = tile_y[r];
                            __syncwarp();
                            rowgp::PivotBlock<2>(prow, zacc, l, 0, c0, lane, fail, rs, al);
                            if (lane < 16) {
                                float *lcol = lp + Lay::Base(j) + r;  // row r of the diagonal block (zero above the diagonal)
#pragma unroll
                                for (int c = 0; c < 16; ++c) { lcol[c * Lay::Stride(j)] = c > r ? 0.f : l[c]; }
                            } else {
                                float *dst = dinv + j * 16 * Lay::kDinvLd + r * Lay::kDinvLd;  // column r of Dinv
#pragma unroll
                                for (int k4 = 0; k4 < 4; ++k4) { *reinterpret_cast<float4 *>(dst + 4 * k4) = make_float4(l[4 * k4], l[4 * k4 + 1], l[4 * k4 + 2], l[4 * k4 + 3]); }
                            }
                        }
                        __syncwarp();
                        MbarArrive(bars + 8 * kBarDinv);
                        if (!last) {
                            MbarWait(bars + 8 * kBarFull, n_full & 1);
                            n_full += 1;
                            FenceAfter();
                            {
                                // batch 0 (mbarrier Mma0): the training rows' columns of the NEXT panel - all the serial chain waits for;
                                // batch 1 (mbarrier Mma1): the query rows' next-panel columns and everything to the right, both groups
                                const uint32_t elected = ElectOne();
                                const int n_rest = npr - c0 - 32;  // columns after the next panel
#pragma unroll 1
                                for (int batch = 0; batch < 4; ++batch) {  // (T, next), (Q, next), (T, rest), (Q, rest)
                                    const int gq = batch & 1;
                                    const int part = batch >> 1;
                                    const int ncols = part == 0 ? 16 : n_rest;
                                    const int col0 = c0 + 16 + 16 * part;
                                    if (ncols > 0 && (gq == 0 || has_q)) {
                                        const uint32_t idesc = InstrDesc(ncols);
                                        const uint32_t brow_off = static_cast<uint32_t>(col0 >> 3) * 128u;  // B = rows col0 .. of L_j
                                        const uint32_t dcol = tmem + 128u * gq + col0;
                                        const uint32_t a_hi = tmem + 128u * gq + c0;
                                        const uint32_t a_lo = tmem + 128u * gq + c0 - 16;
#pragma unroll
                                        for (int ks = 0; ks < 2; ++ks) {
                                            const uint64_t bhi = SmemDesc(s_bhi + brow_off + ks * 2 * kChunkBytes);
                                            const uint64_t blo = SmemDesc(s_blo + brow_off + ks * 2 * kChunkBytes);
                                            const uint32_t acc0 = (j > 0 || ks > 0) ? 1u : 0u;
                                            if (j == 0) {
                                                MmaSS(elected, dcol, SmemDesc((gq == 0 ? s_blo : s_aqlo) + ks * 2 * kChunkBytes), bhi, idesc, acc0);
                                            } else {
                                                MmaTS(elected, dcol, a_lo + 8 * ks, bhi, idesc, acc0);
                                            }
                                            MmaTS(elected, dcol, a_hi + 8 * ks, blo, idesc, 1u);
                                            MmaTS(elected, dcol, a_hi + 8 * ks, bhi, idesc, 1u);
                                        }
                                    }
                                    if (batch == 0) { Commit(elected, bars + 8 * kBarMma0); }
                                }
                                Commit(elected, bars + 8 * kBarMma1);
                            }
                            __syncwarp();
                        }
                    }
                }
                // the last update's second commit has not been consumed by the row threads yet
                if (is_t && nblk > 1) {  // (the query rows consumed this phase in the last panel; the CTA barrier below covers them otherwise)
                    MbarWait(bars + 8 * kBarMma1, n_m1 & 1);
                    n_m1 += 1;
                }
                if (warp == 8 && lane == 0) { *s_fail = fail; }
                FenceBefore();
                __syncthreads();
                FenceAfter();
                const int failed = *s_fail;
                if (failed != 0) {
                    if (tid == 0) { p.info[g] = failed; }
                    if (p.valid != nullptr) {
                        for (long q = q0 + tid; q < q1; q += kThreads) { p.valid[q] = 0; }
                    }
                    __syncthreads();
                    continue;
                }
                // ---- outputs of the first query tile ----
                if (is_q && row < nq) {
                    const long dst = q0 + row;
                    if (p.mean != nullptr) { p.mean[dst] = mean; }
                    if (p.variance != nullptr) { p.variance[dst] = 1.0f - ss; }  // literal prior 1.0f, src/vanilla_gp.cpp:121
                    if (p.valid != nullptr) { p.valid[dst] = 1; }
                }
                // ---- L write-back (coalesced float4 rows of a column, one column per warp and step) ----
                if (p.write_l && warp < 8) {
                    float *gl = p.l + static_cast<long>(g) * p.max_n * p.max_n;
                    if ((p.max_n & 3) == 0) {
                        const int r4 = 4 * lane;
                        for (int c = warp; c < n; c += 8) {
                            const int cb = c >> 4;
                            if (r4 < n) {
                                float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (r4 >= 16 * cb) { v4 = *reinterpret_cast<const float4 *>(lp + Lay::Base(cb) - 16 * cb + (c & 15) * Lay::Stride(cb) + r4); }
                                float *gcol = gl + static_cast<long>(c) * p.max_n;
                                if (r4 + 3 < n) {
                                    *reinterpret_cast<float4 *>(gcol + r4) = v4;
                                } else {
                                    gcol[r4] = v4.x;
                                    if (r4 + 1 < n) { gcol[r4 + 1] = v4.y; }
                                    if (r4 + 2 < n) { gcol[r4 + 2] = v4.z; }
                                }
                            }
                        }
                    } else {
                        for (int c = warp; c < n; c += 8) {
                            const int cb = c >> 4;
                            for (int r = lane; r < n; r += 32) {
                                gl[r + static_cast<long>(c) * p.max_n] = r >= 16 * cb ? lp[Lay::Base(cb) + (c & 15) * Lay::Stride(cb) + (r - 16 * cb)] : 0.f;
                            }
                        }
                    }
                }
                // ---- alpha = L^-T z, written back; alpha also goes beside the points for further query tiles ----
                if (is_t) {
                    BackSolveT(smem, nblk, tid, warp, lane);
                    NamedBarrier(1, 128);
                    if (row < n) {
                        const float a = al[row];
                        p.alpha[static_cast<long>(g) * p.max_n + row] = a;
                        smem[Lay::kPts + 4 * row + 3] = a;
                        smem[Lay::kSoa + 3 * Lay::kNp + row] = a;
                    }
                    if (tid == 0) { p.info[g] = 0; }
                    // ---- query tiles beyond the first 128: the mma.sync predict of erl_gp_rowgp.cuh on the same layout ----
                    if (q1 - q0 > 128) {
                        NamedBarrier(1, 128);
                        for (long qb = q0 + 128; qb < q1; qb += 64) {
                            const int nq2 = static_cast<int>(q1 - qb < 64 ? q1 - qb : 64);
                            rowgp::PredictTileMma<XDIM, 8, false>(p, cov, smem, n, nblk, qb, nq2);
                        }
                    }
                }
                __syncthreads();  // the next GP overwrites the staging area
            }
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
