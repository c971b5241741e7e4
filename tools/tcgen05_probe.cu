// tcgen05 / TMEM probe for the row-GP tensor kernel (erl_gp_rowgp_tc.cuh): checks, on the B200, every hardware assumption that
// kernel is built on, and measures the latencies its schedule is planned with.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tcgen05_probe tools/tcgen05_probe.cu && tools/tcgen05_probe
// 1. kind::tf32 MMA, A and B from shared memory in the K-major no-swizzle canonical layout (core matrix = 8 rows x 16 bytes,
//    LBO = distance of the two 16-byte K chunks of one K = 8 step, SBO = distance of 8-row groups); D (M = 128) in TMEM, read
//    back with tcgen05.ld.32x32b (thread = row); B taken as a row sub-block of the A buffer.
// 2. How the tensor core narrows FP32 bit patterns to TF32 (truncation vs round-to-nearest): decides the 3xTF32 split.
// 3. 3xTF32 (hi*hi + lo*hi + hi*lo, lo = x - trunc(x)) against the FP64 product.
// 4. A from TMEM (written with tcgen05.st), a_negate, accumulate = 0 / 1, N = 16 / 64 / 112.
// 5. Cycle counts: MMA issue -> commit -> mbarrier wake-up -> tcgen05.ld for the shapes of one panel step.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            std::exit(1);                                                                       \
        }                                                                                       \
    } while (0)

namespace tc {
    __device__ __forceinline__ uint32_t
    SmemAddr(const void *p) {
        return static_cast<uint32_t>(__cvta_generic_to_shared(p));
    }

    __device__ __forceinline__ uint64_t
    SmemDesc(const uint32_t saddr, const uint32_t lbo_bytes, const uint32_t sbo_bytes) {
        uint64_t d = 0;
        d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
        d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
        d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
        d |= 1ull << 46;  // descriptor version (Blackwell)
        return d;         // base offset 0, LBO mode 0, SWIZZLE_NONE
    }

    // kind::tf32, FP32 accumulate, K-major A and B
    __host__ __device__ constexpr uint32_t
    InstrDesc(const int m, const int n, const bool neg_a) {
        return (1u << 4) | (2u << 7) | (2u << 10) | (neg_a ? (1u << 13) : 0u) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
    }

    __device__ __forceinline__ void
    MmaSS(const uint32_t d_tmem, const uint64_t a_desc, const uint64_t b_desc, const uint32_t idesc, const uint32_t accumulate) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    }

    __device__ __forceinline__ void
    MmaTS(const uint32_t d_tmem, const uint32_t a_tmem, const uint64_t b_desc, const uint32_t idesc, const uint32_t accumulate) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
            "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    }

    __device__ __forceinline__ void
    Commit(uint64_t *bar) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(SmemAddr(bar)) : "memory");
    }

    __device__ __forceinline__ void
    MbarInit(uint64_t *bar, const uint32_t count) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count) : "memory");
    }

    __device__ __forceinline__ void
    MbarWait(uint64_t *bar, const uint32_t parity) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra DONE_%=;\n\t"
            "bra WAIT_%=;\n\t"
            "DONE_%=:\n\t}\n" ::"r"(SmemAddr(bar)),
            "r"(parity)
            : "memory");
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
    FenceProxyAsync() {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }

    __device__ __forceinline__ void
    TmemLd16(const uint32_t taddr, float (&v)[16]) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
            "tcgen05.wait::ld.sync.aligned;\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
              "=r"(r[14]), "=r"(r[15])
            : "r"(taddr)
            : "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i] = __uint_as_float(r[i]); }
    }

    __device__ __forceinline__ void
    TmemSt16(const uint32_t taddr, const float (&v)[16]) {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n\t"
            "tcgen05.wait::st.sync.aligned;\n" ::"r"(taddr),
            "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
            "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
            "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
            : "memory");
    }

    template<int COLS>
    __device__ __forceinline__ uint32_t
    TmemAlloc(uint32_t *slot) {  // one full warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemAddr(slot)), "n"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        return 0;
    }

    template<int COLS>
    __device__ __forceinline__ void
    TmemFree(const uint32_t taddr) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
    }
}  // namespace tc

constexpr int kM = 128, kK = 16, kNMax = 128;
constexpr int kChunk = kM * 16;  // bytes between the 16-byte K chunks of the operand buffers (LBO)

// element (row, k) of an operand buffer in the canonical K-major no-swizzle layout
__host__ __device__ inline int
OperandIndex(const int row, const int k) {
    return ((k >> 2) * kChunk + (row >> 3) * 128 + (row & 7) * 16 + (k & 3) * 4) / 4;
}

struct Params {
    const float *a;  // [128][16]
    const float *b;  // [128][16]   (B rows)
    float *out;      // [8 tests][128][128]
    long long *cycles;  // [32]
};

__device__ __forceinline__ float
TruncTf32(const float x) {
    return __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

__global__ void __launch_bounds__(128, 1)
ProbeKernel(const Params p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float *a_hi = reinterpret_cast<float *>(smem);                // 8 KB each
    float *a_lo = reinterpret_cast<float *>(smem + 8192);
    float *b_hi = reinterpret_cast<float *>(smem + 2 * 8192);
    float *b_lo = reinterpret_cast<float *>(smem + 3 * 8192);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 4 * 8192);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + 4 * 8192 + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    // thread = row: 16 values each
    float av[16], bv[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        av[k] = p.a[tid * 16 + k];
        bv[k] = p.b[tid * 16 + k];
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        a_hi[OperandIndex(tid, k)] = av[k];
        a_lo[OperandIndex(tid, k)] = av[k] - TruncTf32(av[k]);
        b_hi[OperandIndex(tid, k)] = bv[k];
        b_lo[OperandIndex(tid, k)] = bv[k] - TruncTf32(bv[k]);
    }
    if (tid == 0) { tc::MbarInit(bar, 1); }
    if (warp == 0) { tc::TmemAlloc<512>(slot); }
    tc::FenceProxyAsync();  // generic-proxy writes of the operands -> visible to the tensor core (async proxy)
    tc::FenceBefore();
    __syncthreads();
    tc::FenceAfter();
    const uint32_t tmem = *slot;
    const uint32_t lane_base = static_cast<uint32_t>(32 * warp) << 16;
    uint32_t parity = 0;
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;

    auto desc = [&](const float *buf, const int row0, const int kstep) { return tc::SmemDesc(tc::SmemAddr(buf) + (row0 >> 3) * 128 + kstep * 2 * kChunk, kChunk, 128); };
    auto readback = [&](const int test, const int ncols, const uint32_t col0) {
        tc::FenceAfter();
        for (int c0 = 0; c0 < ncols; c0 += 16) {
            float v[16];
            tc::TmemLd16(tmem + lane_base + col0 + c0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) { p.out[(test * 128 + tid) * 128 + c0 + i] = v[i]; }
        }
        tc::FenceBefore();
        __syncthreads();
    };

    // ---- test 0: plain TF32, SS, N = 64, K = 16 (two K = 8 steps), accumulate 0 then 1 ----
    if (tid == 0) {
        t0 = clock64();
        tc::MmaSS(tmem, desc(a_hi, 0, 0), desc(b_hi, 0, 0), tc::InstrDesc(128, 64, false), 0);
        tc::MmaSS(tmem, desc(a_hi, 0, 1), desc(b_hi, 0, 1), tc::InstrDesc(128, 64, false), 1);
        tc::Commit(bar);
        t1 = clock64();
    }
    tc::MbarWait(bar, parity);
    parity ^= 1;
    if (tid == 0) { t2 = clock64(); }
    readback(0, 64, 0);
    if (tid == 0) {
        t3 = clock64();
        p.cycles[0] = t1 - t0;  // issue of 2 MMAs + commit
        p.cycles[1] = t2 - t0;  // issue -> mbarrier wake-up
        p.cycles[2] = t3 - t2;  // 4 x (tcgen05.ld x16 + wait + 16 STG) + barrier
    }
    // ---- test 1: 3xTF32, SS, N = 112, B = rows 16 .. 127 of the A-side buffers' twin (row sub-block), into columns 128.. ----
    if (tid == 0) {
        t0 = clock64();
        const uint32_t id = tc::InstrDesc(128, 112, false);
        for (int ks = 0; ks < 2; ++ks) {
            tc::MmaSS(tmem + 128, desc(a_lo, 0, ks), desc(b_hi, 16, ks), id, ks > 0);
            tc::MmaSS(tmem + 128, desc(a_hi, 0, ks), desc(b_lo, 16, ks), id, 1);
            tc::MmaSS(tmem + 128, desc(a_hi, 0, ks), desc(b_hi, 16, ks), id, 1);
        }
        tc::Commit(bar);
        t1 = clock64();
    }
    tc::MbarWait(bar, parity);
    parity ^= 1;
    if (tid == 0) {
        t2 = clock64();
        p.cycles[3] = t1 - t0;  // issue of 6 MMAs (N = 112) + commit
        p.cycles[4] = t2 - t0;  // -> wake-up
    }
    readback(1, 112, 128);
    // ---- test 2: a_negate, N = 16, accumulate onto test 0's columns 0 .. 15 (expected: 0 after adding -A B) ----
    if (tid == 0) {
        t0 = clock64();
        const uint32_t id = tc::InstrDesc(128, 16, true);
        tc::MmaSS(tmem, desc(a_hi, 0, 0), desc(b_hi, 0, 0), id, 1);
        tc::MmaSS(tmem, desc(a_hi, 0, 1), desc(b_hi, 0, 1), id, 1);
        tc::Commit(bar);
        t1 = clock64();
    }
    tc::MbarWait(bar, parity);
    parity ^= 1;
    if (tid == 0) {
        t2 = clock64();
        p.cycles[5] = t1 - t0;  // 2 MMAs N = 16 + commit
        p.cycles[6] = t2 - t0;
    }
    readback(2, 16, 0);
    // ---- test 3: A from TMEM (columns 256 .. 271 <- a, via tcgen05.st), N = 64 into columns 288.. ----
    {
        long long s0 = clock64();
        tc::TmemSt16(tmem + lane_base + 256, av);
        long long s1 = clock64();
        if (tid == 0) { p.cycles[7] = s1 - s0; }  // tcgen05.st x16 + wait::st
        tc::FenceBefore();
        __syncthreads();
        tc::FenceAfter();
        if (tid == 0) {
            t0 = clock64();
            const uint32_t id = tc::InstrDesc(128, 64, false);
            tc::MmaTS(tmem + 288, tmem + 256, desc(b_hi, 0, 0), id, 0);
            tc::MmaTS(tmem + 288, tmem + 256 + 8, desc(b_hi, 0, 1), id, 1);
            tc::Commit(bar);
        }
        tc::MbarWait(bar, parity);
        parity ^= 1;
        if (tid == 0) { p.cycles[8] = clock64() - t0; }
        readback(3, 64, 288);
    }
    // ---- test 4: latency of one tcgen05.ld x16 + wait in isolation ----
    {
        float v[16];
        tc::FenceAfter();
        const long long s0 = clock64();
        tc::TmemLd16(tmem + lane_base, v);
        const long long s1 = clock64();
        if (tid == 0) { p.cycles[9] = s1 - s0; }
        if (v[0] == 123.456f) { p.out[0] = v[1]; }
    }
    // ---- test 5: the whole panel step in one go: 12 MMAs N = 16 (look-ahead part), commit, 12 MMAs N = 96, commit ----
    if (tid == 0) {
        t0 = clock64();
        for (int g = 0; g < 2; ++g) {
            for (int ks = 0; ks < 2; ++ks) {
                const uint32_t id = tc::InstrDesc(128, 16, true);
                tc::MmaSS(tmem + 128 * g, desc(a_lo, 0, ks), desc(b_hi, 16, ks), id, 1);
                tc::MmaSS(tmem + 128 * g, desc(a_hi, 0, ks), desc(b_lo, 16, ks), id, 1);
                tc::MmaSS(tmem + 128 * g, desc(a_hi, 0, ks), desc(b_hi, 16, ks), id, 1);
            }
        }
        tc::Commit(bar);
        t1 = clock64();
        for (int g = 0; g < 2; ++g) {
            for (int ks = 0; ks < 2; ++ks) {
                const uint32_t id = tc::InstrDesc(128, 96, true);
                tc::MmaSS(tmem + 128 * g + 16, desc(a_lo, 0, ks), desc(b_hi, 32, ks), id, 1);
                tc::MmaSS(tmem + 128 * g + 16, desc(a_hi, 0, ks), desc(b_lo, 32, ks), id, 1);
                tc::MmaSS(tmem + 128 * g + 16, desc(a_hi, 0, ks), desc(b_hi, 32, ks), id, 1);
            }
        }
        t2 = clock64();
    }
    tc::MbarWait(bar, parity);  // first commit
    parity ^= 1;
    if (tid == 0) {
        t3 = clock64();
        p.cycles[10] = t1 - t0;  // issue 12 MMAs N = 16 + commit
        p.cycles[11] = t2 - t1;  // issue 12 MMAs N = 96
        p.cycles[12] = t3 - t0;  // first commit's wake-up (the second batch was issued in between)
        tc::Commit(bar);
        t0 = clock64();
    }
    tc::MbarWait(bar, parity);
    parity ^= 1;
    if (tid == 0) { p.cycles[13] = clock64() - t0; }  // tail of the N = 96 batch after its commit
    tc::FenceBefore();
    __syncthreads();
    if (warp == 0) { tc::TmemFree<512>(tmem); }
}

static float
TruncHost(float x) {
    uint32_t u;
    std::memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    std::memcpy(&x, &u, 4);
    return x;
}

static float
RoundHost(float x) {  // round to nearest even on 13 dropped bits
    uint32_t u;
    std::memcpy(&u, &x, 4);
    const uint32_t lsb = (u >> 13) & 1u;
    u += 0xfffu + lsb;
    u &= 0xffffe000u;
    std::memcpy(&x, &u, 4);
    return x;
}

int
main() {
    std::vector<float> a(128 * 16), b(128 * 16);
    srand(7);
    for (auto &v: a) { v = static_cast<float>(rand()) / RAND_MAX * 2.f - 1.f; }
    for (auto &v: b) { v = static_cast<float>(rand()) / RAND_MAX * 2.f - 1.f; }
    float *da, *db, *dout;
    long long *dcyc;
    CK(cudaMalloc(&da, a.size() * 4));
    CK(cudaMalloc(&db, b.size() * 4));
    CK(cudaMalloc(&dout, 8 * 128 * 128 * 4));
    CK(cudaMalloc(&dcyc, 32 * 8));
    CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0, 8 * 128 * 128 * 4));
    CK(cudaMemset(dcyc, 0, 32 * 8));
    const int smem = 4 * 8192 + 256;
    CK(cudaFuncSetAttribute(ProbeKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    Params p{da, db, dout, dcyc};
    for (int rep = 0; rep < 2; ++rep) {  // second run: warm instruction cache for the cycle counts
        ProbeKernel<<<1, 128, smem>>>(p);
        CK(cudaDeviceSynchronize());
    }
    std::vector<float> out(8 * 128 * 128);
    std::vector<long long> cyc(32);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cyc.data(), dcyc, 32 * 8, cudaMemcpyDeviceToHost));

    auto dot = [&](int m, int n, int mode) {  // mode 0 exact, 1 truncated inputs, 2 rounded inputs
        double s = 0;
        for (int k = 0; k < 16; ++k) {
            float x = a[m * 16 + k], y = b[n * 16 + k];
            if (mode == 1) { x = TruncHost(x), y = TruncHost(y); }
            if (mode == 2) { x = RoundHost(x), y = RoundHost(y); }
            s += static_cast<double>(x) * y;
        }
        return s;
    };
    double e_exact = 0, e_trunc = 0, e_round = 0;
    for (int m = 0; m < 128; ++m) {
        for (int n = 0; n < 64; ++n) {
            const double got = out[(0 * 128 + m) * 128 + n];
            e_exact = std::fmax(e_exact, std::fabs(got - dot(m, n, 0)));
            e_trunc = std::fmax(e_trunc, std::fabs(got - dot(m, n, 1)));
            e_round = std::fmax(e_round, std::fabs(got - dot(m, n, 2)));
        }
    }
    std::printf("test0 TF32 SS N=64 K=16: max err vs exact %.3e, vs truncated inputs %.3e, vs rounded inputs %.3e  => %s\n", e_exact, e_trunc, e_round,
                e_trunc < 0.2 * e_round ? "TRUNCATION" : (e_round < 0.2 * e_trunc ? "ROUND-TO-NEAREST" : "UNCLEAR"));
    double e3 = 0;
    for (int m = 0; m < 128; ++m) {
        for (int n = 0; n < 112; ++n) { e3 = std::fmax(e3, std::fabs(out[(1 * 128 + m) * 128 + n] - dot(m, n + 16, 0))); }
    }
    std::printf("test1 3xTF32 SS N=112 (B = rows 16.. of the buffer): max err vs exact %.3e (FP32 dot would be ~5e-7)\n", e3);
    double e_neg = 0;
    for (int m = 0; m < 128; ++m) {
        for (int n = 0; n < 16; ++n) { e_neg = std::fmax(e_neg, std::fabs(out[(2 * 128 + m) * 128 + n])); }
    }
    std::printf("test2 a_negate + accumulate, N=16: max |D - A B| = %.3e (expected 0)\n", e_neg);
    double e_ts = 0;
    for (int m = 0; m < 128; ++m) {
        for (int n = 0; n < 64; ++n) { e_ts = std::fmax(e_ts, std::fabs(out[(3 * 128 + m) * 128 + n] - out[(0 * 128 + m) * 128 + n])); }
    }
    std::printf("test3 A from TMEM (tcgen05.st) vs A from shared memory: max diff %.3e (expected 0)\n", e_ts);
    const char *names[] = {"issue 2 MMA (N=64) + commit", "issue -> mbarrier wake-up, 2 MMA N=64", "4 x (tcgen05.ld x16 + 16 STG) + barrier", "issue 6 MMA (N=112) + commit",
                           "issue -> wake-up, 6 MMA N=112", "issue 2 MMA (N=16) + commit", "issue -> wake-up, 2 MMA N=16", "tcgen05.st x16 + wait::st", "TS: issue -> wake-up, 2 MMA N=64",
                           "tcgen05.ld x16 + wait::ld", "issue 12 MMA N=16 + commit", "issue 12 MMA N=96", "first commit wake-up (24 MMAs queued)", "tail after second commit"};
    for (int i = 0; i < 14; ++i) { std::printf("cycles[%2d] %-48s %lld\n", i, names[i], cyc[i]); }
    const bool ok = e_trunc < 1e-5 || e_round < 1e-5;
    const bool ok3 = e3 < 5e-6 && e_neg < 1e-6 && e_ts < 1e-7;
    std::printf("%s\n", ok && ok3 ? "PROBE PASS" : "PROBE FAIL");
    return ok && ok3 ? 0 : 1;
}
