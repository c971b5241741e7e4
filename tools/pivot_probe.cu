// Latency of the 16 x 16 pivot tile routines of the row-GP kernels (rolled rowgp_tc::PivotTile vs unrolled rowgp::PivotBlock<2>),
// alone and next to background warps of one kind (FFMA2 stream, broadcast LDS.128 stream, MUFU stream, mbarrier try_wait spin,
// STS.128 stream): which kind of neighbour slows the serial chain of the pivot warp down.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I erl_gaussian_process_b200/csrc --expt-relaxed-constexpr tools/pivot_probe.cu -o tools/pivot_probe
#include "erl_gp_rowgp_tc.cuh"

#include <cstdio>

using namespace erl_gp;

__global__ void
Probe(float *out, long long *cycles, const int reps, const int mode, const int bg_kind, const int pivot_last) {
    __shared__ __align__(16) float rs[128], al[128], lbuf[16 * 136], dbuf[16 * 20], junk[32 * 64];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ volatile int stop;
    const int lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const int warp = pivot_last ? (nwarps - 1 - (threadIdx.x >> 5)) : (threadIdx.x >> 5);  // logical role: 0 = pivot warp
    if (threadIdx.x == 0) {
        stop = 0;
        rowgp_tc::MbarInit(rowgp_tc::SmemAddr(&bar), 1);
    }
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) { junk[i] = 1.0f + 1e-3f * i; }
    __syncthreads();
    if (warp > 0) {
        // background warps
        float2 acc[8];
        for (int i = 0; i < 8; ++i) { acc[i] = make_float2(1.f + i, 2.f + lane); }
        float s = 0.f;
        while (!stop) {
            if (bg_kind == 0) {
#pragma unroll
                for (int k = 0; k < 64; ++k) { acc[k & 7] = __ffma2_rn(acc[k & 7], make_float2(1.0001f, 0.9999f), make_float2(1e-6f, 1e-6f)); }
            } else if (bg_kind == 1) {
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const float4 v = *reinterpret_cast<const float4 *>(junk + 4 * ((k + warp) & 63));
                    s += v.x + v.w;
                }
            } else if (bg_kind == 2) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    float e;
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(acc[k & 7].x));
                    acc[k & 7].x = e * 0.5f;
                }
            } else if (bg_kind == 3) {
                uint32_t done;
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(rowgp_tc::SmemAddr(&bar)), "r"(0u) : "memory");
                s += done;
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) { *reinterpret_cast<float4 *>(junk + 128 * (warp & 7) + 4 * lane + 1024 * (k & 1)) = make_float4(s, s, s, s); }
                s += 1.f;
            }
        }
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0].x + acc[3].y + s;
        return;
    }
    const int r = lane & 15;
    float base[16];
    for (int c = 0; c < 16; ++c) { base[c] = lane < 16 ? (c == r ? 2.0f + 0.01f * r : 0.3f / (1.0f + abs(c - r))) : (c == r ? 1.0f : 0.f); }
    float acc_out = 0.f;
    int fail = 0;
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < reps; ++it) {
        float a[16];
        for (int c = 0; c < 16; ++c) { a[c] = base[c] + 1e-6f * acc_out; }
        if (mode == 0) {
            float *o = lane < 16 ? lbuf + r : dbuf + r * 20;
            rowgp_tc::PivotTile(a, 0.5f, 0, lane, fail, rs, al, o, lane < 16 ? 136 : 1, nullptr, 0, 16);
            __syncwarp();
            acc_out += lbuf[r] + dbuf[r * 20 + 15];
        } else {
            float l[16];
            float z = 0.5f;
            rowgp::PivotBlock<2>(a, z, l, 0, 0, lane, fail, rs, al);
            acc_out += l[0] + l[15];
        }
    }
    const long long t1 = clock64();
    if (lane == 0) {
        cycles[blockIdx.x] = t1 - t0;
        stop = 1;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc_out + fail;
}

int
main() {
    float *out;
    long long *cyc;
    cudaMalloc(&out, 1024 * 1024 * 4);
    cudaMalloc(&cyc, 4096 * 8);
    const int reps = 100;
    const char *kinds[] = {"FFMA2", "LDS.128 broadcast", "MUFU.EX2", "mbarrier.try_wait spin", "STS.128"};
    for (int mode = 1; mode < 2; ++mode) {
        for (int bg = 0; bg <= 10; bg += (bg == 0 ? 4 : 6)) {  // 0, 4, 10 background warps per CTA, 2 CTAs per SM
            for (int kind = 0; kind < (bg == 0 ? 1 : 5); ++kind) {
                const int grid = 296;
                for (int pl = 0; pl < (bg == 0 ? 1 : 2); ++pl) {
                Probe<<<grid, 32 * (1 + bg)>>>(out, cyc, reps, mode, kind, pl);
                cudaDeviceSynchronize();
                long long h[4096];
                cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
                long long mx = 0, sum = 0;
                for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx, sum += h[i]; }
                printf("%s + %2d background warps (%-22s), pivot = %s warp: %6lld cycles per tile (mean over CTAs), max %6lld  err=%d\n", mode == 0 ? "rolled  " : "unrolled", bg,
                       bg == 0 ? "-" : kinds[kind], pl ? "last " : "first", sum / grid / reps, mx / reps, (int) cudaGetLastError());
                }
            }
        }
    }
    return 0;
}
