// Micro-benchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue rate on sm_100a, alone and fed by warp-uniform
// LDS.128 broadcasts (the inner loop of the thread-per-query forward substitution in erl_gp_rowgp.cuh).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_lds_rate tools/fma_lds_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 2048;
constexpr int NV = 64;  // accumulators per thread

__global__ void k_ffma2(float* out) {
    float2 c[NV / 2];
    for (int i = 0; i < NV / 2; ++i) c[i] = make_float2(i, i + 0.5f);
    const float2 a = make_float2(threadIdx.x * 1e-9f, 1e-9f), b = make_float2(1.0001f, 0.9999f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) c[i] = __ffma2_rn(c[i], b, a);
    }
    float s = 0; for (int i = 0; i < NV / 2; ++i) s += c[i].x + c[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// v[i] -= L[i] * vj with L read by warp-uniform LDS.128, plain FFMA. ROWS_PER_LOAD = 4.
template<int PACKED, int SPLIT>
__global__ void k_sub(float* out, int cols) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < cols * NV + 64; i += blockDim.x) sm[i] = 1e-6f * (i & 255);
    __syncthreads();
    float2 v[NV / 2];
    for (int i = 0; i < NV / 2; ++i) v[i] = make_float2(i, i + 0.5f);
    // SPLIT: lanes alternate between two adjacent 16-byte granules (layout B': 2 distinct addresses / warp)
    const int lane_off = SPLIT ? (threadIdx.x & 1) * 4 : 0;
    for (int it = 0; it < ITERS / 64; ++it) {
        for (int j = 0; j < cols; ++j) {
            const float vj = v[0].x * 1e-3f + j;
            const float2 vj2 = make_float2(vj, vj);
            const float4* col = reinterpret_cast<const float4*>(sm + j * NV + lane_off);
#pragma unroll
            for (int q = 0; q < NV / 4; ++q) {
                const float4 l = col[q];
                if (PACKED) {
                    v[2 * q] = __ffma2_rn(make_float2(l.x, l.y), vj2, v[2 * q]);
                    v[2 * q + 1] = __ffma2_rn(make_float2(l.z, l.w), vj2, v[2 * q + 1]);
                } else {
                    v[2 * q].x = fmaf(l.x, vj, v[2 * q].x);
                    v[2 * q].y = fmaf(l.y, vj, v[2 * q].y);
                    v[2 * q + 1].x = fmaf(l.z, vj, v[2 * q + 1].x);
                    v[2 * q + 1].y = fmaf(l.w, vj, v[2 * q + 1].y);
                }
            }
        }
    }
    float s = 0; for (int i = 0; i < NV / 2; ++i) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// two queries per thread: each LDS.128 feeds 8 FMAs
template<int PACKED>
__global__ void k_sub2(float* out, int cols) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < cols * NV + 64; i += blockDim.x) sm[i] = 1e-6f * (i & 255);
    __syncthreads();
    float2 v[NV / 2], w[NV / 2];
    for (int i = 0; i < NV / 2; ++i) { v[i] = make_float2(i, i + 0.5f); w[i] = make_float2(i, i + 0.25f); }
    const int lane_off = (threadIdx.x & 1) * 4;
    for (int it = 0; it < ITERS / 64; ++it) {
        for (int j = 0; j < cols; ++j) {
            const float vj = v[0].x * 1e-3f + j, wj = w[0].x * 1e-3f + j;
            const float2 vj2 = make_float2(vj, vj), wj2 = make_float2(wj, wj);
            const float4* col = reinterpret_cast<const float4*>(sm + j * NV + lane_off);
#pragma unroll
            for (int q = 0; q < NV / 4; ++q) {
                const float4 l = col[q];
                if (PACKED) {
                    v[2 * q] = __ffma2_rn(make_float2(l.x, l.y), vj2, v[2 * q]);
                    v[2 * q + 1] = __ffma2_rn(make_float2(l.z, l.w), vj2, v[2 * q + 1]);
                    w[2 * q] = __ffma2_rn(make_float2(l.x, l.y), wj2, w[2 * q]);
                    w[2 * q + 1] = __ffma2_rn(make_float2(l.z, l.w), wj2, w[2 * q + 1]);
                } else {
                    v[2 * q].x = fmaf(l.x, vj, v[2 * q].x);
                    v[2 * q].y = fmaf(l.y, vj, v[2 * q].y);
                    v[2 * q + 1].x = fmaf(l.z, vj, v[2 * q + 1].x);
                    v[2 * q + 1].y = fmaf(l.w, vj, v[2 * q + 1].y);
                    w[2 * q].x = fmaf(l.x, wj, w[2 * q].x);
                    w[2 * q].y = fmaf(l.y, wj, w[2 * q].y);
                    w[2 * q + 1].x = fmaf(l.z, wj, w[2 * q + 1].x);
                    w[2 * q + 1].y = fmaf(l.w, wj, w[2 * q + 1].y);
                }
            }
        }
    }
    float s = 0; for (int i = 0; i < NV / 2; ++i) s += v[i].x + v[i].y + w[i].x + w[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_it(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    void* out; CK(cudaMalloc(&out, (size_t)sms * 16 * 256 * 8));
    float ms;
    {
        const int threads = 256, blocks = sms * 4;
        const double warps = (double)blocks * threads / 32;
        ms = time_it([&] { k_ffma2<<<blocks, threads>>>((float*)out); });
        printf("{\"op\":\"ffma2\",\"ms\":%.4f,\"tflops\":%.2f}\n", ms, warps * 32 * ITERS * NV * 2.0 / ms * 1e-9);
    }
    const int cols = 64;
    const size_t smem = (cols * NV + 64) * sizeof(float);
    for (int threads : {128, 256}) {
        for (int cps : {1, 2, 3, 4}) {  // CTAs per SM
            const int blocks = sms * cps;
            const double warps = (double)blocks * threads / 32;
            const double flop = warps * 32 * (ITERS / 64) * cols * NV * 2.0;
            ms = time_it([&] { k_sub<0, 0><<<blocks, threads, smem>>>((float*)out, cols); });
            printf("{\"op\":\"lds128u+4ffma\",\"threads\":%d,\"cta_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.2f}\n", threads, cps, ms, flop / ms * 1e-9);
            ms = time_it([&] { k_sub<1, 0><<<blocks, threads, smem>>>((float*)out, cols); });
            printf("{\"op\":\"lds128u+2ffma2\",\"threads\":%d,\"cta_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.2f}\n", threads, cps, ms, flop / ms * 1e-9);
            ms = time_it([&] { k_sub<1, 1><<<blocks, threads, smem>>>((float*)out, cols); });
            printf("{\"op\":\"lds128(2addr)+2ffma2\",\"threads\":%d,\"cta_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.2f}\n", threads, cps, ms, flop / ms * 1e-9);
            ms = time_it([&] { k_sub2<0><<<blocks, threads, smem>>>((float*)out, cols); });
            printf("{\"op\":\"lds128(2addr)+8ffma\",\"threads\":%d,\"cta_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.2f}\n", threads, cps, ms, 2 * flop / ms * 1e-9);
            ms = time_it([&] { k_sub2<1><<<blocks, threads, smem>>>((float*)out, cols); });
            printf("{\"op\":\"lds128(2addr)+4ffma2\",\"threads\":%d,\"cta_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.2f}\n", threads, cps, ms, 2 * flop / ms * 1e-9);
        }
    }
    CK(cudaGetLastError());
    cudaFree(out);
    return 0;
}
