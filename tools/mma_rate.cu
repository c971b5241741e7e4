// Micro-benchmark: issue rates of legacy mma.sync shapes and plain FMA pipes on sm_100a.
// Used once to choose the math path of the batched / dense GP kernels.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

__global__ void k_tf32(float* out) {
    float c[CHAINS][4];
    for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, b0 = 4, b1 = 5;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bf16(float* out) {
    float c[CHAINS][4];
    for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, b0 = 4, b1 = 5;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_f64_884(double* out) {
    double c[CHAINS][2];
    for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 2; ++j) c[i][j] = 0.;
    double a = threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 2; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_f64_16816(double* out) {
    double c[CHAINS][4];
    for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.;
    double a[8], b[4];
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-9 + j;
    for (int j = 0; j < 4; ++j) b[j] = 1e-9 * j;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0; for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out) {
    float c[CHAINS * 4];
    for (int i = 0; i < CHAINS * 4; ++i) c[i] = i;
    float a = threadIdx.x * 1e-9f, b = 1.0001f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS * 4; ++i) c[i] = fmaf(c[i], b, a);
    }
    float s = 0; for (int i = 0; i < CHAINS * 4; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double* out) {
    double c[CHAINS * 2];
    for (int i = 0; i < CHAINS * 2; ++i) c[i] = i;
    double a = threadIdx.x * 1e-9, b = 1.0001;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS * 2; ++i) c[i] = fma(c[i], b, a);
    }
    double s = 0; for (int i = 0; i < CHAINS * 2; ++i) s += c[i];
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
    const int threads = 256, blocks = sms * 4;
    void* out; CK(cudaMalloc(&out, (size_t)blocks * threads * 8));
    const double warps = (double)blocks * threads / 32;
    struct R { const char* name; double flop_per_warp_iter; float ms; };
    float ms;
    ms = time_it([&] { k_tf32<<<blocks, threads>>>((float*)out); });
    printf("{\"op\":\"mma.m16n8k8.tf32\",\"ms\":%.4f,\"tflops\":%.2f}\n", ms, warps * ITERS * CHAINS * 2.0 * 16 * 8 * 8 / ms * 1e-9);
    ms = time_it([&] { k_bf16<<<blocks, threads>>>((float*)out); });
    printf("{\"op\":\"mma.m16n8k16.bf16\",\"ms\":%.4f,\"tflops\":%.2f}\n", ms, warps * ITERS * CHAINS * 2.0 * 16 * 8 * 16 / ms * 1e-9);
    ms = time_it([&] { k_f64_884<<<blocks, threads>>>((double*)out); });
    printf("{\"op\":\"mma.m8n8k4.f64\",\"ms\":%.4f,\"tflops\":%.2f}\n", ms, warps * ITERS * CHAINS * 2.0 * 8 * 8 * 4 / ms * 1e-9);
    ms = time_it([&] { k_f64_16816<<<blocks, threads>>>((double*)out); });
    printf("{\"op\":\"mma.m16n8k16.f64\",\"ms\":%.4f,\"tflops\":%.2f}\n", ms, warps * ITERS * CHAINS * 2.0 * 16 * 8 * 16 / ms * 1e-9);
    ms = time_it([&] { k_ffma<<<blocks, threads>>>((float*)out); });
    printf("{\"op\":\"ffma\",\"ms\":%.4f,\"tflops\":%.2f}\n", ms, warps * 32 * ITERS * CHAINS * 4 * 2.0 / ms * 1e-9);
    ms = time_it([&] { k_dfma<<<blocks, threads>>>((double*)out); });
    printf("{\"op\":\"dfma\",\"ms\":%.4f,\"tflops\":%.2f}\n", ms, warps * 32 * ITERS * CHAINS * 2 * 2.0 / ms * 1e-9);
    CK(cudaGetLastError());
    cudaFree(out);
    return 0;
}
