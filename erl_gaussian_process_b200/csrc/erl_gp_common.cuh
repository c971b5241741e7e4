// Shared device/host helpers for the erl_gaussian_process_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>

#include "erl_gp_b200.h"

namespace erl_gp {

    constexpr int kNB = 16;  // diagonal / panel block edge of the in-shared-memory factorisation

    // ---- status plumbing -------------------------------------------------------------------
    struct Context;
    int
    SetError(Context *ctx, int status, const char *fmt, ...);

#define ERL_GP_CUDA_OK(ctx, expr)                                                                        \
    do {                                                                                                 \
        cudaError_t err__ = (expr);                                                                      \
        if (err__ != cudaSuccess) {                                                                      \
            return ::erl_gp::SetError(ctx, ERL_GP_STATUS_CUDA_ERROR, "%s:%d %s -> %s", __FILE__, __LINE__, \
                                      #expr, cudaGetErrorString(err__));                                 \
        }                                                                                                \
    } while (0)

    // exp / sqrt of the covariance entries in double.  The library versions cost ~100 FP64 instructions per entry (IEEE sqrt + exp with
    // their special cases); these take ~40 and are accurate to a few ulp on the arguments that occur (r2 >= 0, exponent argument <= 0):
    //   sqrt(r2) = r2 * rsqrt(r2) with one Newton correction (0 at r2 = 0);
    //   exp(x)   = 2^k p(r), k = rint(x log2 e), r = x - k ln2 (two-term Cody-Waite), p = Taylor polynomial of degree 12 on |r| <= 0.347
    //              (truncation error 1.7e-16; 4.7e-16 max relative error measured against numpy on [-60, 0]), 2^k in the exponent field.
    __device__ __forceinline__ double
    FastSqrt(const double r2) {
        const double y = rsqrt(r2);
        double r = r2 * y;
        r = fma(0.5 * y, fma(-r, r, r2), r);
        return r2 > 0.0 ? r : r2;  // 0 at 0, NaN stays NaN
    }

    __device__ __forceinline__ double
    FastExpNeg(const double x) {  // x <= 0
        constexpr double kLog2e = 1.4426950408889634074, kLn2Hi = 6.93147180369123816490e-01, kLn2Lo = 1.90821492927058770002e-10, kMagic = 6755399441055744.0;
        const double kd = fma(x, kLog2e, kMagic) - kMagic;  // rint(x log2 e)
        double r = fma(kd, -kLn2Hi, x);
        r = fma(kd, -kLn2Lo, r);
        double p = 1.0 / 479001600.0;
        p = fma(p, r, 1.0 / 39916800.0);
        p = fma(p, r, 1.0 / 3628800.0);
        p = fma(p, r, 1.0 / 362880.0);
        p = fma(p, r, 1.0 / 40320.0);
        p = fma(p, r, 1.0 / 5040.0);
        p = fma(p, r, 1.0 / 720.0);
        p = fma(p, r, 1.0 / 120.0);
        p = fma(p, r, 1.0 / 24.0);
        p = fma(p, r, 1.0 / 6.0);
        p = fma(p, r, 0.5);
        p = fma(p, r, 1.0);
        p = fma(p, r, 1.0);
        const int k = static_cast<int>(kd);
        const double scale = __hiloint2double((k + 1023) << 20, 0);  // 2^k, k >= -1022
        return x < -708.0 ? 0.0 : p * scale;  // (NaN stays NaN)
    }

    // ---- covariance functions (erl_covariance v0.2.0) ------------------
    //   OrnsteinUhlenbeck : exp(-r / l)
    //   Matern32          : (1 + sqrt(3) r / l) exp(-sqrt(3) r / l)
    //   RadialBiasFunction: exp(-r^2 / (2 l^2))
    // The functor is built once on the host (coefficients precomputed in the working precision,
    // same expression order as the CPU reference restatement) and passed by value to the kernels.
    template<typename T>
    struct Covariance {
        int type;
        T c0;  // OU: l   Matern32: sqrt(3)/l   RBF: 2 l^2

        __host__ static Covariance
        Make(int type, T scale) {
            Covariance c;
            c.type = type;
            if (type == ERL_GP_KERNEL_RBF) {
                c.c0 = T(2) * scale * scale;
            } else if (type == ERL_GP_KERNEL_MATERN32) {
                c.c0 = std::sqrt(T(3)) / scale;
            } else {
                c.c0 = scale;
            }
            return c;
        }

        __device__ __forceinline__ T
        operator()(T r2) const {
            if (type == ERL_GP_KERNEL_RBF) { return exp_(-r2 / c0); }
            const T r = sqrt_(r2);
            if (type == ERL_GP_KERNEL_MATERN32) {
                const T ar = c0 * r;
                return (T(1) + ar) * exp_(-ar);
            }
            return exp_(-r / c0);
        }

        __device__ __forceinline__ static float exp_(float v) { return expf(v); }
        __device__ __forceinline__ static double exp_(double v) { return FastExpNeg(v); }
        __device__ __forceinline__ static float sqrt_(float v) { return sqrtf(v); }
        __device__ __forceinline__ static double sqrt_(double v) { return FastSqrt(v); }
    };

    template<typename T, int XDIM>
    __device__ __forceinline__ T
    SquaredDistance(const T *__restrict__ a, const T *__restrict__ b) {
        T r2 = 0;
#pragma unroll
        for (int d = 0; d < XDIM; ++d) {
            const T diff = a[d] - b[d];
            r2 += diff * diff;
        }
        return r2;
    }

    // ---- Mapping<Dtype>::map / inv — src/mapping.cpp:112-164 --------------------------------
    template<typename T>
    __host__ __device__ __forceinline__ T
    MappingMap(int type, T scale, T x) {
        switch (type) {
            case ERL_GP_MAPPING_IDENTITY: return x;
            case ERL_GP_MAPPING_INVERSE: return T(1) / x;
            case ERL_GP_MAPPING_INVERSE_SQRT: return T(1) / sqrt(x);
            case ERL_GP_MAPPING_EXP: return exp(-scale * x);
            case ERL_GP_MAPPING_LOG: return log(scale * x);
            case ERL_GP_MAPPING_TANH: return tanh(scale * x);
            case ERL_GP_MAPPING_SIGMOID: return T(1) / (T(1) + exp(-scale * x));
            default: return x;
        }
    }

    template<typename T>
    __host__ __device__ __forceinline__ T
    MappingInv(int type, T scale, T y) {
        switch (type) {
            case ERL_GP_MAPPING_IDENTITY: return y;
            case ERL_GP_MAPPING_INVERSE: return T(1) / y;
            case ERL_GP_MAPPING_INVERSE_SQRT: return T(1) / (y * y);
            case ERL_GP_MAPPING_EXP: return -log(y) / scale;
            case ERL_GP_MAPPING_LOG: return exp(y) / scale;
            case ERL_GP_MAPPING_TANH: return atanh(y) / scale;
            case ERL_GP_MAPPING_SIGMOID:
                if (y >= T(1)) { return T(INFINITY) / scale; }
                if (y <= T(0)) { return -T(INFINITY) / scale; }
                return log(y / (T(1) - y)) / scale;
            default: return y;
        }
    }

    __host__ __device__ constexpr long
    CeilDiv(long a, long b) {
        return (a + b - 1) / b;
    }

}  // namespace erl_gp
