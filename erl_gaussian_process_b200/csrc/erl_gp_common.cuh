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
        __device__ __forceinline__ static double exp_(double v) { return exp(v); }
        __device__ __forceinline__ static float sqrt_(float v) { return sqrtf(v); }
        __device__ __forceinline__ static double sqrt_(double v) { return sqrt(v); }
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
