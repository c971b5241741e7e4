// Dense building blocks for large VanillaGaussianProcess / SPGP systems (matrix in HBM):
//   Gemm            C = alpha * op(A) * op(B) + beta * C      register-tiled, double-buffered
//   Potrf           blocked right-looking lower Cholesky, panel 128, diagonal-block inverses kept
//   TrsmLower       W <- L^-1 W  (blocked, as GEMMs with the kept diagonal inverses), optional
//                   per-column ||.||^2 accumulation (the predictive-variance term)
//   TrsmLowerTrans  Z <- L^-T Z
// Everything is column-major with explicit leading dimensions and asynchronous on ctx->stream.
#pragma once

#include "erl_gp_internal.cuh"

namespace erl_gp {

    constexpr int kPanel = 128;  // panel / diagonal block edge of the HBM-resident factorisation

    enum GemmOp : int { kOpN = 0, kOpT = 1 };

    template<typename T>
    int
    Gemm(Context *ctx, int op_a, int op_b, long m, long n, long k, T alpha, const T *a, long lda, const T *b, long ldb, T beta, T *c, long ldc, int lower_only);  // lower_only: 0 full, 1 lower tiles, 2 lower tiles except tile (0, 0)

    // Factor the n x n lower triangle at `l` (ld) in place; linv receives ceil(n/128) inverses of the
    // 128 x 128 diagonal blocks (each 128 x 128 col-major, identity-padded); panel is an n x 128 workspace;
    // info (device int) = 0 or the 1-based failing column.
    template<typename T>
    int
    Potrf(Context *ctx, long n, T *l, long ld, T *linv, T *panel, int *info);

    // W (n x t, ldw) <- L^-1 W.  s_buf: 128 x t workspace.  sumsq (t) accumulates column ||.||^2 when not null
    // (caller zeroes it).  keep: write the solved rows back into W (needed when W itself is the result).
    template<typename T>
    int
    TrsmLower(Context *ctx, long n, long t, const T *l, long ld, const T *linv, T *w, long ldw, T *s_buf, T *sumsq, bool keep);

    // Z (n x t, ldz) <- L^-T Z
    template<typename T>
    int
    TrsmLowerTrans(Context *ctx, long n, long t, const T *l, long ld, const T *linv, T *z, long ldz, T *s_buf);

    // Y (n x y_dim, ldy) <- L^-T L^-1 Y with vector kernels; ERL_GP_STATUS_UNSUPPORTED when y_dim > 4 (use the TRSMs)
    template<typename T>
    int
    TrsvSolve(Context *ctx, long n, long y_dim, const T *l, long ld, const T *linv, T *y, long ldy);

    // L = tril(K), strict upper zero (the reference's `mat_l = ktrain.llt().matrixL()` dense assignment)
    template<typename T>
    int
    CopyLower(Context *ctx, long n, const T *k, long ld_k, T *l, long ld_l);

    // out[j + c * t] = sum_i W[i, j] * alpha[i, c]   (mean = Kt^T alpha)
    template<typename T>
    int
    GemvT(Context *ctx, long n, long t, const T *w, long ldw, const T *alpha, long ld_a, long y_dim, T *out, long ld_out);

    // Fused predict of the dense VanillaGaussianProcess (erl_gp_predict_dense.cu): Ktest generated on the fly.
    //   PredictMean:     out[j + c * ld_out] = sum_i k(x_i, x*_j) alpha[i + c * ld_a]   (ERL_GP_STATUS_UNSUPPORTED when y_dim > 4)
    //   PredictVariance: sumsq[j] = || L^-1 k(X, x*_j) ||^2, left-looking, one CTA per 128 test points; v_slabs needs
    //                    PredictVarianceSlabElems() elements
    template<typename T>
    size_t
    PredictVarianceSlabElems(const Context *ctx, long n, long t);
    template<typename T>
    int
    PredictVariance(Context *ctx, int kernel, T scale, long x_dim, long n, long t, const T *x_train, const T *x_test, const T *l, long ldl, const T *linv, T *v_slabs, T *sumsq);
    template<typename T>
    int
    PredictMean(Context *ctx, int kernel, T scale, long x_dim, long n, long t, const T *x_train, const T *x_test, const T *alpha, long ld_a, long y_dim, T *out, long ld_out);

    // var[j] = 1 - a[j] (+ b[j] when b != nullptr)
    template<typename T>
    int
    VarianceFinalize(Context *ctx, long t, const T *a, const T *b, T *var);

}  // namespace erl_gp
