// VanillaGaussianProcess (any n, device resident) and dense SparsePseudoInputGaussianProcess on top of
// the dense building blocks (erl_gp_dense.cuh) and the fused Gram kernels.
//
//   VanillaGaussianProcess::UpdateKtrain / Solve      src/vanilla_gp.cpp:476-505
//   VanillaGaussianProcess::ComputeKtest / Test       src/vanilla_gp.cpp:521-559
//   TestResult::GetMean / GetVariance                 src/vanilla_gp.cpp:61-150
//   SparsePseudoInputGaussianProcess ctor / UpdateDense / PrepareLqm / TestResult
//                                                     src/sparse_pseudo_input_gp.cpp:313-356, 751-791, 835-842, 43-113, 133-163, 280-310
// Test points are processed in tiles: the n x T Ktest of the reference is never materialised beyond one
// n x tile workspace (config 5: 16384 x 1e6 doubles would be 122 GiB).
#include "erl_gp_dense.cuh"

#include <algorithm>
#include <thread>
#include <vector>

namespace erl_gp {

    // test-point tile: workspace n x tile stays near 256 MiB
    static long
    TestTile(const long n, const size_t elem) {
        long tile = static_cast<long>((size_t(1) << 28) / (static_cast<size_t>(n) * elem));
        tile = (tile / 128) * 128;
        if (tile < 256) { tile = 256; }
        if (tile > 8192) { tile = 8192; }
        return tile;
    }

    template<typename T>
    struct Vanilla {
        Context *ctx = nullptr;
        long n = 0, x_dim = 0, y_dim = 0;
        int kernel = 0;
        T scale = T(1);
        bool trained = false;
        DeviceBuffer<T> x, var, k, l, alpha, linv, panel;
        DeviceBuffer<int> info;
        // test workspaces
        DeviceBuffer<T> xt, w, s_buf, sumsq, mean, variance;
    };

    template<typename T>
    static int
    VanillaTrainDev(Vanilla<T> *gp, int kernel, T scale, long x_dim, long y_dim, long n, const T *x, long ld_x, const T *y, long ld_y, const T *var, cudaMemcpyKind kind) {
        if (gp == nullptr || x == nullptr || y == nullptr || var == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        gp->trained = false;
        if (n <= 0) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "vanilla: num_samples = %ld, it should be > 0", n); }  // src/vanilla_gp.cpp:481-484
        if (x_dim < 1 || x_dim > 3) { return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "vanilla: x_dim=%ld (supported: 1, 2, 3)", x_dim); }
        if (y_dim < 1 || ld_x < x_dim || ld_y < n) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "vanilla: bad y_dim / leading dimensions"); }
        if (kernel < ERL_GP_KERNEL_OU || kernel > ERL_GP_KERNEL_RBF) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "vanilla: unknown kernel %d", kernel); }
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        gp->n = n, gp->x_dim = x_dim, gp->y_dim = y_dim, gp->kernel = kernel, gp->scale = scale;
        const long num_panels = CeilDiv(n, kPanel);
        ERL_GP_CUDA_OK(ctx, gp->x.Reserve(static_cast<size_t>(n) * x_dim));
        ERL_GP_CUDA_OK(ctx, gp->var.Reserve(n));
        ERL_GP_CUDA_OK(ctx, gp->alpha.Reserve(static_cast<size_t>(n) * y_dim));
        ERL_GP_CUDA_OK(ctx, gp->k.Reserve(static_cast<size_t>(n) * n));
        ERL_GP_CUDA_OK(ctx, gp->l.Reserve(static_cast<size_t>(n) * n));
        ERL_GP_CUDA_OK(ctx, gp->linv.Reserve(static_cast<size_t>(num_panels) * kPanel * kPanel));
        ERL_GP_CUDA_OK(ctx, gp->panel.Reserve(static_cast<size_t>(n) * kPanel));
        ERL_GP_CUDA_OK(ctx, gp->s_buf.Reserve(static_cast<size_t>(kPanel) * (y_dim > 8192 ? y_dim : 8192)));
        ERL_GP_CUDA_OK(ctx, gp->info.Reserve(1));
        // stage the training set (x packed to ld = x_dim); alpha <- y (src/vanilla_gp.cpp:485)
        ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(gp->x.ptr, sizeof(T) * x_dim, x, sizeof(T) * ld_x, sizeof(T) * x_dim, n, kind, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(gp->alpha.ptr, sizeof(T) * n, y, sizeof(T) * ld_y, sizeof(T) * n, y_dim, kind, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->var.ptr, var, sizeof(T) * n, kind, ctx->stream));
        int rc = LaunchKtrain<T>(ctx, kernel, scale, x_dim, gp->x.ptr, x_dim, gp->var.ptr, n, gp->k.ptr, n);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        rc = CopyLower<T>(ctx, n, gp->k.ptr, n, gp->l.ptr, n);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        rc = Potrf<T>(ctx, n, gp->l.ptr, n, gp->linv.ptr, gp->panel.ptr, gp->info.ptr);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        // alpha = L^-T L^-1 y (src/vanilla_gp.cpp:501-502)
        rc = TrsvSolve<T>(ctx, n, y_dim, gp->l.ptr, n, gp->linv.ptr, gp->alpha.ptr, n);
        if (rc == ERL_GP_STATUS_UNSUPPORTED) {  // many outputs: GEMM-based triangular solves
            rc = TrsmLower<T>(ctx, n, y_dim, gp->l.ptr, n, gp->linv.ptr, gp->alpha.ptr, n, gp->s_buf.ptr, nullptr, true);
            if (rc != ERL_GP_STATUS_OK) { return rc; }
            rc = TrsmLowerTrans<T>(ctx, n, y_dim, gp->l.ptr, n, gp->linv.ptr, gp->alpha.ptr, n, gp->s_buf.ptr);
        }
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        gp->trained = true;
        return ERL_GP_STATUS_OK;
    }

    // mean: num_test x y_dim (ld = num_test) or null; var: num_test or null.  kind selects host / device pointers.
    template<typename T>
    static int
    VanillaTest(Vanilla<T> *gp, long num_test, const T *x_test, long ld_xt, T *mean, T *var, cudaMemcpyKind in_kind, cudaMemcpyKind out_kind, long ld_mean = 0) {
        if (gp == nullptr || x_test == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        if (ld_mean < num_test) { ld_mean = num_test; }  // rows between two output columns of `mean` (a slice of a longer test set: the whole set's length)
        Context *ctx = gp->ctx;
        if (!gp->trained) { return SetError(ctx, ERL_GP_STATUS_NOT_TRAINED, "vanilla: Test() before Train()"); }  // src/vanilla_gp.cpp:556-558
        if (num_test <= 0) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "vanilla: num_test = %ld, it should be > 0", num_test); }  // :529-532
        if (ld_xt < gp->x_dim) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "vanilla: ld_xt < x_dim"); }
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        const long n = gp->n, d = gp->x_dim;
        // ---- fused path (erl_gp_predict_dense.cu): Ktest is generated where it is consumed, test points in chunks of <= 2^20 ----
        if (gp->y_dim <= 4 || mean == nullptr) {
            const long chunk = std::min<long>(num_test, 1L << 20);
            ERL_GP_CUDA_OK(ctx, gp->xt.Reserve(static_cast<size_t>(chunk) * d));
            if (mean != nullptr) { ERL_GP_CUDA_OK(ctx, gp->mean.Reserve(static_cast<size_t>(chunk) * gp->y_dim)); }
            if (var != nullptr) {
                ERL_GP_CUDA_OK(ctx, gp->sumsq.Reserve(chunk));
                ERL_GP_CUDA_OK(ctx, gp->variance.Reserve(chunk));
                ERL_GP_CUDA_OK(ctx, gp->w.Reserve(PredictVarianceSlabElems<T>(ctx, n, chunk)));
            }
            for (long t0 = 0; t0 < num_test; t0 += chunk) {
                const long tt = std::min(chunk, num_test - t0);
                ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(gp->xt.ptr, sizeof(T) * d, x_test + t0 * ld_xt, sizeof(T) * ld_xt, sizeof(T) * d, tt, in_kind, ctx->stream));
                if (mean != nullptr) {
                    const int rc = PredictMean<T>(ctx, gp->kernel, gp->scale, d, n, tt, gp->x.ptr, gp->xt.ptr, gp->alpha.ptr, n, gp->y_dim, gp->mean.ptr, tt);
                    if (rc != ERL_GP_STATUS_OK) { return rc; }
                    ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(mean + t0, sizeof(T) * ld_mean, gp->mean.ptr, sizeof(T) * tt, sizeof(T) * tt, gp->y_dim, out_kind, ctx->stream));
                }
                if (var != nullptr) {
                    int rc = PredictVariance<T>(ctx, gp->kernel, gp->scale, d, n, tt, gp->x.ptr, gp->xt.ptr, gp->l.ptr, n, gp->linv.ptr, gp->w.ptr, gp->sumsq.ptr);
                    if (rc != ERL_GP_STATUS_OK) { return rc; }
                    rc = VarianceFinalize<T>(ctx, tt, gp->sumsq.ptr, nullptr, gp->variance.ptr);
                    if (rc != ERL_GP_STATUS_OK) { return rc; }
                    ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(var + t0, gp->variance.ptr, sizeof(T) * tt, out_kind, ctx->stream));
                }
            }
            return ERL_GP_STATUS_OK;
        }
        // ---- materialised-tile path (y_dim > 4): Ktest tile, GEMV mean, right-looking solve ----
        const long tile = std::min(TestTile(n, sizeof(T)), ((num_test + 127) / 128) * 128);
        ERL_GP_CUDA_OK(ctx, gp->xt.Reserve(static_cast<size_t>(tile) * d));
        ERL_GP_CUDA_OK(ctx, gp->w.Reserve(static_cast<size_t>(n) * tile));
        ERL_GP_CUDA_OK(ctx, gp->s_buf.Reserve(static_cast<size_t>(kPanel) * tile));
        ERL_GP_CUDA_OK(ctx, gp->sumsq.Reserve(tile));
        ERL_GP_CUDA_OK(ctx, gp->mean.Reserve(static_cast<size_t>(tile) * gp->y_dim));
        ERL_GP_CUDA_OK(ctx, gp->variance.Reserve(tile));
        for (long t0 = 0; t0 < num_test; t0 += tile) {
            const long tt = std::min(tile, num_test - t0);
            ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(gp->xt.ptr, sizeof(T) * d, x_test + t0 * ld_xt, sizeof(T) * ld_xt, sizeof(T) * d, tt, in_kind, ctx->stream));
            int rc = LaunchKtest<T>(ctx, gp->kernel, gp->scale, d, gp->x.ptr, d, n, gp->xt.ptr, d, tt, gp->w.ptr, n);
            if (rc != ERL_GP_STATUS_OK) { return rc; }
            if (mean != nullptr) {
                rc = GemvT<T>(ctx, n, tt, gp->w.ptr, n, gp->alpha.ptr, n, gp->y_dim, gp->mean.ptr, tt);
                if (rc != ERL_GP_STATUS_OK) { return rc; }
                ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(mean + t0, sizeof(T) * ld_mean, gp->mean.ptr, sizeof(T) * tt, sizeof(T) * tt, gp->y_dim, out_kind, ctx->stream));
            }
            if (var != nullptr) {
                ERL_GP_CUDA_OK(ctx, cudaMemsetAsync(gp->sumsq.ptr, 0, sizeof(T) * tt, ctx->stream));
                rc = TrsmLower<T>(ctx, n, tt, gp->l.ptr, n, gp->linv.ptr, gp->w.ptr, n, gp->s_buf.ptr, gp->sumsq.ptr, false);
                if (rc != ERL_GP_STATUS_OK) { return rc; }
                rc = VarianceFinalize<T>(ctx, tt, gp->sumsq.ptr, nullptr, gp->variance.ptr);
                if (rc != ERL_GP_STATUS_OK) { return rc; }
                ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(var + t0, gp->variance.ptr, sizeof(T) * tt, out_kind, ctx->stream));
            }
        }
        return ERL_GP_STATUS_OK;
    }

    // ---- one process, several GPUs (C1 / C5 predict shards over test points, the factorisation is "replicas only") ----
    // Copy the trained state of `src` (training points, L, the inverses of its diagonal blocks, alpha) to `dst`, a VanillaGaussianProcess
    // on another context / device: cudaMemcpyPeerAsync, i.e. NVLink when peer access is available (2 GiB of L at n = 16384 instead of a
    // redundant 62 ms factorisation per device).  K is not copied (predict does not read it): GetKtrain() stays with `src`.
    template<typename T>
    static int
    VanillaReplicate(Vanilla<T> *src, Vanilla<T> *dst) {
        if (src == nullptr || dst == nullptr || src == dst) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *sc = src->ctx, *dc = dst->ctx;
        if (!src->trained) { return SetError(sc, ERL_GP_STATUS_NOT_TRAINED, "vanilla: replicate before Train()"); }
        dst->trained = false;
        const long n = src->n, num_panels = CeilDiv(n, kPanel);
        ERL_GP_CUDA_OK(dc, cudaSetDevice(dc->device));
        if (dc->device != sc->device) {
            int can = 0;
            ERL_GP_CUDA_OK(dc, cudaDeviceCanAccessPeer(&can, dc->device, sc->device));
            if (can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(sc->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ERL_GP_CUDA_OK(dc, e); }
                (void) cudaGetLastError();
            }
        }
        ERL_GP_CUDA_OK(dc, dst->x.Reserve(static_cast<size_t>(n) * src->x_dim));
        ERL_GP_CUDA_OK(dc, dst->alpha.Reserve(static_cast<size_t>(n) * src->y_dim));
        ERL_GP_CUDA_OK(dc, dst->l.Reserve(static_cast<size_t>(n) * n));
        ERL_GP_CUDA_OK(dc, dst->linv.Reserve(static_cast<size_t>(num_panels) * kPanel * kPanel));
        ERL_GP_CUDA_OK(dc, dst->s_buf.Reserve(static_cast<size_t>(kPanel) * (src->y_dim > 8192 ? src->y_dim : 8192)));
        ERL_GP_CUDA_OK(dc, dst->info.Reserve(1));
        ERL_GP_CUDA_OK(sc, cudaSetDevice(sc->device));
        ERL_GP_CUDA_OK(sc, cudaStreamSynchronize(sc->stream));  // the training of `src` is complete
        ERL_GP_CUDA_OK(dc, cudaSetDevice(dc->device));
        auto copy = [&](void *d, const void *s_, size_t bytes) { return cudaMemcpyPeerAsync(d, dc->device, s_, sc->device, bytes, dc->stream); };
        ERL_GP_CUDA_OK(dc, copy(dst->x.ptr, src->x.ptr, sizeof(T) * n * src->x_dim));
        ERL_GP_CUDA_OK(dc, copy(dst->alpha.ptr, src->alpha.ptr, sizeof(T) * n * src->y_dim));
        ERL_GP_CUDA_OK(dc, copy(dst->l.ptr, src->l.ptr, sizeof(T) * n * n));
        ERL_GP_CUDA_OK(dc, copy(dst->linv.ptr, src->linv.ptr, sizeof(T) * num_panels * kPanel * kPanel));
        ERL_GP_CUDA_OK(dc, copy(dst->info.ptr, src->info.ptr, sizeof(int)));
        ERL_GP_CUDA_OK(dc, cudaStreamSynchronize(dc->stream));
        dst->n = n, dst->x_dim = src->x_dim, dst->y_dim = src->y_dim, dst->kernel = src->kernel, dst->scale = src->scale;
        dst->trained = true;
        return ERL_GP_STATUS_OK;
    }

    // Test() of one trained GP replicated on several devices: contiguous ranges of the test points, one host thread per replica, every
    // replica writes its range straight into the caller's arrays (the host gather).
    template<typename T>
    static int
    VanillaTestMulti(Vanilla<T> *const *gps, long num_gps, long num_test, const T *x_test, long ld_xt, T *mean, T *var) {
        if (gps == nullptr || num_gps <= 0 || x_test == nullptr || num_test <= 0) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        for (long i = 0; i < num_gps; ++i) {
            if (gps[i] == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        }
        std::vector<int> status(static_cast<size_t>(num_gps), ERL_GP_STATUS_OK);
        auto run = [&](const long i) {
            const long t0 = num_test * i / num_gps, t1 = num_test * (i + 1) / num_gps;
            if (t1 <= t0) { return; }
            status[i] = VanillaTest<T>(gps[i], t1 - t0, x_test + t0 * ld_xt, ld_xt, mean != nullptr ? mean + t0 : nullptr, var != nullptr ? var + t0 : nullptr, cudaMemcpyHostToDevice,
                                       cudaMemcpyDeviceToHost, num_test);
            if (status[i] == ERL_GP_STATUS_OK) {
                const cudaError_t e = cudaStreamSynchronize(gps[i]->ctx->stream);
                if (e != cudaSuccess) { status[i] = SetError(gps[i]->ctx, ERL_GP_STATUS_CUDA_ERROR, "vanilla multi: %s", cudaGetErrorString(e)); }
            }
        };
        std::vector<std::thread> workers;
        for (long i = 1; i < num_gps; ++i) { workers.emplace_back(run, i); }
        run(0);
        for (std::thread &w : workers) { w.join(); }
        for (long i = 0; i < num_gps; ++i) {
            if (status[i] != ERL_GP_STATUS_OK) { return status[i]; }
        }
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    static int
    VanillaGet(Vanilla<T> *gp, T *k, long ld_k, T *l, long ld_l, T *alpha, long ld_a) {
        if (gp == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        if (!gp->trained) { return SetError(ctx, ERL_GP_STATUS_NOT_TRAINED, "vanilla: not trained"); }
        const long n = gp->n;
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        if (k != nullptr) {
            if (gp->k.capacity < static_cast<size_t>(n) * n) { return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "vanilla: this GP is a replica (erl_gp_vanilla_replicate): Ktrain stays with the GP it was trained on"); }
            ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(k, sizeof(T) * ld_k, gp->k.ptr, sizeof(T) * n, sizeof(T) * n, n, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (l != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(l, sizeof(T) * ld_l, gp->l.ptr, sizeof(T) * n, sizeof(T) * n, n, cudaMemcpyDeviceToHost, ctx->stream)); }
        if (alpha != nullptr) {
            ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(alpha, sizeof(T) * ld_a, gp->alpha.ptr, sizeof(T) * n, sizeof(T) * n, gp->y_dim, cudaMemcpyDeviceToHost, ctx->stream));
        }
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    static int
    ReadInfo(Context *ctx, const int *d_info, int *info) {
        int h = 0;
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(&h, d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        if (info != nullptr) { *info = h; }
        return ERL_GP_STATUS_OK;
    }

    // =========================================================================================
    // SPGP (dense)
    // =========================================================================================
    template<typename T>
    __global__ void
    SpgpScaleKernel(const long m, const long n, const T *__restrict__ k_mn, const T *__restrict__ sumsq, const T *__restrict__ var, T *__restrict__ k_s) {
        // Ks[:, i] = K_MN[:, i] / (lambda_i + var_i), lambda_i = 1 - ||beta_i||^2 — src/sparse_pseudo_input_gp.cpp:768-774
        const long r = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
        const long i = blockIdx.y;
        if (r < m) {
            const T lambda = T(1) - sumsq[i];
            k_s[r + i * m] = k_mn[r + i * m] * (T(1) / (lambda + var[i]));
        }
    }

    // diagonal-Q_M mode (src/sparse_pseudo_input_gp.cpp:775-776): q[r] += sum_i Ks[r][i] K_MN[r][i]
    template<typename T>
    __global__ void
    SpgpDiagUpdateKernel(const long m, const long n, const T *__restrict__ k_s, const T *__restrict__ k_mn, T *__restrict__ q) {
        const long r = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
        if (r >= m) { return; }
        T s = 0;
        for (long i = 0; i < n; ++i) { s += k_s[r + i * m] * k_mn[r + i * m]; }
        q[r] += s;
    }

    // out[r] = fill (a == nullptr) or a[r] / q[r] (TestResult ctor in diagonal mode, :100-101)
    template<typename T>
    __global__ void
    SpgpDiagSolveKernel(const long m, const T *__restrict__ a, const T *__restrict__ q, const T fill, T *__restrict__ out) {
        const long r = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
        if (r < m) { out[r] = a != nullptr ? a[r] / q[r] : fill; }
    }

    template<typename T>
    struct Spgp {
        Context *ctx = nullptr;
        long m = 0, x_dim = 0;
        int kernel = 0;
        T scale = T(1);
        bool l_qm_updated = false;
        bool diagonal_qm = false;  // Setting::diagonal_qm (sparse_pseudo_input_gp.hpp:57): Q_M kept as its diagonal only
        DeviceBuffer<T> q_diag;
        DeviceBuffer<T> z, k_m, l_km, linv_km, q_m, l_qm, linv_qm, alpha, alpha_solved, panel;
        DeviceBuffer<int> info;
        DeviceBuffer<T> x, y, var, k_mn, k_s, w, s_buf, sumsq, sumsq2, xt, mean, variance;
    };

    template<typename T>
    static int
    SpgpCreate(erl_gp_context *c, int kernel, T scale, long x_dim, long m, const T *pseudo, Spgp<T> **out) {
        Context *ctx = Ctx(c);
        if (ctx == nullptr || pseudo == nullptr || out == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        *out = nullptr;
        if (m <= 0) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "spgp: pseudo_points must have at least one column"); }  // :319
        if (x_dim < 1 || x_dim > 3) { return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "spgp: x_dim=%ld (supported: 1, 2, 3)", x_dim); }
        if (kernel < ERL_GP_KERNEL_OU || kernel > ERL_GP_KERNEL_RBF) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "spgp: unknown kernel %d", kernel); }
        auto *gp = new (std::nothrow) Spgp<T>();
        if (gp == nullptr) { return ERL_GP_STATUS_ALLOC_FAILED; }
        gp->ctx = ctx, gp->m = m, gp->x_dim = x_dim, gp->kernel = kernel, gp->scale = scale;
        const size_t mm = static_cast<size_t>(m) * m;
        const size_t lin = static_cast<size_t>(CeilDiv(m, kPanel)) * kPanel * kPanel;
        cudaError_t err = cudaSetDevice(ctx->device);
        if (err == cudaSuccess) { err = gp->z.Reserve(static_cast<size_t>(m) * x_dim); }
        if (err == cudaSuccess) { err = gp->k_m.Reserve(mm); }
        if (err == cudaSuccess) { err = gp->l_km.Reserve(mm); }
        if (err == cudaSuccess) { err = gp->q_m.Reserve(mm); }
        if (err == cudaSuccess) { err = gp->l_qm.Reserve(mm); }
        if (err == cudaSuccess) { err = gp->linv_km.Reserve(lin); }
        if (err == cudaSuccess) { err = gp->linv_qm.Reserve(lin); }
        if (err == cudaSuccess) { err = gp->alpha.Reserve(m); }
        if (err == cudaSuccess) { err = gp->alpha_solved.Reserve(m); }
        if (err == cudaSuccess) { err = gp->panel.Reserve(static_cast<size_t>(m) * kPanel); }
        if (err == cudaSuccess) { err = gp->info.Reserve(2); }
        if (err != cudaSuccess) {
            delete gp;
            return SetError(ctx, ERL_GP_STATUS_ALLOC_FAILED, "spgp: %s", cudaGetErrorString(err));
        }
        int rc = ERL_GP_STATUS_OK;
        auto fail = [&](int code) {
            delete gp;
            return code;
        };
        if (cudaMemcpyAsync(gp->z.ptr, pseudo, sizeof(T) * m * x_dim, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { return fail(ERL_GP_STATUS_CUDA_ERROR); }
        // K_M = ComputeKtest(Z, Z): no noise on the diagonal (:340); L_KM = chol(K_M) (:341); Q_M = K_M (:349); alpha = 0
        rc = LaunchKtest<T>(ctx, kernel, scale, x_dim, gp->z.ptr, x_dim, m, gp->z.ptr, x_dim, m, gp->k_m.ptr, m);
        if (rc != ERL_GP_STATUS_OK) { return fail(rc); }
        rc = CopyLower<T>(ctx, m, gp->k_m.ptr, m, gp->l_km.ptr, m);
        if (rc != ERL_GP_STATUS_OK) { return fail(rc); }
        rc = Potrf<T>(ctx, m, gp->l_km.ptr, m, gp->linv_km.ptr, gp->panel.ptr, gp->info.ptr);
        if (rc != ERL_GP_STATUS_OK) { return fail(rc); }
        if (cudaMemcpyAsync(gp->q_m.ptr, gp->k_m.ptr, sizeof(T) * mm, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess ||
            cudaMemsetAsync(gp->alpha.ptr, 0, sizeof(T) * m, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            return fail(SetError(ctx, ERL_GP_STATUS_CUDA_ERROR, "spgp: %s", cudaGetErrorString(cudaGetLastError())));
        }
        *out = gp;
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    static int
    SpgpUpdate(Spgp<T> *gp, long n, const T *x, long ld_x, const T *y, const T *var) {
        if (gp == nullptr || x == nullptr || y == nullptr || var == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        if (n <= 0) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "spgp: no training samples"); }  // :755
        const long m = gp->m, d = gp->x_dim;
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        ERL_GP_CUDA_OK(ctx, gp->x.Reserve(static_cast<size_t>(n) * d));
        ERL_GP_CUDA_OK(ctx, gp->y.Reserve(n));
        ERL_GP_CUDA_OK(ctx, gp->var.Reserve(n));
        ERL_GP_CUDA_OK(ctx, gp->k_mn.Reserve(static_cast<size_t>(m) * n));
        ERL_GP_CUDA_OK(ctx, gp->k_s.Reserve(static_cast<size_t>(m) * n));
        ERL_GP_CUDA_OK(ctx, gp->w.Reserve(static_cast<size_t>(m) * n));
        ERL_GP_CUDA_OK(ctx, gp->s_buf.Reserve(static_cast<size_t>(kPanel) * n));
        ERL_GP_CUDA_OK(ctx, gp->sumsq.Reserve(n));
        ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(gp->x.ptr, sizeof(T) * d, x, sizeof(T) * ld_x, sizeof(T) * d, n, cudaMemcpyDefault, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->y.ptr, y, sizeof(T) * n, cudaMemcpyDefault, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->var.ptr, var, sizeof(T) * n, cudaMemcpyDefault, ctx->stream));
        // K_MN (:759-762), kept; the triangular solve runs on a copy
        int rc = LaunchKtest<T>(ctx, gp->kernel, gp->scale, d, gp->z.ptr, d, m, gp->x.ptr, d, n, gp->k_mn.ptr, m);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->w.ptr, gp->k_mn.ptr, sizeof(T) * m * n, cudaMemcpyDeviceToDevice, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaMemsetAsync(gp->sumsq.ptr, 0, sizeof(T) * n, ctx->stream));
        rc = TrsmLower<T>(ctx, m, n, gp->l_km.ptr, m, gp->linv_km.ptr, gp->w.ptr, m, gp->s_buf.ptr, gp->sumsq.ptr, false);  // ||L_KM^-1 k_n||^2
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        const dim3 grid(static_cast<unsigned>(CeilDiv(m, 256)), static_cast<unsigned>(n));
        SpgpScaleKernel<T><<<grid, 256, 0, ctx->stream>>>(m, n, gp->k_mn.ptr, gp->sumsq.ptr, gp->var.ptr, gp->k_s.ptr);
        ctx->launches += 1;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        // Q_M += Ks K_MN^T (:778), or its diagonal only (:775-776); alpha += Ks y (:780)
        if (gp->diagonal_qm) {
            SpgpDiagUpdateKernel<T><<<static_cast<unsigned>(CeilDiv(m, 128)), 128, 0, ctx->stream>>>(m, n, gp->k_s.ptr, gp->k_mn.ptr, gp->q_diag.ptr);
            ctx->launches += 1;
            ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        } else {
            rc = Gemm<T>(ctx, kOpN, kOpT, m, m, n, T(1), gp->k_s.ptr, m, gp->k_mn.ptr, m, T(1), gp->q_m.ptr, m, false);
            if (rc != ERL_GP_STATUS_OK) { return rc; }
        }
        rc = Gemm<T>(ctx, kOpN, kOpN, m, 1, n, T(1), gp->k_s.ptr, m, gp->y.ptr, n, T(1), gp->alpha.ptr, m, false);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        gp->l_qm_updated = false;
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    static int
    SpgpPrepareLqm(Spgp<T> *gp) {  // :835-842 + TestResult ctor :100-106
        if (gp->l_qm_updated) { return ERL_GP_STATUS_OK; }
        Context *ctx = gp->ctx;
        const long m = gp->m;
        if (gp->diagonal_qm) {  // no L_QM (:839); alpha / diag(Q_M) (:100-101)
            SpgpDiagSolveKernel<T><<<static_cast<unsigned>(CeilDiv(m, 128)), 128, 0, ctx->stream>>>(m, gp->alpha.ptr, gp->q_diag.ptr, T(0), gp->alpha_solved.ptr);
            ctx->launches += 1;
            ERL_GP_CUDA_OK(ctx, cudaGetLastError());
            gp->l_qm_updated = true;
            return ERL_GP_STATUS_OK;
        }
        int rc = CopyLower<T>(ctx, m, gp->q_m.ptr, m, gp->l_qm.ptr, m);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        rc = Potrf<T>(ctx, m, gp->l_qm.ptr, m, gp->linv_qm.ptr, gp->panel.ptr, gp->info.ptr + 1);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        ERL_GP_CUDA_OK(ctx, gp->s_buf.Reserve(static_cast<size_t>(kPanel) * 8192));
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->alpha_solved.ptr, gp->alpha.ptr, sizeof(T) * m, cudaMemcpyDeviceToDevice, ctx->stream));
        rc = TrsmLower<T>(ctx, m, 1, gp->l_qm.ptr, m, gp->linv_qm.ptr, gp->alpha_solved.ptr, m, gp->s_buf.ptr, nullptr, true);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        rc = TrsmLowerTrans<T>(ctx, m, 1, gp->l_qm.ptr, m, gp->linv_qm.ptr, gp->alpha_solved.ptr, m, gp->s_buf.ptr);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        gp->l_qm_updated = true;
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    static int
    SpgpTest(Spgp<T> *gp, long num_test, const T *x_test, long ld_xt, T *mean, T *var) {
        if (gp == nullptr || x_test == nullptr || num_test <= 0) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        if (gp->diagonal_qm && var != nullptr) {
            return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "spgp: no variance in diagonal_qm mode (the reference never builds the L_QM its variance solves with, src/sparse_pseudo_input_gp.cpp:304-310, 839)");
        }
        int rc = SpgpPrepareLqm(gp);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        const long m = gp->m, d = gp->x_dim;
        // Fused path (erl_gp_predict_dense.cu), as in VanillaTest: Ktest(Z, x*) is generated where it is consumed - once for
        // the mean (k*^T Q_M^-1 alpha, :150-162) and inside each of the two left-looking solves of the variance
        //   var = 1 - ||L_KM^-1 kt||^2 + ||L_QM^-1 kt||^2   (:288-292, :304-309)
        // instead of a materialised M x tile Ktest, two GEMM-based TRSMs and a regenerated Ktest in between (same 7.5 ms per
        // 10 000 points at M = 2048 - 79 tiles of 128 points do not fill the GPU - but no M x T buffer and 27 TFLOP/s for large T).
        const long chunk = std::min<long>(num_test, 1L << 20);
        ERL_GP_CUDA_OK(ctx, gp->xt.Reserve(static_cast<size_t>(chunk) * d));
        if (mean != nullptr) { ERL_GP_CUDA_OK(ctx, gp->mean.Reserve(chunk)); }
        if (var != nullptr) {
            ERL_GP_CUDA_OK(ctx, gp->sumsq.Reserve(chunk));
            ERL_GP_CUDA_OK(ctx, gp->sumsq2.Reserve(chunk));
            ERL_GP_CUDA_OK(ctx, gp->variance.Reserve(chunk));
            ERL_GP_CUDA_OK(ctx, gp->w.Reserve(PredictVarianceSlabElems<T>(ctx, m, chunk)));
        }
        for (long t0 = 0; t0 < num_test; t0 += chunk) {
            const long tt = std::min(chunk, num_test - t0);
            ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(gp->xt.ptr, sizeof(T) * d, x_test + t0 * ld_xt, sizeof(T) * ld_xt, sizeof(T) * d, tt, cudaMemcpyDefault, ctx->stream));
            if (mean != nullptr) {
                rc = PredictMean<T>(ctx, gp->kernel, gp->scale, d, m, tt, gp->z.ptr, gp->xt.ptr, gp->alpha_solved.ptr, m, 1, gp->mean.ptr, tt);
                if (rc != ERL_GP_STATUS_OK) { return rc; }
                ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(mean + t0, gp->mean.ptr, sizeof(T) * tt, cudaMemcpyDefault, ctx->stream));
            }
            if (var != nullptr) {
                rc = PredictVariance<T>(ctx, gp->kernel, gp->scale, d, m, tt, gp->z.ptr, gp->xt.ptr, gp->l_km.ptr, m, gp->linv_km.ptr, gp->w.ptr, gp->sumsq.ptr);
                if (rc != ERL_GP_STATUS_OK) { return rc; }
                rc = PredictVariance<T>(ctx, gp->kernel, gp->scale, d, m, tt, gp->z.ptr, gp->xt.ptr, gp->l_qm.ptr, m, gp->linv_qm.ptr, gp->w.ptr, gp->sumsq2.ptr);
                if (rc != ERL_GP_STATUS_OK) { return rc; }
                rc = VarianceFinalize<T>(ctx, tt, gp->sumsq.ptr, gp->sumsq2.ptr, gp->variance.ptr);
                if (rc != ERL_GP_STATUS_OK) { return rc; }
                ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(var + t0, gp->variance.ptr, sizeof(T) * tt, cudaMemcpyDefault, ctx->stream));
            }
        }
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    // Gradient of the SPGP predictive mean, TestResult::GetGradient (src/sparse_pseudo_input_gp.cpp:187-278): the reference builds the
    // gradient columns of Ktest (ComputeKtestWithGradient with no gradient observation among the pseudo-points, :82-91) and dots them
    // with alpha.  Fused here: one thread per test point accumulates sum_j alpha_j dk(z_j, x*) / dx*_a over the pseudo-points staged in
    // shared memory - no M x T(d + 1) Ktest.   RBF: dk/dx*_a = (z - x*)_a k / l^2;   Matern32: 3 / l^2 exp(-sqrt(3) r / l) (z - x*)_a.
    template<typename T>
    __global__ void
    SpgpGradientKernel(const int type, const T scale, const int x_dim, const long m, const T *__restrict__ z, const T *__restrict__ alpha, const long t, const T *__restrict__ xt,
                       T *__restrict__ grad) {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        T *zs = reinterpret_cast<T *>(smem_raw);  // [256][4]: point + alpha
        const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
        T x[3] = {0, 0, 0}, acc[3] = {0, 0, 0};
        if (i < t) {
            for (int a = 0; a < x_dim; ++a) { x[a] = xt[i * x_dim + a]; }
        }
        const T l2 = scale * scale;
        const T c = sqrt(T(3)) / scale;
        for (long j0 = 0; j0 < m; j0 += blockDim.x) {
            const long j = j0 + threadIdx.x;
            __syncthreads();
            for (int a = 0; a < 3; ++a) { zs[4 * threadIdx.x + a] = (j < m && a < x_dim) ? z[j * x_dim + a] : T(0); }
            zs[4 * threadIdx.x + 3] = j < m ? alpha[j] : T(0);
            __syncthreads();
            const int cnt = static_cast<int>(m - j0 < blockDim.x ? m - j0 : blockDim.x);
            for (int k = 0; k < cnt; ++k) {
                T diff[3], r2 = 0;
                for (int a = 0; a < 3; ++a) {
                    diff[a] = zs[4 * k + a] - x[a];
                    r2 += diff[a] * diff[a];
                }
                const T w = zs[4 * k + 3] * (type == ERL_GP_KERNEL_RBF ? exp(-r2 / (T(2) * l2)) / l2 : c * c * exp(-c * sqrt(r2)));
                for (int a = 0; a < 3; ++a) { acc[a] += w * diff[a]; }
            }
        }
        if (i < t) {
            for (int a = 0; a < x_dim; ++a) { grad[a + i * x_dim] = acc[a]; }
        }
    }

    // grad: x_dim x num_test (host).  raw_alpha != 0 reproduces the reference's batched accessor, which dots with the UNSOLVED alpha
    // (m_gp_->m_mat_alpha_, :212) while its per-index accessor (:252) and GetMean use Q_M^-1 alpha; the default is the consistent one.
    template<typename T>
    static int
    SpgpTestGradient(Spgp<T> *gp, long num_test, const T *x_test, long ld_xt, T *grad, int raw_alpha) {
        if (gp == nullptr || x_test == nullptr || grad == nullptr || num_test <= 0) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        if (gp->kernel == ERL_GP_KERNEL_OU) { return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "spgp: OrnsteinUhlenbeck has no gradient"); }
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        int rc = SpgpPrepareLqm(gp);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        const long m = gp->m, d = gp->x_dim;
        ERL_GP_CUDA_OK(ctx, gp->xt.Reserve(static_cast<size_t>(num_test) * d));
        ERL_GP_CUDA_OK(ctx, gp->w.Reserve(static_cast<size_t>(num_test) * d));
        ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(gp->xt.ptr, sizeof(T) * d, x_test, sizeof(T) * ld_xt, sizeof(T) * d, num_test, cudaMemcpyDefault, ctx->stream));
        SpgpGradientKernel<T><<<static_cast<unsigned>(CeilDiv(num_test, 256)), 256, 256 * 4 * sizeof(T), ctx->stream>>>(gp->kernel, gp->scale, static_cast<int>(d), m, gp->z.ptr,
                                                                                                                     raw_alpha ? gp->alpha.ptr : gp->alpha_solved.ptr, num_test, gp->xt.ptr, gp->w.ptr);
        ctx->launches += 1;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(grad, gp->w.ptr, sizeof(T) * num_test * d, cudaMemcpyDefault, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    // Setting::diagonal_qm (ctor :346-347: Q_M = ones(M)); call before the first update: Q_M and alpha are reset
    template<typename T>
    static int
    SpgpSetDiagonalQm(Spgp<T> *gp, int on) {
        if (gp == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        const long m = gp->m;
        gp->diagonal_qm = on != 0;
        gp->l_qm_updated = false;
        ERL_GP_CUDA_OK(ctx, cudaMemsetAsync(gp->alpha.ptr, 0, sizeof(T) * m, ctx->stream));
        if (on) {
            ERL_GP_CUDA_OK(ctx, gp->q_diag.Reserve(m));
            SpgpDiagSolveKernel<T><<<static_cast<unsigned>(CeilDiv(m, 128)), 128, 0, ctx->stream>>>(m, nullptr, nullptr, T(1), gp->q_diag.ptr);
            ctx->launches += 1;
            ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        } else {
            ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->q_m.ptr, gp->k_m.ptr, sizeof(T) * m * m, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    static int
    SpgpGetQmDiagonal(Spgp<T> *gp, T *q) {
        if (gp == nullptr || q == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        if (!gp->diagonal_qm) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "spgp: not in diagonal_qm mode"); }
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(q, gp->q_diag.ptr, sizeof(T) * gp->m, cudaMemcpyDeviceToHost, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    // Read() of the reference restores Q_M and alpha from the stream (src/sparse_pseudo_input_gp.cpp:721-740); K_M and L_KM follow
    // from the pseudo-points at construction, L_QM is refactored lazily.  q_m: M x M col-major, or M values in diagonal_qm mode.
    template<typename T>
    static int
    SpgpSetState(Spgp<T> *gp, const T *q_m, const T *alpha) {
        if (gp == nullptr || q_m == nullptr || alpha == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        if (gp->diagonal_qm) {
            ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->q_diag.ptr, q_m, sizeof(T) * gp->m, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->q_m.ptr, q_m, sizeof(T) * gp->m * gp->m, cudaMemcpyHostToDevice, ctx->stream));
        }
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->alpha.ptr, alpha, sizeof(T) * gp->m, cudaMemcpyHostToDevice, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));  // the host buffers may go away
        gp->l_qm_updated = false;
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    static int
    SpgpGet(Spgp<T> *gp, T *q_m, T *alpha, T *l_km, T *l_qm) {
        if (gp == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        const size_t mm = static_cast<size_t>(gp->m) * gp->m;
        if (l_qm != nullptr) {
            const int rc = SpgpPrepareLqm(gp);
            if (rc != ERL_GP_STATUS_OK) { return rc; }
            ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(l_qm, gp->l_qm.ptr, sizeof(T) * mm, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (q_m != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(q_m, gp->q_m.ptr, sizeof(T) * mm, cudaMemcpyDeviceToHost, ctx->stream)); }
        if (alpha != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(alpha, gp->alpha.ptr, sizeof(T) * gp->m, cudaMemcpyDeviceToHost, ctx->stream)); }
        if (l_km != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(l_km, gp->l_km.ptr, sizeof(T) * mm, cudaMemcpyDeviceToHost, ctx->stream)); }
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

}  // namespace erl_gp

using namespace erl_gp;

struct erl_gp_vanilla_f32 : Vanilla<float> {};
struct erl_gp_vanilla_f64 : Vanilla<double> {};
struct erl_gp_spgp_f32 : Spgp<float> {};
struct erl_gp_spgp_f64 : Spgp<double> {};

extern "C" {

#define ERL_GP_DEFINE_DENSE(T, SFX)                                                                                                                                          \
    int erl_gp_vanilla_create_##SFX(erl_gp_context *ctx, erl_gp_vanilla_##SFX **gp) {                                                                                        \
        if (ctx == nullptr || gp == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }                                                                                      \
        auto *v = new (std::nothrow) Vanilla<T>();                                                                                                                           \
        if (v == nullptr) { return ERL_GP_STATUS_ALLOC_FAILED; }                                                                                                             \
        v->ctx = Ctx(ctx);                                                                                                                                                   \
        *gp = static_cast<erl_gp_vanilla_##SFX *>(v);                                                                                                                        \
        return ERL_GP_STATUS_OK;                                                                                                                                             \
    }                                                                                                                                                                        \
    int erl_gp_vanilla_destroy_##SFX(erl_gp_vanilla_##SFX *gp) {                                                                                                             \
        if (gp != nullptr) {                                                                                                                                                 \
            cudaSetDevice(gp->ctx->device);                                                                                                                                  \
            cudaStreamSynchronize(gp->ctx->stream);                                                                                                                          \
            delete static_cast<Vanilla<T> *>(gp);                                                                                                                            \
        }                                                                                                                                                                    \
        return ERL_GP_STATUS_OK;                                                                                                                                             \
    }                                                                                                                                                                        \
    int erl_gp_vanilla_train_##SFX(erl_gp_vanilla_##SFX *gp, int kernel, T scale, long x_dim, long y_dim, long n, const T *x, long ld_x, const T *y, long ld_y,              \
                                   const T *var, int *info) {                                                                                                                \
        const int rc = VanillaTrainDev<T>(gp, kernel, scale, x_dim, y_dim, n, x, ld_x, y, ld_y, var, cudaMemcpyHostToDevice);                                                \
        if (rc != ERL_GP_STATUS_OK) { return rc; }                                                                                                                           \
        return ReadInfo<T>(gp->ctx, gp->info.ptr, info);                                                                                                                     \
    }                                                                                                                                                                        \
    int erl_gp_vanilla_train_dev_##SFX(erl_gp_vanilla_##SFX *gp, int kernel, T scale, long x_dim, long y_dim, long n, const T *x, long ld_x, const T *y, long ld_y,          \
                                       const T *var) {                                                                                                                       \
        return VanillaTrainDev<T>(gp, kernel, scale, x_dim, y_dim, n, x, ld_x, y, ld_y, var, cudaMemcpyDeviceToDevice);                                                      \
    }                                                                                                                                                                        \
    int erl_gp_vanilla_info_##SFX(erl_gp_vanilla_##SFX *gp, int *info) {                                                                                                     \
        if (gp == nullptr || info == nullptr || gp->info.ptr == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }                                                          \
        return ReadInfo<T>(gp->ctx, gp->info.ptr, info);                                                                                                                     \
    }                                                                                                                                                                        \
    int erl_gp_vanilla_get_##SFX(erl_gp_vanilla_##SFX *gp, T *k, long ld_k, T *l, long ld_l, T *alpha, long ld_a) { return VanillaGet<T>(gp, k, ld_k, l, ld_l, alpha, ld_a); } \
    int erl_gp_vanilla_test_##SFX(erl_gp_vanilla_##SFX *gp, long num_test, const T *x_test, long ld_xt, T *mean, T *var) {                                                   \
        const int rc = VanillaTest<T>(gp, num_test, x_test, ld_xt, mean, var, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost);                                               \
        if (rc != ERL_GP_STATUS_OK) { return rc; }                                                                                                                           \
        return erl_gp_context_synchronize(gp->ctx);                                                                                                                          \
    }                                                                                                                                                                        \
    int erl_gp_vanilla_replicate_##SFX(erl_gp_vanilla_##SFX *src, erl_gp_vanilla_##SFX *dst) { return VanillaReplicate<T>(src, dst); }                                      \
    int erl_gp_vanilla_test_multi_##SFX(erl_gp_vanilla_##SFX *const *gps, long num_gps, long num_test, const T *x_test, long ld_xt, T *mean, T *var) {                        \
        return VanillaTestMulti<T>(reinterpret_cast<Vanilla<T> *const *>(gps), num_gps, num_test, x_test, ld_xt, mean, var);                                                 \
    }                                                                                                                                                                        \
    int erl_gp_vanilla_test_dev_##SFX(erl_gp_vanilla_##SFX *gp, long num_test, const T *x_test, long ld_xt, T *mean, T *var) {                                               \
        return VanillaTest<T>(gp, num_test, x_test, ld_xt, mean, var, cudaMemcpyDeviceToDevice, cudaMemcpyDeviceToDevice);                                                   \
    }                                                                                                                                                                        \
    int erl_gp_spgp_create_##SFX(erl_gp_context *ctx, int kernel, T scale, long x_dim, long num_pseudo, const T *pseudo_points, erl_gp_spgp_##SFX **gp) {                    \
        return SpgpCreate<T>(ctx, kernel, scale, x_dim, num_pseudo, pseudo_points, reinterpret_cast<Spgp<T> **>(gp));                                                        \
    }                                                                                                                                                                        \
    int erl_gp_spgp_destroy_##SFX(erl_gp_spgp_##SFX *gp) {                                                                                                                   \
        if (gp != nullptr) {                                                                                                                                                 \
            cudaSetDevice(gp->ctx->device);                                                                                                                                  \
            cudaStreamSynchronize(gp->ctx->stream);                                                                                                                          \
            delete static_cast<Spgp<T> *>(gp);                                                                                                                               \
        }                                                                                                                                                                    \
        return ERL_GP_STATUS_OK;                                                                                                                                             \
    }                                                                                                                                                                        \
    int erl_gp_spgp_update_##SFX(erl_gp_spgp_##SFX *gp, long n, const T *x, long ld_x, const T *y, const T *var) { return SpgpUpdate<T>(gp, n, x, ld_x, y, var); }           \
    int erl_gp_spgp_test_##SFX(erl_gp_spgp_##SFX *gp, long num_test, const T *x_test, long ld_xt, T *mean, T *var) {                                                         \
        return SpgpTest<T>(gp, num_test, x_test, ld_xt, mean, var);                                                                                                          \
    }                                                                                                                                                                        \
    int erl_gp_spgp_set_diagonal_qm_##SFX(erl_gp_spgp_##SFX *gp, int on) { return SpgpSetDiagonalQm<T>(gp, on); }                                                            \
    int erl_gp_spgp_get_qm_diagonal_##SFX(erl_gp_spgp_##SFX *gp, T *q) { return SpgpGetQmDiagonal<T>(gp, q); }                                                               \
    int erl_gp_spgp_test_gradient_##SFX(erl_gp_spgp_##SFX *gp, long num_test, const T *x_test, long ld_xt, T *grad, int raw_alpha) {                                         \
        return SpgpTestGradient<T>(gp, num_test, x_test, ld_xt, grad, raw_alpha);                                                                                            \
    }                                                                                                                                                                        \
    int erl_gp_spgp_set_state_##SFX(erl_gp_spgp_##SFX *gp, const T *q_m, const T *alpha) { return SpgpSetState<T>(gp, q_m, alpha); }                                        \
    int erl_gp_spgp_get_##SFX(erl_gp_spgp_##SFX *gp, T *q_m, T *alpha, T *l_km, T *l_qm) { return SpgpGet<T>(gp, q_m, alpha, l_km, l_qm); }

ERL_GP_DEFINE_DENSE(float, f32)
ERL_GP_DEFINE_DENSE(double, f64)

}  // extern "C"
