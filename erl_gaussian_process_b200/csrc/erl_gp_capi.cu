// C ABI: context, covariance and batched small-GP entry points (see include/erl_gp_b200.h).
#include "erl_gp_internal.cuh"

#include <cstdlib>
#include <thread>
#include <vector>

namespace erl_gp {

    int
    SetError(Context *ctx, const int status, const char *fmt, ...) {
        if (ctx != nullptr) {
            va_list args;
            va_start(args, fmt);
            vsnprintf(ctx->last_error, sizeof(ctx->last_error), fmt, args);
            va_end(args);
        }
        return status;
    }

    template<typename T>
    int
    BatchCreate(erl_gp_context *c, long num_gps, long max_n, long x_dim, int kernel, T scale, Batch<T> **out) {
        Context *ctx = Ctx(c);
        if (ctx == nullptr || out == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        *out = nullptr;
        if (num_gps <= 0 || max_n <= 0) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "batch: num_gps=%ld max_n=%ld", num_gps, max_n); }
        if (x_dim < 1 || x_dim > 3) { return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "batch: x_dim=%ld (supported: 1, 2, 3)", x_dim); }
        if (max_n > LargeGpMaxN()) { return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "batch: max_n=%ld exceeds %ld", max_n, LargeGpMaxN()); }
        if (kernel < ERL_GP_KERNEL_OU || kernel > ERL_GP_KERNEL_RBF) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "batch: unknown kernel %d", kernel); }
        auto *b = new (std::nothrow) Batch<T>();
        if (b == nullptr) { return ERL_GP_STATUS_ALLOC_FAILED; }
        b->ctx = ctx;
        b->num_gps = num_gps;
        b->max_n = max_n;
        b->x_dim = x_dim;
        b->kernel = kernel;
        b->scale = scale;
        const size_t bn = static_cast<size_t>(num_gps) * max_n;
        cudaError_t err = cudaSetDevice(ctx->device);
        if (err == cudaSuccess) { err = b->n_train.Reserve(num_gps); }
        if (err == cudaSuccess) { err = b->info.Reserve(num_gps); }
        if (err == cudaSuccess) { err = b->x.Reserve(bn * x_dim); }
        if (err == cudaSuccess) { err = b->y.Reserve(bn); }
        if (err == cudaSuccess) { err = b->var.Reserve(bn); }
        if (err == cudaSuccess) { err = b->alpha.Reserve(bn); }
        if (err == cudaSuccess) { err = b->l.Reserve(bn * max_n); }
        if (err == cudaSuccess) { err = cudaMemsetAsync(b->info.ptr, 0xff, sizeof(int) * num_gps, ctx->stream); }  // -1 = untrained
        // the kernels write rows / columns [0, n) of a GP's alpha / L slice only: start from zeros so that the padding a caller downloads
        // does not depend on what the allocator handed out
        if (err == cudaSuccess) { err = cudaMemsetAsync(b->alpha.ptr, 0, sizeof(T) * bn, ctx->stream); }
        if (err == cudaSuccess) { err = cudaMemsetAsync(b->l.ptr, 0, sizeof(T) * bn * max_n, ctx->stream); }
        if (err != cudaSuccess) {
            delete b;
            return SetError(ctx, ERL_GP_STATUS_ALLOC_FAILED, "batch: %s", cudaGetErrorString(err));
        }
        *out = b;
        return ERL_GP_STATUS_OK;
    }

    // O(num_gps) host checks of caller metadata: a training-set size beyond the capacity would index past the per-GP slices
    // (and the kernel's shared memory), an inconsistent CSR query list past the query / output buffers
    static bool
    TrainSizesOk(const int *n_train, long num_gps, long max_n) {
        for (long g = 0; g < num_gps; ++g) {
            if (n_train[g] < 0 || n_train[g] > max_n) { return false; }
        }
        return true;
    }

    static bool
    QueryOffsetsOk(const long *q_offsets, long num_gps, long num_q) {
        if (q_offsets[0] != 0 || q_offsets[num_gps] != num_q) { return false; }
        for (long g = 0; g < num_gps; ++g) {
            if (q_offsets[g + 1] < q_offsets[g]) { return false; }
        }
        return true;
    }

    template<typename T>
    int
    BatchUpload(Batch<T> *b, const int *n_train, const T *x, const T *y, const T *var) {
        if (b == nullptr || n_train == nullptr || x == nullptr || y == nullptr || var == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = b->ctx;
        if (!TrainSizesOk(n_train, b->num_gps, b->max_n)) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "batch: n_train[g] must lie in [0, max_n = %ld]", b->max_n); }
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        const size_t bn = static_cast<size_t>(b->num_gps) * b->max_n;
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(b->n_train.ptr, n_train, sizeof(int) * b->num_gps, cudaMemcpyHostToDevice, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(b->x.ptr, x, sizeof(T) * bn * b->x_dim, cudaMemcpyHostToDevice, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(b->y.ptr, y, sizeof(T) * bn, cudaMemcpyHostToDevice, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(b->var.ptr, var, sizeof(T) * bn, cudaMemcpyHostToDevice, ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    constexpr long kDefaultHostChunks = 8;  // chunks of the host-buffer pipeline (BatchTrainPredictHost)

    template<typename T>
    int
    BatchTrainDev(Batch<T> *b, long min_num_samples, int write_l) {
        if (b == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        ERL_GP_CUDA_OK(b->ctx, cudaSetDevice(b->ctx->device));
        const BatchParams<T> p = b->Params(min_num_samples, write_l);
        return LaunchBatch<T>(b->ctx, p, static_cast<int>(b->x_dim), kBatchTrain, 1);
    }

    template<typename T>
    int
    BatchPredictDev(Batch<T> *b, const long *q_offsets, const T *q_x, const int *q_out_index, long num_q, int mapping, T mapping_scale, T *mean, T *var, uint8_t *valid) {
        if (b == nullptr || q_offsets == nullptr || q_x == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        if (num_q <= 0) { return ERL_GP_STATUS_OK; }
        ERL_GP_CUDA_OK(b->ctx, cudaSetDevice(b->ctx->device));
        BatchParams<T> p = b->Params(0, 0);
        p.q_offsets = q_offsets;
        p.q_x = q_x;
        p.q_out_index = q_out_index;
        p.mean = mean;
        p.variance = var;
        p.valid = valid;
        p.mapping = mapping;
        p.mapping_scale = mapping_scale;
        // CTAs sharing one GP's query list: aim at >= 4 CTAs per SM over the whole grid
        const long tq = 64;
        long tiles = CeilDiv(CeilDiv(num_q, b->num_gps), tq);
        const long want = CeilDiv(4L * b->ctx->sm_count, b->num_gps);
        if (tiles > want) { tiles = want; }
        if (tiles < 1) { tiles = 1; }
        if (tiles > 65535) { tiles = 65535; }
        return LaunchBatch<T>(b->ctx, p, static_cast<int>(b->x_dim), kBatchPredict, static_cast<int>(tiles));
    }

    template<typename T>
    int
    BatchTrainPredictDev(Batch<T> *b, long min_num_samples, int write_l, const long *q_offsets, const T *q_x, long num_q, T *mean, T *var, uint8_t *valid) {
        if (b == nullptr || q_offsets == nullptr || (num_q > 0 && q_x == nullptr)) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        ERL_GP_CUDA_OK(b->ctx, cudaSetDevice(b->ctx->device));
        BatchParams<T> p = b->Params(min_num_samples, write_l);
        p.q_offsets = q_offsets;
        p.q_x = q_x;
        p.mean = mean;
        p.variance = var;
        p.valid = valid;
        return LaunchBatch<T>(b->ctx, p, static_cast<int>(b->x_dim), kBatchTrainPredict, 1);
    }

    template<typename T>
    int
    BatchDownload(Batch<T> *b, T *l, T *alpha, int *info) {
        if (b == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = b->ctx;
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        const size_t bn = static_cast<size_t>(b->num_gps) * b->max_n;
        if (l != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(l, b->l.ptr, sizeof(T) * bn * b->max_n, cudaMemcpyDeviceToHost, ctx->stream)); }
        if (alpha != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(alpha, b->alpha.ptr, sizeof(T) * bn, cudaMemcpyDeviceToHost, ctx->stream)); }
        if (info != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(info, b->info.ptr, sizeof(int) * b->num_gps, cudaMemcpyDeviceToHost, ctx->stream)); }
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    static int
    BatchTrainPredictHost(
        Batch<T> *b,
        long min_num_samples,
        const int *n_train,
        const T *x,
        const T *y,
        const T *var,
        const long *q_offsets,
        const T *q_x,
        long num_q,
        T *l,
        T *alpha,
        int *info,
        T *mean,
        T *variance,
        uint8_t *valid) {
        if (b == nullptr || q_offsets == nullptr || num_q < 0 || n_train == nullptr || x == nullptr || y == nullptr || var == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        if (num_q > 0 && q_x == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = b->ctx;
        if (!TrainSizesOk(n_train, b->num_gps, b->max_n)) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "batch: n_train[g] must lie in [0, max_n = %ld]", b->max_n); }
        if (!QueryOffsetsOk(q_offsets, b->num_gps, num_q)) {
            return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "batch: q_offsets must start at 0, be non-decreasing and end at num_q = %ld", num_q);
        }
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        const long num_gps = b->num_gps, max_n = b->max_n, d = b->x_dim;
        const size_t nq = static_cast<size_t>(num_q > 0 ? num_q : 1);
        ERL_GP_CUDA_OK(ctx, b->q_offsets.Reserve(num_gps + 1));
        ERL_GP_CUDA_OK(ctx, b->q_x.Reserve(nq * d));
        ERL_GP_CUDA_OK(ctx, b->mean.Reserve(nq));
        ERL_GP_CUDA_OK(ctx, b->variance.Reserve(nq));
        ERL_GP_CUDA_OK(ctx, b->valid.Reserve(nq));
        ERL_GP_CUDA_OK(ctx, b->info_host.Reserve(num_gps));
        // Pipeline: the GP stream is cut into chunks; chunk c+1 is uploaded (copy-in stream) and chunk c-1 downloaded
        // (copy-out stream) while chunk c is computed (the context's stream).  H2D, the kernel and D2H of a 50k-GP batch
        // cost 4.7 + 7.6 + 1.0 ms back to back; overlapped the step is bounded by the kernel.
        static const long env_chunks = std::getenv("ERL_GP_BATCH_CHUNKS") != nullptr ? std::atol(std::getenv("ERL_GP_BATCH_CHUNKS")) : 0;
        const long want_chunks = env_chunks > 0 ? env_chunks : kDefaultHostChunks;
        const long num_chunks = num_gps >= 1024 * want_chunks ? want_chunks : (num_gps >= 8 * 1024 ? 8 : 1);
        if (b->copy_in == nullptr) {
            ERL_GP_CUDA_OK(ctx, cudaStreamCreateWithFlags(&b->copy_in, cudaStreamNonBlocking));
            ERL_GP_CUDA_OK(ctx, cudaStreamCreateWithFlags(&b->copy_out, cudaStreamNonBlocking));
            ERL_GP_CUDA_OK(ctx, cudaStreamCreateWithFlags(&b->compute2, cudaStreamNonBlocking));
        }
        while (static_cast<long>(b->ev_in.size()) < num_chunks) {
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            ERL_GP_CUDA_OK(ctx, cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
            ERL_GP_CUDA_OK(ctx, cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
            b->ev_in.push_back(e0);
            b->ev_kernel.push_back(e1);
        }
        // the copy streams start after whatever is already queued on the compute stream (re-use of the device buffers)
        ERL_GP_CUDA_OK(ctx, cudaEventRecord(b->ev_kernel[0], ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaStreamWaitEvent(b->copy_in, b->ev_kernel[0], 0));
        ERL_GP_CUDA_OK(ctx, cudaStreamWaitEvent(b->copy_out, b->ev_kernel[0], 0));
        ERL_GP_CUDA_OK(ctx, cudaStreamWaitEvent(b->compute2, b->ev_kernel[0], 0));
        // Chunks alternate between two compute streams: kernels of one stream run back to back, so the last, partly filled wave of chunk
        // c would leave SMs idle until chunk c + 1 may start; on two streams the head of c + 1 fills them (the chunks are independent).
        static const bool one_stream = std::getenv("ERL_GP_BATCH_ONE_COMPUTE_STREAM") != nullptr;  // A/B
        cudaStream_t const main_stream = ctx->stream;
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(b->q_offsets.ptr, q_offsets, sizeof(long) * (num_gps + 1), cudaMemcpyHostToDevice, b->copy_in));
        // Uneven cut: a short first chunk (the first kernel starts after 1/4 of a nominal chunk's upload instead of a whole one)
        // and a short last chunk (a short download after the last kernel); the interior is cut evenly.
        const long edge = num_chunks >= 4 ? num_gps / (4 * num_chunks) : 0;
        auto chunk_begin = [&](long c) {
            if (edge == 0) { return c * num_gps / num_chunks; }
            if (c <= 0) { return 0L; }
            if (c >= num_chunks) { return num_gps; }
            return edge + (c - 1) * (num_gps - 2 * edge) / (num_chunks - 2);
        };
        // ---- enqueue every upload and every kernel ----
        for (long c = 0; c < num_chunks; ++c) {
            const long g0 = chunk_begin(c), g1 = chunk_begin(c + 1);
            const long t0 = q_offsets[g0], t1 = q_offsets[g1];
            const size_t gn = static_cast<size_t>(g1 - g0) * max_n;
            ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(b->n_train.ptr + g0, n_train + g0, sizeof(int) * (g1 - g0), cudaMemcpyHostToDevice, b->copy_in));
            ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(b->x.ptr + g0 * max_n * d, x + g0 * max_n * d, sizeof(T) * gn * d, cudaMemcpyHostToDevice, b->copy_in));
            ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(b->y.ptr + g0 * max_n, y + g0 * max_n, sizeof(T) * gn, cudaMemcpyHostToDevice, b->copy_in));
            ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(b->var.ptr + g0 * max_n, var + g0 * max_n, sizeof(T) * gn, cudaMemcpyHostToDevice, b->copy_in));
            if (t1 > t0) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(b->q_x.ptr + t0 * d, q_x + t0 * d, sizeof(T) * (t1 - t0) * d, cudaMemcpyHostToDevice, b->copy_in)); }
            ERL_GP_CUDA_OK(ctx, cudaEventRecord(b->ev_in[c], b->copy_in));
            cudaStream_t const cs = (!one_stream && (c & 1)) ? b->compute2 : main_stream;
            ERL_GP_CUDA_OK(ctx, cudaStreamWaitEvent(cs, b->ev_in[c], 0));
            if (t1 > t0) { ERL_GP_CUDA_OK(ctx, cudaMemsetAsync(b->valid.ptr + t0, 0, t1 - t0, cs)); }
            // L is always materialised in HBM (downloaded only on demand)
            BatchParams<T> p = b->Params(min_num_samples, 1);
            p.num_gps = static_cast<int>(g1 - g0);
            p.n_train += g0;
            p.x += g0 * max_n * d;
            p.y += g0 * max_n;
            p.var += g0 * max_n;
            p.l += g0 * max_n * max_n;
            p.alpha += g0 * max_n;
            p.info += g0;
            p.q_offsets = b->q_offsets.ptr + g0;  // offsets stay global: q_x / mean / variance / valid are not shifted
            p.q_x = b->q_x.ptr;
            p.mean = mean != nullptr ? b->mean.ptr : nullptr;
            p.variance = variance != nullptr ? b->variance.ptr : nullptr;
            p.valid = b->valid.ptr;
            ctx->stream = cs;  // LaunchBatch launches on the context's stream
            const int rc = LaunchBatch<T>(ctx, p, static_cast<int>(d), kBatchTrainPredict, 1);
            ctx->stream = main_stream;
            if (rc != ERL_GP_STATUS_OK) { return rc; }
            ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(b->info_host.ptr + g0, b->info.ptr + g0, sizeof(int) * (g1 - g0), cudaMemcpyDeviceToHost, cs));
            ERL_GP_CUDA_OK(ctx, cudaEventRecord(b->ev_kernel[c], cs));
        }
        // ---- downloads, chunk by chunk as the kernels finish ----
        std::vector<T> tmp;
        std::vector<uint8_t> tmp_valid;
        for (long c = 0; c < num_chunks; ++c) {
            const long g0 = chunk_begin(c), g1 = chunk_begin(c + 1);
            const long t0 = q_offsets[g0], t1 = q_offsets[g1];
            const size_t gn = static_cast<size_t>(g1 - g0) * max_n;
            ERL_GP_CUDA_OK(ctx, cudaEventSynchronize(b->ev_kernel[c]));
            bool all_trained = true;
            for (long g = g0; g < g1; ++g) { all_trained = all_trained && b->info_host.ptr[g] == 0; }
            if (info != nullptr) { std::memcpy(info + g0, b->info_host.ptr + g0, sizeof(int) * (g1 - g0)); }
            if (alpha != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(alpha + g0 * max_n, b->alpha.ptr + g0 * max_n, sizeof(T) * gn, cudaMemcpyDeviceToHost, b->copy_out)); }
            if (l != nullptr) {
                ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(l + g0 * max_n * max_n, b->l.ptr + g0 * max_n * max_n, sizeof(T) * gn * max_n, cudaMemcpyDeviceToHost, b->copy_out));
            }
            if (t1 <= t0) { continue; }
            if (valid != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(valid + t0, b->valid.ptr + t0, t1 - t0, cudaMemcpyDeviceToHost, b->copy_out)); }
            if (all_trained) {
                // every query of the chunk was written: straight into the caller's buffers
                if (mean != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(mean + t0, b->mean.ptr + t0, sizeof(T) * (t1 - t0), cudaMemcpyDeviceToHost, b->copy_out)); }
                if (variance != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(variance + t0, b->variance.ptr + t0, sizeof(T) * (t1 - t0), cudaMemcpyDeviceToHost, b->copy_out)); }
            } else {
                // outputs of untrained / failed GPs must come back untouched (src/lidar_gp_2d.cpp:112,120): merge on the host
                tmp.resize(static_cast<size_t>(t1 - t0));
                tmp_valid.resize(static_cast<size_t>(t1 - t0));
                ERL_GP_CUDA_OK(ctx, cudaMemcpy(tmp_valid.data(), b->valid.ptr + t0, t1 - t0, cudaMemcpyDeviceToHost));
                if (mean != nullptr) {
                    ERL_GP_CUDA_OK(ctx, cudaMemcpy(tmp.data(), b->mean.ptr + t0, sizeof(T) * (t1 - t0), cudaMemcpyDeviceToHost));
                    for (long q = 0; q < t1 - t0; ++q) {
                        if (tmp_valid[q]) { mean[t0 + q] = tmp[q]; }
                    }
                }
                if (variance != nullptr) {
                    ERL_GP_CUDA_OK(ctx, cudaMemcpy(tmp.data(), b->variance.ptr + t0, sizeof(T) * (t1 - t0), cudaMemcpyDeviceToHost));
                    for (long q = 0; q < t1 - t0; ++q) {
                        if (tmp_valid[q]) { variance[t0 + q] = tmp[q]; }
                    }
                }
            }
        }
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(b->copy_out));
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(b->compute2));
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    // One process, several GPUs (src/lidar_gp_2d.cpp:366-392: the partitions are independent, the reference loops over them with
    // OpenMP): batch i takes the next batches[i]->num_gps GPs of the caller's stream - contiguous GP ranges, no exchange between
    // devices.  One host thread per batch drives that device's upload / kernel / download pipeline; every device writes its
    // results straight into the caller's arrays, which is the whole "host gather".
    template<typename T>
    static int
    BatchTrainPredictMulti(
        Batch<T> *const *batches,
        long num_batches,
        long min_num_samples,
        const int *n_train,
        const T *x,
        const T *y,
        const T *var,
        const long *q_offsets,
        const T *q_x,
        long num_q,
        T *l,
        T *alpha,
        int *info,
        T *mean,
        T *variance,
        uint8_t *valid) {
        if (batches == nullptr || num_batches <= 0 || q_offsets == nullptr || n_train == nullptr || x == nullptr || y == nullptr || var == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        long total = 0;
        for (long i = 0; i < num_batches; ++i) {
            if (batches[i] == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
            if (batches[i]->max_n != batches[0]->max_n || batches[i]->x_dim != batches[0]->x_dim) {
                return SetError(batches[0]->ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "multi: batch %ld has another max_n / x_dim than batch 0", i);
            }
            total += batches[i]->num_gps;
        }
        if (!QueryOffsetsOk(q_offsets, total, num_q)) {
            return SetError(batches[0]->ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "multi: q_offsets must start at 0, be non-decreasing and end at num_q = %ld over the %ld GPs of all batches", num_q, total);
        }
        const long max_n = batches[0]->max_n, d = batches[0]->x_dim;
        std::vector<int> status(static_cast<size_t>(num_batches), ERL_GP_STATUS_OK);
        std::vector<std::vector<long>> offsets(static_cast<size_t>(num_batches));
        std::vector<long> first(static_cast<size_t>(num_batches) + 1, 0);
        for (long i = 0; i < num_batches; ++i) { first[i + 1] = first[i] + batches[i]->num_gps; }
        auto run = [&](const long i) {
            const long g0 = first[i], g1 = first[i + 1];
            const long t0 = q_offsets[g0], t1 = q_offsets[g1];
            std::vector<long> &off = offsets[i];  // the query lists of this range, re-based to 0
            off.resize(static_cast<size_t>(g1 - g0) + 1);
            for (long g = g0; g <= g1; ++g) { off[g - g0] = q_offsets[g] - t0; }
            status[i] = BatchTrainPredictHost<T>(batches[i], min_num_samples, n_train + g0, x + g0 * max_n * d, y + g0 * max_n, var + g0 * max_n, off.data(),
                                                 q_x != nullptr ? q_x + t0 * d : nullptr, t1 - t0, l != nullptr ? l + g0 * max_n * max_n : nullptr,
                                                 alpha != nullptr ? alpha + g0 * max_n : nullptr, info != nullptr ? info + g0 : nullptr, mean != nullptr ? mean + t0 : nullptr,
                                                 variance != nullptr ? variance + t0 : nullptr, valid != nullptr ? valid + t0 : nullptr);
        };
        std::vector<std::thread> workers;
        for (long i = 1; i < num_batches; ++i) { workers.emplace_back(run, i); }
        run(0);
        for (std::thread &w : workers) { w.join(); }
        for (long i = 0; i < num_batches; ++i) {
            if (status[i] != ERL_GP_STATUS_OK) { return status[i]; }
        }
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    int
    BatchGetGp(Batch<T> *b, long g, int *info, long *n, T *l, long ld_l, T *alpha) {
        if (b == nullptr || g < 0 || g >= b->num_gps) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = b->ctx;
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        int h_info = -1, h_n = 0;
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(&h_info, b->info.ptr + g, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(&h_n, b->n_train.ptr + g, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        if (info != nullptr) { *info = h_info; }
        if (n != nullptr) { *n = h_n; }
        if (h_info != 0 || h_n <= 0) { return ERL_GP_STATUS_OK; }
        if (l != nullptr) {
            if (ld_l < h_n) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "get_gp: ld_l=%ld < n=%d", ld_l, h_n); }
            ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(l, sizeof(T) * ld_l, b->l.ptr + static_cast<size_t>(g) * b->max_n * b->max_n, sizeof(T) * b->max_n, sizeof(T) * h_n, h_n,
                                                  cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (alpha != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(alpha, b->alpha.ptr + static_cast<size_t>(g) * b->max_n, sizeof(T) * h_n, cudaMemcpyDeviceToHost, ctx->stream)); }
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    // ---- covariance, host-pointer flavour: stage, launch, copy back ------------------------
    template<typename T>
    static int
    GramHost(erl_gp_context *c, bool train, int kernel, T scale, long x_dim, const T *x1, long ld_x1, long n1, const T *x2, long ld_x2, long n2, const T *var, T *k, long ld_k) {
        Context *ctx = Ctx(c);
        if (ctx == nullptr || x1 == nullptr || x2 == nullptr || k == nullptr || (train && var == nullptr)) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        if (n1 <= 0 || n2 <= 0 || ld_k < n1) { return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "gram: bad shape"); }
        DeviceBuffer<T> d_x1, d_x2, d_var, d_k;
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        ERL_GP_CUDA_OK(ctx, d_x1.Reserve(static_cast<size_t>(n1) * ld_x1));
        ERL_GP_CUDA_OK(ctx, d_k.Reserve(static_cast<size_t>(n1) * n2));
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(d_x1.ptr, x1, sizeof(T) * n1 * ld_x1, cudaMemcpyHostToDevice, ctx->stream));
        int rc;
        if (train) {
            ERL_GP_CUDA_OK(ctx, d_var.Reserve(n1));
            ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(d_var.ptr, var, sizeof(T) * n1, cudaMemcpyHostToDevice, ctx->stream));
            rc = LaunchKtrain<T>(ctx, kernel, scale, x_dim, d_x1.ptr, ld_x1, d_var.ptr, n1, d_k.ptr, n1);
        } else {
            ERL_GP_CUDA_OK(ctx, d_x2.Reserve(static_cast<size_t>(n2) * ld_x2));
            ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(d_x2.ptr, x2, sizeof(T) * n2 * ld_x2, cudaMemcpyHostToDevice, ctx->stream));
            rc = LaunchKtest<T>(ctx, kernel, scale, x_dim, d_x1.ptr, ld_x1, n1, d_x2.ptr, ld_x2, n2, d_k.ptr, n1);
        }
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        ERL_GP_CUDA_OK(ctx, cudaMemcpy2DAsync(k, sizeof(T) * ld_k, d_k.ptr, sizeof(T) * n1, sizeof(T) * n1, n2, cudaMemcpyDeviceToHost, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    // explicit instantiations used by the other translation units
#define ERL_GP_INSTANTIATE_BATCH(T)                                                                                                            \
    template int BatchCreate<T>(erl_gp_context *, long, long, long, int, T, Batch<T> **);                                                      \
    template int BatchUpload<T>(Batch<T> *, const int *, const T *, const T *, const T *);                                                     \
    template int BatchTrainDev<T>(Batch<T> *, long, int);                                                                                      \
    template int BatchPredictDev<T>(Batch<T> *, const long *, const T *, const int *, long, int, T, T *, T *, uint8_t *);                      \
    template int BatchTrainPredictDev<T>(Batch<T> *, long, int, const long *, const T *, long, T *, T *, uint8_t *);                           \
    template int BatchDownload<T>(Batch<T> *, T *, T *, int *);                                                                                \
    template int BatchGetGp<T>(Batch<T> *, long, int *, long *, T *, long, T *);
    ERL_GP_INSTANTIATE_BATCH(float)
    ERL_GP_INSTANTIATE_BATCH(double)
#undef ERL_GP_INSTANTIATE_BATCH

}  // namespace erl_gp

using namespace erl_gp;

struct erl_gp_batch_f32 : Batch<float> {};
struct erl_gp_batch_f64 : Batch<double> {};

extern "C" {

int
erl_gp_version(void) {
    return ERL_GP_B200_VERSION;
}

const char *
erl_gp_status_string(const int status) {
    switch (status) {
        case ERL_GP_STATUS_OK: return "ok";
        case ERL_GP_STATUS_INVALID_ARGUMENT: return "invalid argument";
        case ERL_GP_STATUS_CUDA_ERROR: return "CUDA error";
        case ERL_GP_STATUS_NOT_TRAINED: return "not trained";
        case ERL_GP_STATUS_UNSUPPORTED: return "unsupported shape or option";
        case ERL_GP_STATUS_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
        case ERL_GP_STATUS_ALLOC_FAILED: return "allocation failed";
        default: return "unknown status";
    }
}

int
erl_gp_device_count(int *count) {
    if (count == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void) cudaGetLastError();
        n = 0;
    }
    *count = n;
    return n > 0 ? ERL_GP_STATUS_OK : ERL_GP_STATUS_NO_DEVICE;
}

int
erl_gp_context_create(const int device, erl_gp_context **out) {
    if (out == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        (void) cudaGetLastError();
        return ERL_GP_STATUS_NO_DEVICE;
    }
    if (device < 0 || device >= n) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
    auto *ctx = new (std::nothrow) Context();
    if (ctx == nullptr) { return ERL_GP_STATUS_ALLOC_FAILED; }
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return ERL_GP_STATUS_CUDA_ERROR;
    }
    ctx->own_stream = true;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    *out = ctx;
    return ERL_GP_STATUS_OK;
}

int
erl_gp_context_destroy(erl_gp_context *c) {
    Context *ctx = Ctx(c);
    if (ctx == nullptr) { return ERL_GP_STATUS_OK; }
    cudaSetDevice(ctx->device);
    if (ctx->own_stream && ctx->stream != nullptr) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    if (ctx->side_stream != nullptr) {
        cudaStreamSynchronize(ctx->side_stream);
        cudaStreamDestroy(ctx->side_stream);
    }
    if (ctx->sync_ints != nullptr) { cudaFree(ctx->sync_ints); }
    if (ctx->side_stream2 != nullptr) {
        cudaStreamSynchronize(ctx->side_stream2);
        cudaStreamDestroy(ctx->side_stream2);
    }
    for (cudaEvent_t ev: {ctx->ev_panel, ctx->ev_diag, ctx->ev_la_start, ctx->ev_la_done}) {
        if (ev != nullptr) { cudaEventDestroy(ev); }
    }
    delete ctx;
    return ERL_GP_STATUS_OK;
}

int
erl_gp_context_set_stream(erl_gp_context *c, void *cuda_stream) {
    Context *ctx = Ctx(c);
    if (ctx == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
    ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    if (ctx->own_stream && ctx->stream != nullptr) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    ctx->own_stream = false;
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);  // NULL = the CUDA legacy default stream
    return ERL_GP_STATUS_OK;
}

int
erl_gp_context_synchronize(erl_gp_context *c) {
    Context *ctx = Ctx(c);
    if (ctx == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
    ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return ERL_GP_STATUS_OK;
}

const char *
erl_gp_context_last_error(const erl_gp_context *ctx) {
    return ctx == nullptr ? "null context" : ctx->last_error;
}

int
erl_gp_context_set_rowgp_tc(erl_gp_context *ctx, int on) {
    if (ctx == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
    ctx->rowgp_tc = on < 0 ? -1 : (on != 0 ? 1 : 0);
    return ERL_GP_STATUS_OK;
}

int
erl_gp_context_kernel_launches(const erl_gp_context *ctx, long *count) {
    if (ctx == nullptr || count == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
    *count = ctx->launches;
    return ERL_GP_STATUS_OK;
}

#define ERL_GP_DEFINE_TYPED(T, SFX)                                                                                                                                          \
    int erl_gp_compute_ktrain_##SFX(erl_gp_context *ctx, int kernel, T scale, long x_dim, const T *x, long ld_x, const T *var, long n, T *k, long ld_k) {                    \
        return GramHost<T>(ctx, true, kernel, scale, x_dim, x, ld_x, n, x, ld_x, n, var, k, ld_k);                                                                           \
    }                                                                                                                                                                        \
    int erl_gp_compute_ktest_##SFX(erl_gp_context *ctx, int kernel, T scale, long x_dim, const T *x1, long ld_x1, long n1, const T *x2, long ld_x2, long n2, T *k,           \
                                   long ld_k) {                                                                                                                              \
        return GramHost<T>(ctx, false, kernel, scale, x_dim, x1, ld_x1, n1, x2, ld_x2, n2, nullptr, k, ld_k);                                                                \
    }                                                                                                                                                                        \
    int erl_gp_compute_ktrain_dev_##SFX(erl_gp_context *ctx, int kernel, T scale, long x_dim, const T *x, long ld_x, const T *var, long n, T *k, long ld_k) {                \
        if (ctx == nullptr || x == nullptr || var == nullptr || k == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }                                                     \
        ERL_GP_CUDA_OK(Ctx(ctx), cudaSetDevice(ctx->device));                                                                                                                \
        return LaunchKtrain<T>(Ctx(ctx), kernel, scale, x_dim, x, ld_x, var, n, k, ld_k);                                                                                    \
    }                                                                                                                                                                        \
    int erl_gp_compute_ktest_dev_##SFX(erl_gp_context *ctx, int kernel, T scale, long x_dim, const T *x1, long ld_x1, long n1, const T *x2, long ld_x2, long n2, T *k,       \
                                       long ld_k) {                                                                                                                          \
        if (ctx == nullptr || x1 == nullptr || x2 == nullptr || k == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }                                                     \
        ERL_GP_CUDA_OK(Ctx(ctx), cudaSetDevice(ctx->device));                                                                                                                \
        return LaunchKtest<T>(Ctx(ctx), kernel, scale, x_dim, x1, ld_x1, n1, x2, ld_x2, n2, k, ld_k);                                                                        \
    }                                                                                                                                                                        \
    int erl_gp_batch_create_##SFX(erl_gp_context *ctx, long num_gps, long max_n, long x_dim, int kernel, T scale, erl_gp_batch_##SFX **batch) {                              \
        return BatchCreate<T>(ctx, num_gps, max_n, x_dim, kernel, scale, reinterpret_cast<Batch<T> **>(batch));                                                              \
    }                                                                                                                                                                        \
    int erl_gp_batch_destroy_##SFX(erl_gp_batch_##SFX *batch) {                                                                                                              \
        if (batch != nullptr) {                                                                                                                                              \
            cudaSetDevice(batch->ctx->device);                                                                                                                               \
            cudaStreamSynchronize(batch->ctx->stream);                                                                                                                       \
            delete static_cast<Batch<T> *>(batch);                                                                                                                           \
        }                                                                                                                                                                    \
        return ERL_GP_STATUS_OK;                                                                                                                                             \
    }                                                                                                                                                                        \
    int erl_gp_batch_device_buffers_##SFX(erl_gp_batch_##SFX *batch, int **n_train, T **x, T **y, T **var, T **l, T **alpha, int **info) {                                   \
        if (batch == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }                                                                                                     \
        if (n_train != nullptr) { *n_train = batch->n_train.ptr; }                                                                                                           \
        if (x != nullptr) { *x = batch->x.ptr; }                                                                                                                             \
        if (y != nullptr) { *y = batch->y.ptr; }                                                                                                                             \
        if (var != nullptr) { *var = batch->var.ptr; }                                                                                                                       \
        if (l != nullptr) { *l = batch->l.ptr; }                                                                                                                             \
        if (alpha != nullptr) { *alpha = batch->alpha.ptr; }                                                                                                                 \
        if (info != nullptr) { *info = batch->info.ptr; }                                                                                                                    \
        return ERL_GP_STATUS_OK;                                                                                                                                             \
    }                                                                                                                                                                        \
    int erl_gp_batch_upload_##SFX(erl_gp_batch_##SFX *batch, const int *n_train, const T *x, const T *y, const T *var) { return BatchUpload<T>(batch, n_train, x, y, var); } \
    int erl_gp_batch_train_dev_##SFX(erl_gp_batch_##SFX *batch, long min_num_samples, int write_l) { return BatchTrainDev<T>(batch, min_num_samples, write_l); }             \
    int erl_gp_batch_predict_dev_##SFX(erl_gp_batch_##SFX *batch, const long *q_offsets, const T *q_x, const int *q_out_index, long num_q, int mapping, T mapping_scale,     \
                                       T *mean, T *var, uint8_t *valid) {                                                                                                    \
        return BatchPredictDev<T>(batch, q_offsets, q_x, q_out_index, num_q, mapping, mapping_scale, mean, var, valid);                                                      \
    }                                                                                                                                                                        \
    int erl_gp_batch_train_predict_dev_##SFX(erl_gp_batch_##SFX *batch, long min_num_samples, int write_l, const long *q_offsets, const T *q_x, long num_q, T *mean, T *var, \
                                             uint8_t *valid) {                                                                                                               \
        return BatchTrainPredictDev<T>(batch, min_num_samples, write_l, q_offsets, q_x, num_q, mean, var, valid);                                                            \
    }                                                                                                                                                                        \
    int erl_gp_batch_train_predict_##SFX(erl_gp_batch_##SFX *batch, long min_num_samples, const int *n_train, const T *x, const T *y, const T *var, const long *q_offsets,   \
                                         const T *q_x, long num_q, T *l, T *alpha, int *info, T *mean, T *variance, uint8_t *valid) {                                        \
        return BatchTrainPredictHost<T>(batch, min_num_samples, n_train, x, y, var, q_offsets, q_x, num_q, l, alpha, info, mean, variance, valid);                           \
    }                                                                                                                                                                        \
    int erl_gp_batch_train_predict_multi_##SFX(erl_gp_batch_##SFX *const *batches, long num_batches, long min_num_samples, const int *n_train, const T *x, const T *y,        \
                                               const T *var, const long *q_offsets, const T *q_x, long num_q, T *l, T *alpha, int *info, T *mean, T *variance, uint8_t *valid) { \
        return BatchTrainPredictMulti<T>(reinterpret_cast<Batch<T> *const *>(batches), num_batches, min_num_samples, n_train, x, y, var, q_offsets, q_x, num_q, l, alpha, info,  \
                                         mean, variance, valid);                                                                                                             \
    }                                                                                                                                                                        \
    int erl_gp_batch_download_##SFX(erl_gp_batch_##SFX *batch, T *l, T *alpha, int *info) { return BatchDownload<T>(batch, l, alpha, info); }                                \
    int erl_gp_batch_get_gp_##SFX(erl_gp_batch_##SFX *batch, long gp_index, int *info, long *n, T *l, long ld_l, T *alpha) {                                                 \
        return BatchGetGp<T>(batch, gp_index, info, n, l, ld_l, alpha);                                                                                                      \
    }

ERL_GP_DEFINE_TYPED(float, f32)
ERL_GP_DEFINE_TYPED(double, f64)

}  // extern "C"
