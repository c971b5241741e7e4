// Internal declarations shared by the .cu translation units of liberl_gp_b200.so.
#pragma once

#include "erl_gp_common.cuh"

#include <cstdarg>
#include <cstring>
#include <new>
#include <vector>

struct erl_gp_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    int max_smem_optin = 0;
    long launches = 0;
    int rowgp_tc = -1;  // fused FP32 train + predict, n <= 128, on the tcgen05 / TMEM kernel: 1 / 0, -1 = ERL_GP_ROWGP_TC or the default
    char last_error[512] = {0};
    // look-ahead of the blocked Cholesky (erl_gp_dense.cu): side streams + events, created on first use
    cudaStream_t side_stream = nullptr, side_stream2 = nullptr;
    cudaEvent_t ev_panel = nullptr, ev_diag = nullptr, ev_la_start = nullptr, ev_la_done = nullptr;
    // ticket counter + per-block flags of the wavefront TRSV (erl_gp_dense.cu), grow-only
    int *sync_ints = nullptr;
    int sync_ints_capacity = 0;
};

namespace erl_gp {

    struct Context : erl_gp_context {};

    inline Context *
    Ctx(erl_gp_context *c) {
        return static_cast<Context *>(c);
    }

    // Grow-only device buffer (mirrors the reference's grow-only Eigen buffers,
    // src/vanilla_gp.cpp:152-161, 805-812): cudaMalloc is never on the steady-state path.
    template<typename T>
    struct DeviceBuffer {
        T *ptr = nullptr;
        size_t capacity = 0;  // elements

        DeviceBuffer() = default;
        DeviceBuffer(const DeviceBuffer &) = delete;
        DeviceBuffer &
        operator=(const DeviceBuffer &) = delete;

        ~DeviceBuffer() { Free(); }

        void
        Free() {
            if (ptr != nullptr) { cudaFree(ptr); }
            ptr = nullptr;
            capacity = 0;
        }

        cudaError_t
        Reserve(size_t count) {
            if (count <= capacity) { return cudaSuccess; }
            Free();
            const cudaError_t err = cudaMalloc(&ptr, count * sizeof(T));
            if (err == cudaSuccess) { capacity = count; }
            return err;
        }
    };

    // Pinned host staging buffer, grow-only.
    template<typename T>
    struct PinnedBuffer {
        T *ptr = nullptr;
        size_t capacity = 0;

        PinnedBuffer() = default;
        PinnedBuffer(const PinnedBuffer &) = delete;
        PinnedBuffer &
        operator=(const PinnedBuffer &) = delete;

        ~PinnedBuffer() {
            if (ptr != nullptr) { cudaFreeHost(ptr); }
        }

        cudaError_t
        Reserve(size_t count) {
            if (count <= capacity) { return cudaSuccess; }
            if (ptr != nullptr) { cudaFreeHost(ptr); }
            ptr = nullptr;
            capacity = 0;
            const cudaError_t err = cudaMallocHost(&ptr, count * sizeof(T));
            if (err == cudaSuccess) { capacity = count; }
            return err;
        }
    };

    // ---- launchers implemented in the kernel translation units (all async on ctx->stream) ----
    template<typename T>
    int
    LaunchKtrain(Context *ctx, int kernel, T scale, long x_dim, const T *x, long ld_x, const T *var, long n, T *k, long ld_k);

    template<typename T>
    int
    LaunchKtest(Context *ctx, int kernel, T scale, long x_dim, const T *x1, long ld_x1, long n1, const T *x2, long ld_x2, long n2, T *k, long ld_k);

    template<typename T>
    struct BatchParams {
        Covariance<T> cov;
        int num_gps;
        int max_n;
        int min_train;  // train iff n > min_train
        int write_l;
        const int *n_train;
        const T *x;
        const T *y;
        const T *var;
        T *l;
        T *alpha;
        int *info;
        // prediction
        const long *q_offsets;  // [num_gps + 1]
        const T *q_x;           // [T][x_dim]
        const int *q_out_index; // optional scatter index
        T *mean;
        T *variance;
        uint8_t *valid;
        int mapping;  // ERL_GP_MAPPING_NONE = no un-map
        T mapping_scale;
        // row-GP kernel, fused train + predict: CTA slot k of an SM (first wave only) starts k * stagger_cycles late
        int stagger_cycles = 0;
        int sm_count = 148;
        // row-GP kernel: L goes to HBM with one bulk asynchronous copy (TMA, cp.async.bulk shared -> global) per column instead of
        // LDS + STG by the warps; needs the L slices zero above the diagonal blocks (erl_gp_batch_create zeroes them once)
        int tma_writeback = 0;
    };

    enum BatchMode : int { kBatchTrain = 1, kBatchPredict = 2, kBatchTrainPredict = 3 };

    // tiles_per_gp: how many CTAs share one GP's query list (predict-only mode), >= 1.
    template<typename T>
    int
    LaunchBatch(Context *ctx, const BatchParams<T> &params, int x_dim, int mode, int tiles_per_gp);

    // maximum capacity (max_n) the one-CTA-per-GP shared-memory kernels support for this Dtype
    template<typename T>
    long
    BatchMaxN();

    // larger partition GPs (L resident in HBM / L2, erl_gp_largegp.cu): capacity limit and launcher
    long
    LargeGpMaxN();

    template<typename T>
    int
    LaunchLargeGp(Context *ctx, const BatchParams<T> &params, int x_dim, int mode, int tiles_per_gp);

    template<typename T>
    struct Batch {
        Context *ctx = nullptr;
        long num_gps = 0, max_n = 0, x_dim = 0;
        int kernel = 0;
        T scale = T(1);
        DeviceBuffer<int> n_train, info;
        DeviceBuffer<T> x, y, var, l, alpha;
        // query-side staging for the host-pointer entry points
        DeviceBuffer<long> q_offsets;
        DeviceBuffer<T> q_x, mean, variance;
        DeviceBuffer<uint8_t> valid;
        // host-buffer pipeline (BatchTrainPredictHost): copy streams and per-chunk events, created on first use
        cudaStream_t copy_in = nullptr, copy_out = nullptr, compute2 = nullptr;  // compute2: odd chunks of the host-buffer pipeline
        std::vector<cudaEvent_t> ev_in, ev_kernel;
        PinnedBuffer<int> info_host;

        ~Batch() {
            if (copy_in != nullptr) { cudaStreamDestroy(copy_in); }
            if (copy_out != nullptr) { cudaStreamDestroy(copy_out); }
            if (compute2 != nullptr) { cudaStreamDestroy(compute2); }
            for (cudaEvent_t e: ev_in) { cudaEventDestroy(e); }
            for (cudaEvent_t e: ev_kernel) { cudaEventDestroy(e); }
        }

        BatchParams<T>
        Params(const long min_num_samples, const int write_l) const {
            BatchParams<T> p{};
            p.cov = Covariance<T>::Make(kernel, scale);
            p.num_gps = static_cast<int>(num_gps);
            p.max_n = static_cast<int>(max_n);
            p.min_train = static_cast<int>(min_num_samples < 0 ? 0 : min_num_samples);
            p.write_l = write_l;
            p.n_train = n_train.ptr;
            p.x = x.ptr;
            p.y = y.ptr;
            p.var = var.ptr;
            p.l = l.ptr;
            p.alpha = alpha.ptr;
            p.info = info.ptr;
            p.mapping = ERL_GP_MAPPING_NONE;
            p.mapping_scale = T(1);
            return p;
        }
    };

    template<typename T>
    int
    BatchCreate(erl_gp_context *c, long num_gps, long max_n, long x_dim, int kernel, T scale, Batch<T> **out);
    template<typename T>
    int
    BatchUpload(Batch<T> *b, const int *n_train, const T *x, const T *y, const T *var);
    template<typename T>
    int
    BatchTrainDev(Batch<T> *b, long min_num_samples, int write_l);
    template<typename T>
    int
    BatchPredictDev(Batch<T> *b, const long *q_offsets, const T *q_x, const int *q_out_index, long num_q, int mapping, T mapping_scale, T *mean, T *var, uint8_t *valid);
    template<typename T>
    int
    BatchTrainPredictDev(Batch<T> *b, long min_num_samples, int write_l, const long *q_offsets, const T *q_x, long num_q, T *mean, T *var, uint8_t *valid);
    template<typename T>
    int
    BatchDownload(Batch<T> *b, T *l, T *alpha, int *info);
    template<typename T>
    int
    BatchGetGp(Batch<T> *b, long g, int *info, long *n, T *l, long ld_l, T *alpha);

}  // namespace erl_gp
