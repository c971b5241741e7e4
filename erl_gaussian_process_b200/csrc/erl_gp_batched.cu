// Dispatcher of the one-CTA-per-GP batched kernels (instances live in erl_gp_batched_<dtype>_x<dim>.cu).
#include "erl_gp_internal.cuh"

namespace erl_gp {

    template<typename T, int XDIM>
    int
    LaunchBatchXdim(Context *ctx, const BatchParams<T> &params, int mode, int tiles_per_gp);

    extern template int LaunchBatchXdim<float, 1>(Context *, const BatchParams<float> &, int, int);
    extern template int LaunchBatchXdim<float, 2>(Context *, const BatchParams<float> &, int, int);
    extern template int LaunchBatchXdim<float, 3>(Context *, const BatchParams<float> &, int, int);
    extern template int LaunchBatchXdim<double, 1>(Context *, const BatchParams<double> &, int, int);
    extern template int LaunchBatchXdim<double, 2>(Context *, const BatchParams<double> &, int, int);
    extern template int LaunchBatchXdim<double, 3>(Context *, const BatchParams<double> &, int, int);

    template<typename T>
    long
    BatchMaxN() {
        return sizeof(T) == 4 ? 256 : 192;
    }

    template long BatchMaxN<float>();
    template long BatchMaxN<double>();

    template<typename T>
    int
    LaunchBatch(Context *ctx, const BatchParams<T> &params, const int x_dim, const int mode, const int tiles_per_gp) {
        if (params.num_gps <= 0) { return ERL_GP_STATUS_OK; }
        if (x_dim < 1 || x_dim > 3) { return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "batch: x_dim=%d (supported: 1, 2, 3)", x_dim); }
        // beyond the shared-memory kernels: L stays in HBM / L2 (erl_gp_largegp.cu); src/range_sensor_gp_3d.cpp:213 puts no cap on n
        if (params.max_n > BatchMaxN<T>()) { return LaunchLargeGp<T>(ctx, params, x_dim, mode, tiles_per_gp); }
        switch (x_dim) {
            case 1: return LaunchBatchXdim<T, 1>(ctx, params, mode, tiles_per_gp);
            case 2: return LaunchBatchXdim<T, 2>(ctx, params, mode, tiles_per_gp);
            case 3: return LaunchBatchXdim<T, 3>(ctx, params, mode, tiles_per_gp);
            default: return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "batch: x_dim=%d (supported: 1, 2, 3)", x_dim);
        }
    }

    template int LaunchBatch<float>(Context *, const BatchParams<float> &, int, int, int);
    template int LaunchBatch<double>(Context *, const BatchParams<double> &, int, int, int);

}  // namespace erl_gp
