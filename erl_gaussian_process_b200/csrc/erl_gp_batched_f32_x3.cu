// Instantiates the one-CTA-per-GP kernels for Dtype = float, x_dim = 3 (own translation unit: build time).
#define ERL_GP_ROWGP_EXTERN_INSTANCES  // the row-GP kernels live in erl_gp_rowgp_x3_<a|b|c>.cu
#include "erl_gp_batched.cuh"

namespace erl_gp {
    template int LaunchBatchXdim<float, 3>(Context *, const BatchParams<float> &, int, int);
}  // namespace erl_gp
