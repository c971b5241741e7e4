// Instantiates the one-CTA-per-GP kernels for Dtype = float, x_dim = 2 (own translation unit: build time).
#define ERL_GP_ROWGP_EXTERN_INSTANCES  // the row-GP kernels live in erl_gp_rowgp_x2_<a|b|c>.cu
#include "erl_gp_batched.cuh"

namespace erl_gp {
    template int LaunchBatchXdim<float, 2>(Context *, const BatchParams<float> &, int, int);
}  // namespace erl_gp
