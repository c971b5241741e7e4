// Instantiates the one-CTA-per-GP kernels for Dtype = float, x_dim = 3 (own translation unit: build time).
#include "erl_gp_batched.cuh"

namespace erl_gp {
    template int LaunchBatchXdim<float, 3>(Context *, const BatchParams<float> &, int, int);
}  // namespace erl_gp
