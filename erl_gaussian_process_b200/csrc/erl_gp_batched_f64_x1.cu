// Instantiates the one-CTA-per-GP kernels for Dtype = double, x_dim = 1 (own translation unit: build time).
#include "erl_gp_batched.cuh"

namespace erl_gp {
    template int LaunchBatchXdim<double, 1>(Context *, const BatchParams<double> &, int, int);
}  // namespace erl_gp
