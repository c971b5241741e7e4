// Instantiates the FP32 row-GP kernels (erl_gp_rowgp.cuh) for x_dim = 1, n <= 256 (own translation unit: build time).
#include "erl_gp_rowgp.cuh"

namespace erl_gp {
    namespace rowgp {
#ifndef ERL_GP_ROWGP_FAST_BUILD
        template int LaunchMode<1, 16>(Context *, const BatchParams<float> &, int, int);
#endif
    }  // namespace rowgp
}  // namespace erl_gp
