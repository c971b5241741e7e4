// Instantiates the FP32 row-GP kernels (erl_gp_rowgp.cuh) for x_dim = 3, n <= 192 (own translation unit: build time).
#include "erl_gp_rowgp.cuh"

namespace erl_gp {
    namespace rowgp {
#ifndef ERL_GP_ROWGP_FAST_BUILD
        template int LaunchMode<3, 12>(Context *, const BatchParams<float> &, int, int);
#endif
    }  // namespace rowgp
}  // namespace erl_gp
