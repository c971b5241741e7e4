// Instantiates the FP32 row-GP kernels (erl_gp_rowgp.cuh) for x_dim = 1, n <= 128 (own translation unit: build time).
#include "erl_gp_rowgp.cuh"

namespace erl_gp {
    namespace rowgp {
#ifndef ERL_GP_ROWGP_FAST_BUILD
        template int LaunchMode<1, 2>(Context *, const BatchParams<float> &, int, int);
#endif
#ifndef ERL_GP_ROWGP_FAST_BUILD
        template int LaunchMode<1, 4>(Context *, const BatchParams<float> &, int, int);
#endif
#ifndef ERL_GP_ROWGP_FAST_BUILD
        template int LaunchMode<1, 6>(Context *, const BatchParams<float> &, int, int);
#endif
        template int LaunchMode<1, 8>(Context *, const BatchParams<float> &, int, int);
    }  // namespace rowgp
}  // namespace erl_gp
