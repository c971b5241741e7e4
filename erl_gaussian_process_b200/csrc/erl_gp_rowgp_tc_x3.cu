// Instantiates the tcgen05 fused train + predict kernel (erl_gp_rowgp_tc.cuh) for x_dim = 3 (own translation unit: build time).
#define ERL_GP_ROWGP_EXTERN_INSTANCES
#include "erl_gp_rowgp_tc.cuh"

namespace erl_gp {
    namespace rowgp_tc {
        template int Launch<3>(Context *, const BatchParams<float> &);
    }  // namespace rowgp_tc
}  // namespace erl_gp
