// Instantiates the FP64 row-GP kernels (erl_gp_rowgp64.cuh) for x_dim = 1 (own translation unit: build time).
#include "erl_gp_rowgp64.cuh"

namespace erl_gp {
    namespace rowgp64 {
        template int Launch<1>(Context *, const BatchParams<double> &, int, int);
    }  // namespace rowgp64
}  // namespace erl_gp
