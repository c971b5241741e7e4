// The slice of erl_covariance v0.2.0 that the GP classes touch: Covariance<Dtype>::Setting and the
// kernel_type strings (e.g. "erl::covariance::Matern32<float, 2>", config/spgp_occupancy_map_2d.yaml:2).
#pragma once

#include "erl_gp_b200.h"

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace erl::covariance {

    template<typename Dtype>
    class Covariance {
    public:
        struct Setting {  // config/spgp_occupancy_map_2d.yaml:4-8
            long x_dim = -1;
            Dtype scale = 1.0;
            Dtype scale_mix = 1.0;
            std::vector<Dtype> weights{};
        };
    };

    // "erl::covariance::OrnsteinUhlenbeck<double, 1>" / "...OrnsteinUhlenbeck1d" -> ERL_GP_KERNEL_OU, etc.
    inline int
    KernelFromTypeName(const std::string &type_name) {
        if (type_name.find("OrnsteinUhlenbeck") != std::string::npos) { return ERL_GP_KERNEL_OU; }
        if (type_name.find("Matern32") != std::string::npos) { return ERL_GP_KERNEL_MATERN32; }
        if (type_name.find("RadialBiasFunction") != std::string::npos) { return ERL_GP_KERNEL_RBF; }
        throw std::logic_error("failed to create kernel of type " + type_name + " (supported: OrnsteinUhlenbeck, Matern32, RadialBiasFunction)");
    }

}  // namespace erl::covariance
