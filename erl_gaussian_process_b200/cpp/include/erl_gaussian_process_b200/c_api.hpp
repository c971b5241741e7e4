// Typed access to the C ABI (include/erl_gp_b200.h) for the templated drop-in classes.
#pragma once

#include "erl_gp_b200.h"

#include <memory>
#include <stdexcept>
#include <string>

namespace erl::gaussian_process::b200 {

    // ERL_ASSERTM equivalent: hard failure on misuse (the reference aborts; we throw so tests can observe it)
    inline void
    AssertM(const bool ok, const std::string &msg) {
        if (!ok) { throw std::logic_error(msg); }
    }

    // One process-wide context per device, created on first use (the reference has no such object: its
    // state lives in the GP instances, which keep a shared_ptr to this).
    class DeviceContext {
        erl_gp_context *m_ctx_ = nullptr;

    public:
        explicit DeviceContext(const int device) {
            const int rc = erl_gp_context_create(device, &m_ctx_);
            if (rc != ERL_GP_STATUS_OK) { throw std::runtime_error(std::string("erl_gp_context_create: ") + erl_gp_status_string(rc)); }
        }

        DeviceContext(const DeviceContext &) = delete;
        DeviceContext &
        operator=(const DeviceContext &) = delete;

        ~DeviceContext() { erl_gp_context_destroy(m_ctx_); }

        [[nodiscard]] erl_gp_context *
        Get() const {
            return m_ctx_;
        }

        void
        Check(const int rc, const char *where) const {
            if (rc != ERL_GP_STATUS_OK) { throw std::runtime_error(std::string(where) + ": " + erl_gp_status_string(rc) + " — " + erl_gp_context_last_error(m_ctx_)); }
        }

        static std::shared_ptr<DeviceContext>
        Default(const int device = 0) {
            static std::shared_ptr<DeviceContext> ctx = std::make_shared<DeviceContext>(device);
            return ctx;
        }
    };

    template<typename Dtype>
    struct Api;

#define ERL_GP_API_STRUCT(T, SFX)                                                        \
    template<>                                                                           \
    struct Api<T> {                                                                      \
        using Vanilla = erl_gp_vanilla_##SFX;                                            \
        using Batch = erl_gp_batch_##SFX;                                                \
        using Lidar2d = erl_gp_lidar2d_##SFX;                                            \
        using Range3d = erl_gp_range3d_##SFX;                                            \
        using Noisy = erl_gp_noisy_##SFX;                                                \
        static constexpr auto noisy_create = erl_gp_noisy_create_##SFX;                  \
        static constexpr auto noisy_destroy = erl_gp_noisy_destroy_##SFX;                \
        static constexpr auto noisy_train = erl_gp_noisy_train_##SFX;                    \
        static constexpr auto noisy_get = erl_gp_noisy_get_##SFX;                        \
        static constexpr auto noisy_test = erl_gp_noisy_test_##SFX;                      \
        static constexpr auto compute_ktrain = erl_gp_compute_ktrain_##SFX;              \
        static constexpr auto compute_ktest = erl_gp_compute_ktest_##SFX;                \
        static constexpr auto vanilla_create = erl_gp_vanilla_create_##SFX;              \
        static constexpr auto vanilla_destroy = erl_gp_vanilla_destroy_##SFX;            \
        static constexpr auto vanilla_train = erl_gp_vanilla_train_##SFX;                \
        static constexpr auto vanilla_get = erl_gp_vanilla_get_##SFX;                    \
        static constexpr auto vanilla_test = erl_gp_vanilla_test_##SFX;                  \
        static constexpr auto batch_create = erl_gp_batch_create_##SFX;                  \
        static constexpr auto batch_destroy = erl_gp_batch_destroy_##SFX;                \
        static constexpr auto batch_train_predict = erl_gp_batch_train_predict_##SFX;    \
        static constexpr auto lidar2d_create = erl_gp_lidar2d_create_##SFX;              \
        static constexpr auto lidar2d_destroy = erl_gp_lidar2d_destroy_##SFX;            \
        static constexpr auto lidar2d_num_partitions = erl_gp_lidar2d_num_partitions_##SFX; \
        static constexpr auto lidar2d_partitions = erl_gp_lidar2d_partitions_##SFX;      \
        static constexpr auto lidar2d_train = erl_gp_lidar2d_train_##SFX;                \
        static constexpr auto lidar2d_test = erl_gp_lidar2d_test_##SFX;                  \
        static constexpr auto lidar2d_get_gp = erl_gp_lidar2d_get_gp_##SFX;              \
        static constexpr auto lidar2d_compute_occ = erl_gp_lidar2d_compute_occ_##SFX;    \
        static constexpr auto range3d_create = erl_gp_range3d_create_##SFX;              \
        static constexpr auto range3d_destroy = erl_gp_range3d_destroy_##SFX;            \
        static constexpr auto range3d_grid = erl_gp_range3d_grid_##SFX;                  \
        static constexpr auto range3d_partitions = erl_gp_range3d_partitions_##SFX;      \
        static constexpr auto range3d_train = erl_gp_range3d_train_##SFX;                \
        static constexpr auto range3d_test = erl_gp_range3d_test_##SFX;                  \
        static constexpr auto range3d_get_gp = erl_gp_range3d_get_gp_##SFX;              \
        static constexpr auto range3d_compute_occ = erl_gp_range3d_compute_occ_##SFX;    \
    };
    ERL_GP_API_STRUCT(float, f32)
    ERL_GP_API_STRUCT(double, f64)
#undef ERL_GP_API_STRUCT

}  // namespace erl::gaussian_process::b200
