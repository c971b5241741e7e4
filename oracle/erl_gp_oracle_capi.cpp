// ORACLE — TEST INFRASTRUCTURE ONLY (see erl_gp_oracle.hpp).
// Flat C exports of the CPU restatement so pytest / bench.py can drive it through ctypes.
// Every function exists in an _f32 and an _f64 flavour.
#include "erl_gp_oracle.hpp"

#include <chrono>
#include <omp.h>

using namespace erl_gp_oracle;

namespace {
    template<typename T>
    int
    BatchedTrainPredict(
        const int kernel,
        const T scale,
        const long x_dim,
        const long num_gps,
        const long max_n,
        const int *n_train,
        const T *x,
        const T *y,
        const T *var,
        const long *q_offsets,
        const T *q_x,
        T *mat_l,
        T *alpha,
        int *info,
        T *mean,
        T *variance) {
        // The batched small-GP stream (BASELINE.json config 4): every GP is an independent
        // VanillaGaussianProcess Reset/Train/Test — the loop the reference runs under
        // `#pragma omp parallel for` at src/lidar_gp_2d.cpp:366-392.
#pragma omp parallel
        {
            VanillaGp<T> gp;
            gp.kernel_type = kernel;
            gp.scale = scale;
            gp.max_num_samples_setting = max_n;
#pragma omp for schedule(dynamic, 4)
            for (long g = 0; g < num_gps; ++g) {
                const long n = n_train[g];
                gp.Reset(max_n, x_dim, 1);
                if (n <= 0) {
                    if (info != nullptr) { info[g] = -1; }
                    continue;
                }
                std::memcpy(gp.x.data(), x + g * max_n * x_dim, sizeof(T) * n * x_dim);
                std::memcpy(gp.y.data(), y + g * max_n, sizeof(T) * n);
                std::memcpy(gp.var.data(), var + g * max_n, sizeof(T) * n);
                gp.num_samples = n;
                (void) gp.Train();
                if (info != nullptr) { info[g] = gp.llt_info; }
                if (mat_l != nullptr) {
                    T *dst = mat_l + g * max_n * max_n;
                    for (long c = 0; c < n; ++c) { std::memcpy(dst + c * max_n, gp.mat_l.data() + c * gp.ld, sizeof(T) * n); }
                }
                if (alpha != nullptr) { std::memcpy(alpha + g * max_n, gp.mat_alpha.data(), sizeof(T) * n); }
                if (q_offsets != nullptr) {
                    const long q0 = q_offsets[g], q1 = q_offsets[g + 1];
                    if (q1 > q0) {
                        gp.Test(q_x + q0 * x_dim, x_dim, q1 - q0, mean != nullptr ? mean + q0 : nullptr,
                                variance != nullptr ? variance + q0 : nullptr, /*parallel=*/false);
                    }
                }
            }
        }
        return 0;
    }
}  // namespace

#define ORACLE_EXPORTS(T, SFX)                                                                                        \
    extern "C" int oracle_gram_train_##SFX(int kernel, T scale, long x_dim, const T *x, long ld_x, const T *var,      \
                                           long n, T *k, long ld_k) {                                                 \
        ComputeKtrain<T>(kernel, scale, x_dim, x, ld_x, var, n, k, ld_k);                                             \
        return 0;                                                                                                     \
    }                                                                                                                 \
    extern "C" int oracle_gram_test_##SFX(int kernel, T scale, long x_dim, const T *x1, long ld_x1, long n1,          \
                                          const T *x2, long ld_x2, long n2, T *k, long ld_k) {                        \
        ComputeKtest<T>(kernel, scale, x_dim, x1, ld_x1, n1, x2, ld_x2, n2, k, ld_k);                                 \
        return 0;                                                                                                     \
    }                                                                                                                 \
    extern "C" int oracle_llt_##SFX(const T *k, long ld_k, long n, T *l, long ld_l) {                                 \
        return Llt<T>(k, ld_k, n, l, ld_l);                                                                           \
    }                                                                                                                 \
    extern "C" T oracle_mapping_map_##SFX(int type, T scale, T x) { return MappingMap<T>(type, scale, x); }           \
    extern "C" T oracle_mapping_inv_##SFX(int type, T scale, T y) { return MappingInv<T>(type, scale, y); }           \
    /* ---- VanillaGaussianProcess ---- */                                                                            \
    extern "C" void *oracle_vanilla_create_##SFX(int kernel, T scale, long max_num_samples) {                         \
        auto *gp = new VanillaGp<T>();                                                                                \
        gp->kernel_type = kernel;                                                                                     \
        gp->scale = scale;                                                                                            \
        gp->max_num_samples_setting = max_num_samples;                                                                \
        return gp;                                                                                                    \
    }                                                                                                                 \
    extern "C" void oracle_vanilla_destroy_##SFX(void *h) { delete static_cast<VanillaGp<T> *>(h); }                  \
    extern "C" int oracle_vanilla_train_##SFX(void *h, long n, long x_dim, long y_dim, const T *x, long ld_x,         \
                                              const T *y, long ld_y, const T *var) {                                  \
        auto *gp = static_cast<VanillaGp<T> *>(h);                                                                    \
        if (!gp->Reset(n, x_dim, y_dim)) { return -1; }                                                               \
        for (long i = 0; i < n; ++i) {                                                                                \
            for (long d = 0; d < x_dim; ++d) { gp->x[d + i * gp->x_rows] = x[d + i * ld_x]; }                         \
            for (long c = 0; c < y_dim; ++c) { gp->y[i + c * gp->y_rows] = y[i + c * ld_y]; }                         \
            gp->var[i] = var[i];                                                                                      \
        }                                                                                                             \
        gp->num_samples = n;                                                                                          \
        return gp->Train() ? gp->llt_info : -2;                                                                       \
    }                                                                                                                 \
    extern "C" int oracle_vanilla_get_##SFX(void *h, T *k, long ld_k, T *l, long ld_l, T *alpha, long ld_a) {         \
        auto *gp = static_cast<VanillaGp<T> *>(h);                                                                    \
        const long n = gp->k_cols;                                                                                    \
        for (long c = 0; c < n; ++c) {                                                                                \
            for (long r = 0; r < n; ++r) {                                                                            \
                if (k != nullptr) { k[r + c * ld_k] = gp->mat_k[r + c * gp->ld]; }                                    \
                if (l != nullptr) { l[r + c * ld_l] = gp->mat_l[r + c * gp->ld]; }                                    \
            }                                                                                                         \
        }                                                                                                             \
        if (alpha != nullptr) {                                                                                       \
            for (long c = 0; c < gp->y_dim; ++c) {                                                                    \
                for (long r = 0; r < n; ++r) { alpha[r + c * ld_a] = gp->mat_alpha[r + c * gp->alpha_rows]; }         \
            }                                                                                                         \
        }                                                                                                             \
        return static_cast<int>(n);                                                                                   \
    }                                                                                                                 \
    extern "C" int oracle_vanilla_test_##SFX(void *h, const T *x_test, long ld_xt, long num_test, T *mean,            \
                                             T *variance, int parallel) {                                             \
        auto *gp = static_cast<VanillaGp<T> *>(h);                                                                    \
        return gp->Test(x_test, ld_xt, num_test, mean, variance, parallel != 0) ? 0 : -1;                             \
    }                                                                                                                 \
    /* ---- partitions ---- */                                                                                        \
    extern "C" long oracle_make_partitions_##SFX(const T *coords, long stride, long n, long group_size,               \
                                                 long overlap_size, long margin, int symmetric, long capacity,        \
                                                 long *index_left, long *index_right, T *coord_left,                  \
                                                 T *coord_right) {                                                    \
        const auto parts = MakePartitions<T>(coords, stride, n, group_size, overlap_size, margin, symmetric != 0);    \
        const long np = static_cast<long>(parts.size());                                                              \
        for (long i = 0; i < std::min(np, capacity); ++i) {                                                           \
            index_left[i] = parts[i].index_left;                                                                      \
            index_right[i] = parts[i].index_right;                                                                    \
            coord_left[i] = parts[i].coord_left;                                                                      \
            coord_right[i] = parts[i].coord_right;                                                                    \
        }                                                                                                             \
        return np;                                                                                                    \
    }                                                                                                                 \
    /* ---- LidarGaussianProcess2D ---- */                                                                            \
    extern "C" void *oracle_lidar_create_##SFX(int kernel, T scale, long group_size, long overlap_size, long margin,  \
                                               int symmetric, T sensor_range_var, T discontinuity_var,                \
                                               int discontinuity_detection, int mapping_type, T mapping_scale,        \
                                               T max_valid_range_var, T occ_test_temperature, const T *angles,        \
                                               long n) {                                                              \
        auto *gp = new LidarGp2D<T>();                                                                                \
        gp->kernel_type = kernel;                                                                                     \
        gp->kernel_scale = scale;                                                                                     \
        gp->group_size = group_size;                                                                                  \
        gp->overlap_size = overlap_size;                                                                              \
        gp->margin = margin;                                                                                          \
        gp->symmetric_partitions = symmetric != 0;                                                                    \
        gp->sensor_range_var = sensor_range_var;                                                                      \
        gp->discontinuity_var = discontinuity_var;                                                                    \
        gp->discontinuity_detection = discontinuity_detection != 0;                                                   \
        gp->mapping_type = mapping_type;                                                                              \
        gp->mapping_scale = mapping_scale;                                                                            \
        gp->max_valid_range_var = max_valid_range_var;                                                                \
        gp->occ_test_temperature = occ_test_temperature;                                                              \
        gp->Init(angles, n);                                                                                          \
        return gp;                                                                                                    \
    }                                                                                                                 \
    extern "C" void oracle_lidar_destroy_##SFX(void *h) { delete static_cast<LidarGp2D<T> *>(h); }                    \
    /* Setting::partition_on_hit_rays; call right after create (the table is then built by every Train) */           \
    extern "C" void oracle_lidar_set_partition_on_hit_rays_##SFX(void *h, int flag) {                                 \
        auto *lg = static_cast<LidarGp2D<T> *>(h);                                                                    \
        lg->partition_on_hit_rays = flag != 0;                                                                        \
        if (flag != 0) {                                                                                              \
            lg->partitions.clear();                                                                                   \
            lg->gps.clear();                                                                                          \
        }                                                                                                             \
    }                                                                                                                 \
    extern "C" long oracle_lidar_partitions_##SFX(void *h, long *il, long *ir, T *cl, T *cr) {                        \
        auto *lg = static_cast<LidarGp2D<T> *>(h);                                                                    \
        for (std::size_t i = 0; i < lg->partitions.size(); ++i) {                                                     \
            il[i] = lg->partitions[i].index_left, ir[i] = lg->partitions[i].index_right;                              \
            cl[i] = lg->partitions[i].coord_left, cr[i] = lg->partitions[i].coord_right;                              \
        }                                                                                                             \
        return static_cast<long>(lg->partitions.size());                                                              \
    }                                                                                                                 \
    extern "C" long oracle_lidar_num_partitions_##SFX(void *h) {                                                      \
        return static_cast<long>(static_cast<LidarGp2D<T> *>(h)->partitions.size());                                  \
    }                                                                                                                 \
    extern "C" int oracle_lidar_train_##SFX(void *h, const T *rotation, const T *ranges, const uint8_t *mask_hit,     \
                                            const uint8_t *mask_con, int frame_valid) {                               \
        return static_cast<LidarGp2D<T> *>(h)->Train(rotation, ranges, mask_hit, mask_con, frame_valid != 0) ? 0      \
                                                                                                              : -1;   \
    }                                                                                                                 \
    extern "C" int oracle_lidar_get_gp_##SFX(void *h, long p, int *trained, long *n, T *l, long ld_l, T *alpha) {     \
        auto *lg = static_cast<LidarGp2D<T> *>(h);                                                                    \
        const auto &gp = lg->gps[p];                                                                                  \
        *trained = gp.trained ? 1 : 0;                                                                                \
        *n = gp.trained ? gp.k_cols : gp.num_samples;                                                                 \
        if (!gp.trained) { return 0; }                                                                                \
        for (long c = 0; c < gp.k_cols; ++c) {                                                                        \
            for (long r = 0; r < gp.k_cols; ++r) {                                                                    \
                if (l != nullptr) { l[r + c * ld_l] = gp.mat_l[r + c * gp.ld]; }                                      \
            }                                                                                                         \
            if (alpha != nullptr) { alpha[c] = gp.mat_alpha[c]; }                                                     \
        }                                                                                                             \
        return 0;                                                                                                     \
    }                                                                                                                 \
    extern "C" int oracle_lidar_test_##SFX(void *h, const T *angles, long num_test, int angles_are_local,             \
                                           int un_map, T *mean, T *variance, uint8_t *valid) {                        \
        return static_cast<LidarGp2D<T> *>(h)->Test(angles, num_test, angles_are_local != 0, un_map != 0, mean,       \
                                                    variance, valid)                                                  \
                   ? 0                                                                                                \
                   : -1;                                                                                              \
    }                                                                                                                 \
    extern "C" int oracle_lidar_compute_occ_##SFX(void *h, T px, T py, T *dist, T *range_pred, T *occ) {              \
        return static_cast<LidarGp2D<T> *>(h)->ComputeOcc(px, py, *dist, *range_pred, *occ) ? 1 : 0;                  \
    }                                                                                                                 \
    /* ---- RangeSensorGaussianProcess3D ---- */                                                                      \
    extern "C" void *oracle_range3d_create_##SFX(int kernel, T scale, long row_group_size, long row_overlap_size,     \
                                                 long row_margin, long col_group_size, long col_overlap_size,         \
                                                 long col_margin, long min_num_samples_per_group,                     \
                                                 T sensor_range_var, int mapping_type, T mapping_scale,               \
                                                 const T *frame_coords, long rows, long cols) {                       \
        auto *gp = new RangeSensorGp3D<T>();                                                                          \
        gp->kernel_type = kernel;                                                                                     \
        gp->kernel_scale = scale;                                                                                     \
        gp->row_group_size = row_group_size;                                                                          \
        gp->row_overlap_size = row_overlap_size;                                                                      \
        gp->row_margin = row_margin;                                                                                  \
        gp->col_group_size = col_group_size;                                                                          \
        gp->col_overlap_size = col_overlap_size;                                                                      \
        gp->col_margin = col_margin;                                                                                  \
        gp->min_num_samples_per_group = min_num_samples_per_group;                                                    \
        gp->sensor_range_var = sensor_range_var;                                                                      \
        gp->mapping_type = mapping_type;                                                                              \
        gp->mapping_scale = mapping_scale;                                                                            \
        if (!gp->Init(frame_coords, rows, cols)) {                                                                    \
            delete gp;                                                                                                \
            return nullptr;                                                                                           \
        }                                                                                                             \
        return gp;                                                                                                    \
    }                                                                                                                 \
    extern "C" void oracle_range3d_destroy_##SFX(void *h) { delete static_cast<RangeSensorGp3D<T> *>(h); }            \
    extern "C" void oracle_range3d_grid_##SFX(void *h, long *row_parts, long *col_parts) {                            \
        auto *gp = static_cast<RangeSensorGp3D<T> *>(h);                                                              \
        *row_parts = static_cast<long>(gp->row_partitions.size());                                                    \
        *col_parts = static_cast<long>(gp->col_partitions.size());                                                    \
    }                                                                                                                 \
    extern "C" int oracle_range3d_train_##SFX(void *h, const T *ranges, const uint8_t *mask_hit, int frame_valid) {   \
        return static_cast<RangeSensorGp3D<T> *>(h)->Train(ranges, mask_hit, frame_valid != 0) ? 0 : -1;              \
    }                                                                                                                 \
    extern "C" int oracle_range3d_get_gp_##SFX(void *h, long g, int *trained, long *n, T *l, long ld_l, T *alpha) {   \
        auto *rg = static_cast<RangeSensorGp3D<T> *>(h);                                                              \
        const auto &gp = rg->gps[g];                                                                                  \
        *trained = gp.trained ? 1 : 0;                                                                                \
        *n = gp.trained ? gp.k_cols : gp.num_samples;                                                                 \
        if (!gp.trained) { return 0; }                                                                                \
        for (long c = 0; c < gp.k_cols; ++c) {                                                                        \
            for (long r = 0; r < gp.k_cols; ++r) {                                                                    \
                if (l != nullptr) { l[r + c * ld_l] = gp.mat_l[r + c * gp.ld]; }                                      \
            }                                                                                                         \
            if (alpha != nullptr) { alpha[c] = gp.mat_alpha[c]; }                                                     \
        }                                                                                                             \
        return 0;                                                                                                     \
    }                                                                                                                 \
    extern "C" int oracle_range3d_test_##SFX(void *h, const T *coords, const uint8_t *coords_ok, long num_test,       \
                                             int un_map, T *mean, T *variance, uint8_t *valid) {                      \
        return static_cast<RangeSensorGp3D<T> *>(h)->TestFrameCoords(coords, coords_ok, num_test, un_map != 0, mean,  \
                                                                     variance, valid)                                 \
                   ? 0                                                                                                \
                   : -1;                                                                                              \
    }                                                                                                                 \
    /* ---- batched independent GPs (config 4) ---- */                                                                \
    extern "C" int oracle_batched_train_predict_##SFX(int kernel, T scale, long x_dim, long num_gps, long max_n,      \
                                                      const int *n_train, const T *x, const T *y, const T *var,       \
                                                      const long *q_offsets, const T *q_x, T *mat_l, T *alpha,        \
                                                      int *info, T *mean, T *variance) {                              \
        return BatchedTrainPredict<T>(kernel, scale, x_dim, num_gps, max_n, n_train, x, y, var, q_offsets, q_x,       \
                                      mat_l, alpha, info, mean, variance);                                            \
    }                                                                                                                 \
    /* ---- NoisyInputGaussianProcess ---- */                                                                         \
    extern "C" void *oracle_noisy_create_##SFX(int kernel, T scale, int no_gradient_observation) {                    \
        auto *gp = new NoisyInputGp<T>();                                                                             \
        gp->kernel_type = kernel;                                                                                     \
        gp->scale = scale;                                                                                            \
        gp->no_gradient_observation = no_gradient_observation != 0;                                                   \
        return gp;                                                                                                    \
    }                                                                                                                 \
    extern "C" void oracle_noisy_destroy_##SFX(void *h) { delete static_cast<NoisyInputGp<T> *>(h); }                 \
    extern "C" long oracle_noisy_train_##SFX(void *h, long n, long x_dim, long y_dim, const T *x, const T *y,         \
                                             const T *grad, const T *var_x, const T *var_y, const T *var_grad,        \
                                             const long *grad_flag) {                                                 \
        auto *gp = static_cast<NoisyInputGp<T> *>(h);                                                                 \
        gp->x_dim = x_dim, gp->y_dim = y_dim, gp->num_samples = n;                                                    \
        gp->x.assign(x, x + n * x_dim);                                                                               \
        gp->y.assign(y, y + n * y_dim);                                                                               \
        gp->grad.assign(static_cast<std::size_t>(n * x_dim * y_dim), T(0));                                           \
        if (grad != nullptr) { gp->grad.assign(grad, grad + n * x_dim * y_dim); }                                     \
        gp->var_x.assign(var_x, var_x + n);                                                                           \
        gp->var_y.assign(var_y, var_y + n);                                                                           \
        gp->var_grad.assign(static_cast<std::size_t>(n), T(0));                                                       \
        if (var_grad != nullptr) { gp->var_grad.assign(var_grad, var_grad + n); }                                     \
        gp->grad_flag.assign(grad_flag, grad_flag + n);                                                               \
        if (!gp->Train()) { return -1; }                                                                              \
        return gp->m;                                                                                                 \
    }                                                                                                                 \
    extern "C" int oracle_noisy_get_##SFX(void *h, T *k, T *l, T *alpha) {                                            \
        auto *gp = static_cast<NoisyInputGp<T> *>(h);                                                                 \
        const std::size_t mm = static_cast<std::size_t>(gp->m * gp->m);                                               \
        if (k != nullptr) { std::memcpy(k, gp->mat_k.data(), mm * sizeof(T)); }                                       \
        if (l != nullptr) { std::memcpy(l, gp->mat_l.data(), mm * sizeof(T)); }                                       \
        if (alpha != nullptr) { std::memcpy(alpha, gp->mat_alpha.data(), gp->m * gp->y_dim * sizeof(T)); }            \
        return gp->info;                                                                                              \
    }                                                                                                                 \
    extern "C" int oracle_noisy_test_##SFX(void *h, const T *x_test, long num_test, int predict_gradient, T *mean,    \
                                           T *gradient, T *var, T *grad_var, T *cov) {                                \
        return static_cast<NoisyInputGp<T> *>(h)->Test(x_test, num_test, predict_gradient != 0, mean, gradient, var,  \
                                                       grad_var, cov)                                                 \
                   ? 0                                                                                                \
                   : -1;                                                                                              \
    }                                                                                                                 \
    /* ---- SPGP (dense) ---- */                                                                                      \
    extern "C" void *oracle_spgp_create_##SFX(int kernel, T scale, long x_dim, long m, const T *pseudo) {             \
        auto *gp = new Spgp<T>();                                                                                     \
        gp->kernel_type = kernel;                                                                                     \
        gp->scale = scale;                                                                                            \
        gp->Init(pseudo, x_dim, m);                                                                                   \
        return gp;                                                                                                    \
    }                                                                                                                 \
    extern "C" void oracle_spgp_destroy_##SFX(void *h) { delete static_cast<Spgp<T> *>(h); }                          \
    extern "C" int oracle_spgp_update_##SFX(void *h, const T *x, const T *y, const T *var, long n) {                  \
        return static_cast<Spgp<T> *>(h)->Update(x, y, var, n) ? 0 : -1;                                              \
    }                                                                                                                 \
    extern "C" int oracle_spgp_test_##SFX(void *h, const T *x_test, long num_test, T *mean, T *variance) {            \
        static_cast<Spgp<T> *>(h)->Test(x_test, num_test, mean, variance);                                            \
        return 0;                                                                                                     \
    }                                                                                                                 \
    extern "C" void oracle_spgp_set_diagonal_qm_##SFX(void *h, int on) {                                             \
        auto *gp = static_cast<Spgp<T> *>(h);                                                                         \
        gp->diagonal_qm = on != 0;                                                                                    \
        std::fill(gp->q_diag.begin(), gp->q_diag.end(), T(1));                                                        \
        std::fill(gp->alpha.begin(), gp->alpha.end(), T(0));                                                          \
        gp->q_m = gp->k_m;                                                                                            \
        gp->l_qm_updated = false;                                                                                     \
    }                                                                                                                 \
    extern "C" int oracle_spgp_get_qm_diagonal_##SFX(void *h, T *q) {                                                 \
        auto *gp = static_cast<Spgp<T> *>(h);                                                                         \
        std::memcpy(q, gp->q_diag.data(), gp->m * sizeof(T));                                                         \
        return 0;                                                                                                     \
    }                                                                                                                 \
    extern "C" int oracle_spgp_test_gradient_##SFX(void *h, const T *x_test, long num_test, T *grad, int raw_alpha) {  \
        static_cast<Spgp<T> *>(h)->TestGradient(x_test, num_test, grad, raw_alpha != 0);                              \
        return 0;                                                                                                     \
    }                                                                                                                 \
    extern "C" int oracle_spgp_get_##SFX(void *h, T *q_m, T *alpha, T *l_km, T *l_qm) {                               \
        auto *gp = static_cast<Spgp<T> *>(h);                                                                         \
        const std::size_t mm = static_cast<std::size_t>(gp->m * gp->m);                                               \
        if (q_m != nullptr) { std::memcpy(q_m, gp->q_m.data(), mm * sizeof(T)); }                                     \
        if (alpha != nullptr) { std::memcpy(alpha, gp->alpha.data(), gp->m * sizeof(T)); }                            \
        if (l_km != nullptr) { std::memcpy(l_km, gp->l_km.data(), mm * sizeof(T)); }                                  \
        if (l_qm != nullptr && gp->l_qm.size() == mm) { std::memcpy(l_qm, gp->l_qm.data(), mm * sizeof(T)); }         \
        return 0;                                                                                                     \
    }

ORACLE_EXPORTS(float, f32)
ORACLE_EXPORTS(double, f64)

extern "C" int
oracle_num_threads() {
    return omp_get_max_threads();
}

extern "C" void
oracle_set_num_threads(int n) {
    omp_set_num_threads(n);
}
