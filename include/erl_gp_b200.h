/*
 * erl_gp_b200.h — C ABI of the B200-native (sm_100a) train/predict hot path of
 * ExistentialRobotics/erl_gaussian_process.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry
 * point names the reference interface it replaces (file:line under the reference tree).
 * All matrices are COLUMN-MAJOR with explicit leading dimensions, exactly as the reference's
 * Eigen buffers: x is x_dim x n (one sample = x_dim contiguous scalars), K / L are the
 * top-left n x n of an ld x ld buffer, L's strict upper triangle is zero, Ktest is n x T.
 *
 * Conventions
 *   - every function returns an erl_gp_status (0 = OK) and never throws;
 *   - `_f32` / `_f64` suffix = Dtype float / double (the reference instantiates both);
 *   - functions without `_dev` take HOST pointers and return after the stream is drained
 *     (what the reference-side C++ classes call); `_dev` functions take DEVICE pointers on
 *     the context's device and are asynchronous on the context's stream;
 *   - nothing here falls back to the CPU: without a CUDA device every call fails with
 *     ERL_GP_STATUS_NO_DEVICE.
 */
#ifndef ERL_GP_B200_H_
#define ERL_GP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the library is built with -fvisibility=hidden; only this ABI is exported */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define ERL_GP_B200_VERSION 100

typedef enum erl_gp_status {
    ERL_GP_STATUS_OK = 0,
    ERL_GP_STATUS_INVALID_ARGUMENT = 1,
    ERL_GP_STATUS_CUDA_ERROR = 2,
    ERL_GP_STATUS_NOT_TRAINED = 3, /* Test() before Train(): the reference returns nullptr/false */
    ERL_GP_STATUS_UNSUPPORTED = 4, /* shape outside what the kernels cover (e.g. x_dim > 3) */
    ERL_GP_STATUS_NO_DEVICE = 5,
    ERL_GP_STATUS_ALLOC_FAILED = 6
} erl_gp_status;

/* erl_covariance kernel classes named by Setting::kernel_type */
typedef enum erl_gp_kernel {
    ERL_GP_KERNEL_OU = 0,       /* erl::covariance::OrnsteinUhlenbeck<Dtype, Dim> */
    ERL_GP_KERNEL_MATERN32 = 1, /* erl::covariance::Matern32<Dtype, Dim>          */
    ERL_GP_KERNEL_RBF = 2       /* erl::covariance::RadialBiasFunction<Dtype, Dim>*/
} erl_gp_kernel;

/* include/erl_gaussian_process/mapping.hpp:11-20 (same numeric values) */
typedef enum erl_gp_mapping {
    ERL_GP_MAPPING_NONE = -1, /* Test(..., un_map=false) */
    ERL_GP_MAPPING_IDENTITY = 0,
    ERL_GP_MAPPING_INVERSE = 1,
    ERL_GP_MAPPING_INVERSE_SQRT = 2,
    ERL_GP_MAPPING_EXP = 3,
    ERL_GP_MAPPING_LOG = 4,
    ERL_GP_MAPPING_TANH = 5,
    ERL_GP_MAPPING_SIGMOID = 6
} erl_gp_mapping;

typedef struct erl_gp_context erl_gp_context; /* device, stream, workspaces */

/* ------------------------------------------------------------------------------------------
 * Library / context
 * ---------------------------------------------------------------------------------------- */
int erl_gp_version(void);
const char *erl_gp_status_string(int status);
int erl_gp_device_count(int *count);
/* One context per host thread and device (the reference objects are not thread-safe either). */
int erl_gp_context_create(int device, erl_gp_context **ctx);
int erl_gp_context_destroy(erl_gp_context *ctx);
/* Borrow an external CUDA stream (cudaStream_t as void*, e.g. torch's current stream); the context's own
 * stream is released.  NULL selects the CUDA legacy default stream. */
int erl_gp_context_set_stream(erl_gp_context *ctx, void *cuda_stream);
int erl_gp_context_synchronize(erl_gp_context *ctx);
const char *erl_gp_context_last_error(const erl_gp_context *ctx);
/* Kernel choice for the fused FP32 train + predict of small GPs (n <= 128; replaces src/batch_gp_update_torch.cpp:74-82 and
 * the per-partition loops src/lidar_gp_2d.cpp:366-392): on != 0 selects the tcgen05 / TMEM kernel (csrc/erl_gp_rowgp_tc.cuh),
 * 0 the mma.sync kernel, a negative value the default (environment variable ERL_GP_ROWGP_TC, else the mma.sync kernel). */
int erl_gp_context_set_rowgp_tc(erl_gp_context *ctx, int on);
/* Number of this library's kernels launched through the context so far (bench.py `gpu_launches`). */
int erl_gp_context_kernel_launches(const erl_gp_context *ctx, long *count);

/* ------------------------------------------------------------------------------------------
 * Covariance::ComputeKtrain / ComputeKtest  (erl_covariance v0.2.0; call sites
 * src/vanilla_gp.cpp:486-487 and :537, src/sparse_pseudo_input_gp.cpp:340, 761-762).
 * Fused pairwise distance + kernel + noise diagonal; K[i,i] = 1 + var[i]; full square written.
 * ---------------------------------------------------------------------------------------- */
int erl_gp_compute_ktrain_f32(erl_gp_context *ctx, int kernel, float scale, long x_dim, const float *x, long ld_x,
                              const float *var, long n, float *k, long ld_k);
int erl_gp_compute_ktrain_f64(erl_gp_context *ctx, int kernel, double scale, long x_dim, const double *x, long ld_x,
                              const double *var, long n, double *k, long ld_k);
int erl_gp_compute_ktest_f32(erl_gp_context *ctx, int kernel, float scale, long x_dim, const float *x1, long ld_x1,
                             long n1, const float *x2, long ld_x2, long n2, float *k, long ld_k);
int erl_gp_compute_ktest_f64(erl_gp_context *ctx, int kernel, double scale, long x_dim, const double *x1, long ld_x1,
                             long n1, const double *x2, long ld_x2, long n2, double *k, long ld_k);
int erl_gp_compute_ktrain_dev_f32(erl_gp_context *ctx, int kernel, float scale, long x_dim, const float *x, long ld_x,
                                  const float *var, long n, float *k, long ld_k);
int erl_gp_compute_ktrain_dev_f64(erl_gp_context *ctx, int kernel, double scale, long x_dim, const double *x,
                                  long ld_x, const double *var, long n, double *k, long ld_k);
int erl_gp_compute_ktest_dev_f32(erl_gp_context *ctx, int kernel, float scale, long x_dim, const float *x1,
                                 long ld_x1, long n1, const float *x2, long ld_x2, long n2, float *k, long ld_k);
int erl_gp_compute_ktest_dev_f64(erl_gp_context *ctx, int kernel, double scale, long x_dim, const double *x1,
                                 long ld_x1, long n1, const double *x2, long ld_x2, long n2, double *k, long ld_k);

/* ------------------------------------------------------------------------------------------
 * VanillaGaussianProcess<Dtype>  (src/vanilla_gp.cpp) — one dense GP of any n, device resident.
 *   train  = UpdateKtrain (:476-490) + Solve (:492-505): K, L = chol(K) (blocked, right-looking,
 *            FP64 panels on DMMA), alpha = L^-T L^-1 y for y_dim right-hand sides.
 *   test   = ComputeKtest (:521-552) + TestResult::GetMean (:61-82) + GetVariance (:106-150):
 *            mean = Kt^T alpha, var = 1 - ||L^-1 k*||^2, tiled over test points so the n x T
 *            Ktest is never materialised beyond one tile.
 *   get    = materialise K / L / alpha into the caller's (Eigen) buffers on demand.
 * info: 0, or k > 0 when the leading minor of order k is not positive (Eigen's NumericalIssue,
 * which the reference ignores at :499).
 * ---------------------------------------------------------------------------------------- */
typedef struct erl_gp_vanilla_f32 erl_gp_vanilla_f32;
typedef struct erl_gp_vanilla_f64 erl_gp_vanilla_f64;

int erl_gp_vanilla_create_f32(erl_gp_context *ctx, erl_gp_vanilla_f32 **gp);
int erl_gp_vanilla_create_f64(erl_gp_context *ctx, erl_gp_vanilla_f64 **gp);
int erl_gp_vanilla_destroy_f32(erl_gp_vanilla_f32 *gp);
int erl_gp_vanilla_destroy_f64(erl_gp_vanilla_f64 *gp);
int erl_gp_vanilla_train_f32(erl_gp_vanilla_f32 *gp, int kernel, float scale, long x_dim, long y_dim, long n,
                             const float *x, long ld_x, const float *y, long ld_y, const float *var, int *info);
int erl_gp_vanilla_train_f64(erl_gp_vanilla_f64 *gp, int kernel, double scale, long x_dim, long y_dim, long n,
                             const double *x, long ld_x, const double *y, long ld_y, const double *var, int *info);
int erl_gp_vanilla_train_dev_f32(erl_gp_vanilla_f32 *gp, int kernel, float scale, long x_dim, long y_dim, long n,
                                 const float *x, long ld_x, const float *y, long ld_y, const float *var);
int erl_gp_vanilla_train_dev_f64(erl_gp_vanilla_f64 *gp, int kernel, double scale, long x_dim, long y_dim, long n,
                                 const double *x, long ld_x, const double *y, long ld_y, const double *var);
int erl_gp_vanilla_info_f32(erl_gp_vanilla_f32 *gp, int *info); /* synchronises */
int erl_gp_vanilla_info_f64(erl_gp_vanilla_f64 *gp, int *info);
int erl_gp_vanilla_get_f32(erl_gp_vanilla_f32 *gp, float *k, long ld_k, float *l, long ld_l, float *alpha, long ld_a);
int erl_gp_vanilla_get_f64(erl_gp_vanilla_f64 *gp, double *k, long ld_k, double *l, long ld_l, double *alpha,
                           long ld_a);
/* mean: num_test x y_dim col-major (ld = num_test) or NULL; var: num_test or NULL */
int erl_gp_vanilla_test_f32(erl_gp_vanilla_f32 *gp, long num_test, const float *x_test, long ld_xt, float *mean,
                            float *var);
int erl_gp_vanilla_test_f64(erl_gp_vanilla_f64 *gp, long num_test, const double *x_test, long ld_xt, double *mean,
                            double *var);
/* One process, several GPUs (north_star: "sharded across the 8 GPUs of one box with only a host gather"; the factorisation is
 * single-GPU, the predict of src/vanilla_gp.cpp:521-559, 61-150 shards over test points).  replicate: copy the trained state of
 * `src` (x_train, L, alpha) to `dst`, a GP created on another context / device (cudaMemcpyPeerAsync: NVLink with peer access)
 * instead of one redundant factorisation per device.  test_multi: Test() + GetMean + GetVariance over contiguous ranges of the
 * test points, one host thread per replica, results written straight into the caller's arrays. */
int erl_gp_vanilla_replicate_f32(erl_gp_vanilla_f32 *src, erl_gp_vanilla_f32 *dst);
int erl_gp_vanilla_replicate_f64(erl_gp_vanilla_f64 *src, erl_gp_vanilla_f64 *dst);
int erl_gp_vanilla_test_multi_f32(erl_gp_vanilla_f32 *const *gps, long num_gps, long num_test, const float *x_test,
                                  long ld_xt, float *mean, float *var);
int erl_gp_vanilla_test_multi_f64(erl_gp_vanilla_f64 *const *gps, long num_gps, long num_test, const double *x_test,
                                  long ld_xt, double *mean, double *var);
int erl_gp_vanilla_test_dev_f32(erl_gp_vanilla_f32 *gp, long num_test, const float *x_test, long ld_xt, float *mean,
                                float *var);
int erl_gp_vanilla_test_dev_f64(erl_gp_vanilla_f64 *gp, long num_test, const double *x_test, long ld_xt,
                                double *mean, double *var);

/* ------------------------------------------------------------------------------------------
 * Batched small GPs — one CTA per GP (n <= 256 float, n <= 192 double).
 * Replaces the per-partition OpenMP loops (src/lidar_gp_2d.cpp:366-392,
 * src/range_sensor_gp_3d.cpp:334-360) and BatchGaussianProcessUpdateTorch
 * (src/batch_gp_update_torch.cpp:10-98; semantics of VanillaGaussianProcess::Solve, i.e. the
 * K^-1-applied-twice bug of :76-78 is NOT reproduced).
 *
 * Layout of a batch of B GPs with capacity max_n:
 *   n_train int32[B]                      samples per GP (ragged)
 *   x       Dtype[B][max_n][x_dim]        == each GP's x_dim x max_n col-major TrainSet::x
 *   y, var  Dtype[B][max_n]
 *   L       Dtype[B][max_n*max_n]         col-major, ld = max_n, strict upper zero, top-left n x n
 *   alpha   Dtype[B][max_n]
 *   info    int32[B]                      0 trained; k>0 LLT failed at column k; -1 not trained
 *                                         (n <= min_num_samples, the reference's `cnt > min` gate)
 * Queries are grouped per GP (CSR): q_offsets int64[B+1], q_x Dtype[T][x_dim].
 * Outputs for queries of untrained GPs are left untouched and flagged valid = 0, as the
 * reference leaves them unwritten (src/lidar_gp_2d.cpp:112,120).
 * ---------------------------------------------------------------------------------------- */
typedef struct erl_gp_batch_f32 erl_gp_batch_f32;
typedef struct erl_gp_batch_f64 erl_gp_batch_f64;

int erl_gp_batch_create_f32(erl_gp_context *ctx, long num_gps, long max_n, long x_dim, int kernel, float scale,
                            erl_gp_batch_f32 **batch);
int erl_gp_batch_create_f64(erl_gp_context *ctx, long num_gps, long max_n, long x_dim, int kernel, double scale,
                            erl_gp_batch_f64 **batch);
int erl_gp_batch_destroy_f32(erl_gp_batch_f32 *batch);
int erl_gp_batch_destroy_f64(erl_gp_batch_f64 *batch);
/* Device buffers owned by the batch (for callers that fill them on the device, e.g. bench.py). */
int erl_gp_batch_device_buffers_f32(erl_gp_batch_f32 *batch, int **n_train, float **x, float **y, float **var,
                                    float **l, float **alpha, int **info);
int erl_gp_batch_device_buffers_f64(erl_gp_batch_f64 *batch, int **n_train, double **x, double **y, double **var,
                                    double **l, double **alpha, int **info);
/* Host -> device upload of the training sets. */
int erl_gp_batch_upload_f32(erl_gp_batch_f32 *batch, const int *n_train, const float *x, const float *y,
                            const float *var);
int erl_gp_batch_upload_f64(erl_gp_batch_f64 *batch, const int *n_train, const double *x, const double *y,
                            const double *var);
/* Train every GP with n_train > min_num_samples from the resident buffers (async). write_l = 0
 * skips the L write-back ("fused, no L write-back" mode of SURVEY.md 8d). */
int erl_gp_batch_train_dev_f32(erl_gp_batch_f32 *batch, long min_num_samples, int write_l);
int erl_gp_batch_train_dev_f64(erl_gp_batch_f64 *batch, long min_num_samples, int write_l);
/* Predict from resident L / alpha (async). q_out_index (int32[T] or NULL) scatters result i to
 * mean[q_out_index[i]] — used after the device-side ray->partition bucketing. */
int erl_gp_batch_predict_dev_f32(erl_gp_batch_f32 *batch, const long *q_offsets, const float *q_x,
                                 const int *q_out_index, long num_q, int mapping, float mapping_scale, float *mean,
                                 float *var, uint8_t *valid);
int erl_gp_batch_predict_dev_f64(erl_gp_batch_f64 *batch, const long *q_offsets, const double *q_x,
                                 const int *q_out_index, long num_q, int mapping, double mapping_scale, double *mean,
                                 double *var, uint8_t *valid);
/* Fused train + predict in ONE kernel (L stays in shared memory between the two phases). */
int erl_gp_batch_train_predict_dev_f32(erl_gp_batch_f32 *batch, long min_num_samples, int write_l,
                                       const long *q_offsets, const float *q_x, long num_q, float *mean, float *var,
                                       uint8_t *valid);
int erl_gp_batch_train_predict_dev_f64(erl_gp_batch_f64 *batch, long min_num_samples, int write_l,
                                       const long *q_offsets, const double *q_x, long num_q, double *mean,
                                       double *var, uint8_t *valid);
/* Whole path with HOST buffers: upload, fused train+predict, download (the e2e call).
 * l / alpha / info / valid may be NULL to skip their download. */
int erl_gp_batch_train_predict_f32(erl_gp_batch_f32 *batch, long min_num_samples, const int *n_train,
                                   const float *x, const float *y, const float *var, const long *q_offsets,
                                   const float *q_x, long num_q, float *l, float *alpha, int *info, float *mean,
                                   float *variance, uint8_t *valid);
int erl_gp_batch_train_predict_f64(erl_gp_batch_f64 *batch, long min_num_samples, const int *n_train,
                                   const double *x, const double *y, const double *var, const long *q_offsets,
                                   const double *q_x, long num_q, double *l, double *alpha, int *info, double *mean,
                                   double *variance, uint8_t *valid);
/* The same call over SEVERAL GPUs from one process (north_star: "sharded across the 8 GPUs of one box with only a host
 * gather"; the loop it replaces, src/lidar_gp_2d.cpp:366-392 / src/range_sensor_gp_3d.cpp:334-360, has no dependency
 * between partitions).  batches[i] was created on a context of device i with its share of the GPs; batch i takes the next
 * batches[i]->num_gps GPs of the arrays (contiguous ranges, sum = the number of GPs described by q_offsets).  One host
 * thread per batch drives that device; every device writes straight into the caller's arrays.  Host buffers should be
 * page-locked (cudaHostRegister / cudaHostAlloc) for the copies to overlap. */
int erl_gp_batch_train_predict_multi_f32(erl_gp_batch_f32 *const *batches, long num_batches, long min_num_samples,
                                         const int *n_train, const float *x, const float *y, const float *var,
                                         const long *q_offsets, const float *q_x, long num_q, float *l, float *alpha,
                                         int *info, float *mean, float *variance, uint8_t *valid);
int erl_gp_batch_train_predict_multi_f64(erl_gp_batch_f64 *const *batches, long num_batches, long min_num_samples,
                                         const int *n_train, const double *x, const double *y, const double *var,
                                         const long *q_offsets, const double *q_x, long num_q, double *l, double *alpha,
                                         int *info, double *mean, double *variance, uint8_t *valid);
/* Device -> host download of results; any pointer may be NULL. */
int erl_gp_batch_download_f32(erl_gp_batch_f32 *batch, float *l, float *alpha, int *info);
int erl_gp_batch_download_f64(erl_gp_batch_f64 *batch, double *l, double *alpha, int *info);
/* One GP's state (what LidarGaussianProcess2D::GetGps()[p] exposes). l is n x n with ld_l. */
int erl_gp_batch_get_gp_f32(erl_gp_batch_f32 *batch, long gp_index, int *info, long *n, float *l, long ld_l,
                            float *alpha);
int erl_gp_batch_get_gp_f64(erl_gp_batch_f64 *batch, long gp_index, int *info, long *n, double *l, long ld_l,
                            double *alpha);

/* ------------------------------------------------------------------------------------------
 * LidarGaussianProcess2D<Dtype>  (src/lidar_gp_2d.cpp).  erl_geometry::LidarFrame2D stays on
 * the caller's side; its outputs (angles in frame, valid ranges, hit / continuity masks, the
 * sensor rotation) are the inputs here.
 *   create : PartitionOnAngles (:238-300) on the host, table uploaded once
 *   train  : StoreData mapping (:227-236) + per-partition gather (:379-389) + batched train
 *   test   : world->frame angle (:69-75), SearchPartition (:398-411), bucketing, batched predict,
 *            GetMean with Mapping::inv (:102-126), GetVariance (:128-167)
 * ---------------------------------------------------------------------------------------- */
typedef struct erl_gp_lidar2d_setting {
    int symmetric_partitions; /* Setting::symmetric_partitions (lidar_gp_2d.hpp:33) */
    long group_size;          /* :35 */
    long overlap_size;        /* :37 */
    long margin;              /* :40 */
    double sensor_range_var;  /* :44 */
    double discontinuity_var; /* :47 */
    int discontinuity_detection; /* sensor_frame->discontinuity_detection, src/lidar_gp_2d.cpp:376 */
    int kernel;               /* gp->kernel_type */
    double kernel_scale;      /* gp->kernel->scale */
    int mapping;              /* mapping->type, default kInverseSqrt (:57-62) */
    double mapping_scale;     /* mapping->scale */
    int partition_on_hit_rays; /* Setting::partition_on_hit_rays (:31): the table is rebuilt from the hit rays of every
                                  Train() (src/lidar_gp_2d.cpp:302-348, 364); out-of-range indices of the reference clamped */
} erl_gp_lidar2d_setting;

typedef struct erl_gp_lidar2d_f32 erl_gp_lidar2d_f32;
typedef struct erl_gp_lidar2d_f64 erl_gp_lidar2d_f64;

int erl_gp_lidar2d_create_f32(erl_gp_context *ctx, const erl_gp_lidar2d_setting *setting, const float *angles,
                              long num_rays, erl_gp_lidar2d_f32 **gp);
int erl_gp_lidar2d_create_f64(erl_gp_context *ctx, const erl_gp_lidar2d_setting *setting, const double *angles,
                              long num_rays, erl_gp_lidar2d_f64 **gp);
int erl_gp_lidar2d_destroy_f32(erl_gp_lidar2d_f32 *gp);
int erl_gp_lidar2d_destroy_f64(erl_gp_lidar2d_f64 *gp);
int erl_gp_lidar2d_num_partitions_f32(erl_gp_lidar2d_f32 *gp, long *num);
int erl_gp_lidar2d_num_partitions_f64(erl_gp_lidar2d_f64 *gp, long *num);
/* GetAnglePartitions(): (index_left, index_right, coord_left, coord_right) per partition */
int erl_gp_lidar2d_partitions_f32(erl_gp_lidar2d_f32 *gp, long *index_left, long *index_right, float *coord_left,
                                  float *coord_right);
int erl_gp_lidar2d_partitions_f64(erl_gp_lidar2d_f64 *gp, long *index_left, long *index_right, double *coord_left,
                                  double *coord_right);
/* rotation: 2x2 col-major sensor->world; ranges / masks: num_rays (frame outputs). */
int erl_gp_lidar2d_train_f32(erl_gp_lidar2d_f32 *gp, const float *rotation, const float *ranges,
                             const uint8_t *mask_hit, const uint8_t *mask_continuous);
int erl_gp_lidar2d_train_f64(erl_gp_lidar2d_f64 *gp, const double *rotation, const double *ranges,
                             const uint8_t *mask_hit, const uint8_t *mask_continuous);
int erl_gp_lidar2d_test_f32(erl_gp_lidar2d_f32 *gp, const float *angles, long num_test, int angles_are_local,
                            int un_map, float *mean, float *var, uint8_t *valid);
int erl_gp_lidar2d_test_f64(erl_gp_lidar2d_f64 *gp, const double *angles, long num_test, int angles_are_local,
                            int un_map, double *mean, double *var, uint8_t *valid);
/* partition GP p: info (0 trained / -1 untrained / k>0 failed), n, L (n x n, ld_l), alpha */
int erl_gp_lidar2d_get_gp_f32(erl_gp_lidar2d_f32 *gp, long p, int *info, long *n, float *l, long ld_l, float *alpha);
int erl_gp_lidar2d_get_gp_f64(erl_gp_lidar2d_f64 *gp, long p, int *info, long *n, double *l, long ld_l,
                              double *alpha);
/* Batched ComputeOcc (:428-459): pos 2 x T col-major in the sensor frame.  ok[i] = 0 where the
 * reference returns false (no partition / untrained / var > max_valid_range_var). */
int erl_gp_lidar2d_compute_occ_f32(erl_gp_lidar2d_f32 *gp, const float *pos_local, long num, float max_valid_range_var,
                                   float occ_test_temperature, float *dist, float *range_pred, float *occ,
                                   uint8_t *ok);
int erl_gp_lidar2d_compute_occ_f64(erl_gp_lidar2d_f64 *gp, const double *pos_local, long num,
                                   double max_valid_range_var, double occ_test_temperature, double *dist,
                                   double *range_pred, double *occ, uint8_t *ok);

/* ------------------------------------------------------------------------------------------
 * RangeSensorGaussianProcess3D<Dtype>  (src/range_sensor_gp_3d.cpp).  The
 * erl_geometry::RangeSensorFrame3D outputs are the inputs: frame_coords (rows x cols of
 * Vector2, Eigen col-major: element (r,c) at ((r + c*rows)*2 + k)), valid ranges, hit mask;
 * queries are the frame coordinates returned by ComputeFrameCoords plus its bool.
 *   create : row / col partition tables (:199-259); GP grid (row_part, col_part) col-major
 *   train  : gather col-outer / row-inner (:348-356), train iff cnt > min_num_samples_per_group
 *   test   : SearchPartition (:366-393; row [l,r), col [l,r]), bucketing, batched predict
 * ---------------------------------------------------------------------------------------- */
typedef struct erl_gp_range3d_setting {
    long row_group_size, row_overlap_size, row_margin; /* range_sensor_gp_3d.hpp:33-36 */
    long col_group_size, col_overlap_size, col_margin; /* :38-41 */
    long min_num_samples_per_group;                    /* :43 */
    double sensor_range_var;                           /* :47 */
    int kernel;
    double kernel_scale;
    int mapping;
    double mapping_scale;
} erl_gp_range3d_setting;

typedef struct erl_gp_range3d_f32 erl_gp_range3d_f32;
typedef struct erl_gp_range3d_f64 erl_gp_range3d_f64;

int erl_gp_range3d_create_f32(erl_gp_context *ctx, const erl_gp_range3d_setting *setting, const float *frame_coords,
                              long rows, long cols, erl_gp_range3d_f32 **gp);
int erl_gp_range3d_create_f64(erl_gp_context *ctx, const erl_gp_range3d_setting *setting,
                              const double *frame_coords, long rows, long cols, erl_gp_range3d_f64 **gp);
int erl_gp_range3d_destroy_f32(erl_gp_range3d_f32 *gp);
int erl_gp_range3d_destroy_f64(erl_gp_range3d_f64 *gp);
int erl_gp_range3d_grid_f32(erl_gp_range3d_f32 *gp, long *num_row_partitions, long *num_col_partitions);
int erl_gp_range3d_grid_f64(erl_gp_range3d_f64 *gp, long *num_row_partitions, long *num_col_partitions);
/* axis: 0 = GetRowPartitions(), 1 = GetColPartitions() */
int erl_gp_range3d_partitions_f32(erl_gp_range3d_f32 *gp, int axis, long *index_left, long *index_right,
                                  float *coord_left, float *coord_right);
int erl_gp_range3d_partitions_f64(erl_gp_range3d_f64 *gp, int axis, long *index_left, long *index_right,
                                  double *coord_left, double *coord_right);
/* ranges, mask_hit: rows x cols col-major */
int erl_gp_range3d_train_f32(erl_gp_range3d_f32 *gp, const float *ranges, const uint8_t *mask_hit);
int erl_gp_range3d_train_f64(erl_gp_range3d_f64 *gp, const double *ranges, const uint8_t *mask_hit);
/* coords: 2 x T col-major frame coordinates; coords_ok (or NULL) = ComputeFrameCoords' return */
int erl_gp_range3d_test_f32(erl_gp_range3d_f32 *gp, const float *coords, const uint8_t *coords_ok, long num_test,
                            int un_map, float *mean, float *var, uint8_t *valid);
int erl_gp_range3d_test_f64(erl_gp_range3d_f64 *gp, const double *coords, const uint8_t *coords_ok, long num_test,
                            int un_map, double *mean, double *var, uint8_t *valid);
/* Batched ComputeOcc (src/range_sensor_gp_3d.cpp:409-439).  coords (2 x num), coords_ok and dist are what
 * RangeSensorFrame3D::ComputeFrameCoords (+ CoordsIsInFrame) returns for the positions; ok[i] = 0 where the reference
 * returns false (bad coords / no partition / untrained / var > max_valid_range_var); range_pred / occ of those stay
 * untouched. */
int erl_gp_range3d_compute_occ_f32(erl_gp_range3d_f32 *gp, const float *coords, const uint8_t *coords_ok, const float *dist,
                                   long num, float max_valid_range_var, float occ_test_temperature, float *range_pred,
                                   float *occ, uint8_t *ok);
int erl_gp_range3d_compute_occ_f64(erl_gp_range3d_f64 *gp, const double *coords, const uint8_t *coords_ok,
                                   const double *dist, long num, double max_valid_range_var, double occ_test_temperature,
                                   double *range_pred, double *occ, uint8_t *ok);
int erl_gp_range3d_get_gp_f32(erl_gp_range3d_f32 *gp, long row_part, long col_part, int *info, long *n, float *l,
                              long ld_l, float *alpha);
int erl_gp_range3d_get_gp_f64(erl_gp_range3d_f64 *gp, long row_part, long col_part, int *info, long *n, double *l,
                              long ld_l, double *alpha);

/* ------------------------------------------------------------------------------------------
 * SparsePseudoInputGaussianProcess<Dtype>, dense mode (src/sparse_pseudo_input_gp.cpp):
 *   create : K_M = ComputeKtest(Z,Z), L_KM = chol(K_M), Q_M = K_M, alpha = 0     (:313-356)
 *   update : K_MN, beta = L_KM^-1 K_MN, lambda, Q_M += Ks K_MN^T, alpha += Ks y  (:751-791)
 *   test   : L_QM = chol(Q_M) lazily (:835-842); mean = Kt^T Q_M^-1 alpha;
 *            var = 1 - ||L_KM^-1 Kt||^2 + ||L_QM^-1 Kt||^2                        (:43-113, :280-310)
 * ---------------------------------------------------------------------------------------- */
typedef struct erl_gp_spgp_f32 erl_gp_spgp_f32;
typedef struct erl_gp_spgp_f64 erl_gp_spgp_f64;

int erl_gp_spgp_create_f32(erl_gp_context *ctx, int kernel, float scale, long x_dim, long num_pseudo,
                           const float *pseudo_points, erl_gp_spgp_f32 **gp);
int erl_gp_spgp_create_f64(erl_gp_context *ctx, int kernel, double scale, long x_dim, long num_pseudo,
                           const double *pseudo_points, erl_gp_spgp_f64 **gp);
int erl_gp_spgp_destroy_f32(erl_gp_spgp_f32 *gp);
int erl_gp_spgp_destroy_f64(erl_gp_spgp_f64 *gp);
int erl_gp_spgp_update_f32(erl_gp_spgp_f32 *gp, long n, const float *x, long ld_x, const float *y, const float *var);
int erl_gp_spgp_update_f64(erl_gp_spgp_f64 *gp, long n, const double *x, long ld_x, const double *y,
                           const double *var);
int erl_gp_spgp_test_f32(erl_gp_spgp_f32 *gp, long num_test, const float *x_test, long ld_xt, float *mean,
                         float *var);
int erl_gp_spgp_test_f64(erl_gp_spgp_f64 *gp, long num_test, const double *x_test, long ld_xt, double *mean,
                         double *var);
/* Setting::diagonal_qm (sparse_pseudo_input_gp.hpp:57): Q_M is kept as its diagonal (ctor :346-347 ones(M), update :775-776,
 * alpha / diag(Q_M) at test time :100-101).  Call before the first update (Q_M and alpha are reset).  Mean and gradient only: the
 * reference's variance solves with an L_QM this mode never builds (:304-310, :839), erl_gp_spgp_test_* rejects var != NULL. */
int erl_gp_spgp_set_diagonal_qm_f32(erl_gp_spgp_f32 *gp, int on);
int erl_gp_spgp_set_diagonal_qm_f64(erl_gp_spgp_f64 *gp, int on);
int erl_gp_spgp_get_qm_diagonal_f32(erl_gp_spgp_f32 *gp, float *q);
int erl_gp_spgp_get_qm_diagonal_f64(erl_gp_spgp_f64 *gp, double *q);
/* TestResult::GetGradient (src/sparse_pseudo_input_gp.cpp:187-278): gradient of the predictive mean, x_dim x num_test col-major.
 * raw_alpha = 0: dotted with Q_M^-1 alpha like GetMean and the per-index accessor (:252); raw_alpha = 1: with the unsolved alpha,
 * as the reference's batched accessor does (:212).  RBF and Matern32 only. */
int erl_gp_spgp_test_gradient_f32(erl_gp_spgp_f32 *gp, long num_test, const float *x_test, long ld_xt, float *grad,
                                  int raw_alpha);
int erl_gp_spgp_test_gradient_f64(erl_gp_spgp_f64 *gp, long num_test, const double *x_test, long ld_xt, double *grad,
                                  int raw_alpha);
/* Read() (src/sparse_pseudo_input_gp.cpp:721-740): restores the accumulated Q_M (M x M col-major; M values in diagonal_qm mode)
 * and alpha (M) of a saved SPGP into a handle created with the same pseudo-points; L_QM is refactored at the next test. */
int erl_gp_spgp_set_state_f32(erl_gp_spgp_f32 *gp, const float *q_m, const float *alpha);
int erl_gp_spgp_set_state_f64(erl_gp_spgp_f64 *gp, const double *q_m, const double *alpha);
/* Q_M, L_KM, L_QM: M x M col-major (ld = M); alpha: M.  Any may be NULL. */
int erl_gp_spgp_get_f32(erl_gp_spgp_f32 *gp, float *q_m, float *alpha, float *l_km, float *l_qm);
int erl_gp_spgp_get_f64(erl_gp_spgp_f64 *gp, double *q_m, double *alpha, double *l_km, double *l_qm);

#if defined(__GNUC__)
/* ---------------------------------------------------------------------------------------------
 * NoisyInputGaussianProcess<Dtype> (include/erl_gaussian_process/noisy_input_gp.hpp, src/noisy_input_gp.cpp): a GP with
 * noisy inputs and gradient observations.  Row layout of K / L / alpha (m = n + x_dim * ng rows, ng = samples with
 * grad_flag != 0): rows [0, n) the values, row n + j + a * ng the derivative d/dx_a at the j-th flagged sample
 * (src/noisy_input_gp.cpp:835-846).  Kernels: RadialBiasFunction and Matern32 (OrnsteinUhlenbeck only with
 * no_gradient_observation and without gradient prediction: it is not differentiable at r = 0).
 *   train  = Reset + TrainSet fill + UpdateKtrain + Train (:807-899).  x: x_dim x n (ld_x); y: n x y_dim (ld_y);
 *            grad: (x_dim * y_dim) x n (ld_grad), column i = [dh_1/dx_1 .. dh_1/dx_m, dh_2/dx_1 ..] at sample i
 *            (noisy_input_gp.hpp:175-177); var_x / var_y / var_grad: n; grad_flag: n (Eigen::VectorXl).
 *   get    = GetKtrainSized / GetCholeskyDecomposition / GetAlphaSized (:766-800); info = 0 or the failing LLT column.
 *   test   = Test + TestResult::{GetMean :125-145, GetGradient :168-207, GetMeanVariance :233-246,
 *            GetGradientVariance :258-277, GetCovariance :300-333}.  Host pointers, any output may be null:
 *            mean T x y_dim; grad x_dim x T x y_dim; var T; grad_var x_dim x T; cov x_dim (x_dim + 1) / 2 x T.
 * ------------------------------------------------------------------------------------------- */
typedef struct erl_gp_noisy_f32 erl_gp_noisy_f32;
typedef struct erl_gp_noisy_f64 erl_gp_noisy_f64;

int erl_gp_noisy_create_f32(erl_gp_context *ctx, erl_gp_noisy_f32 **gp);
int erl_gp_noisy_create_f64(erl_gp_context *ctx, erl_gp_noisy_f64 **gp);
int erl_gp_noisy_destroy_f32(erl_gp_noisy_f32 *gp);
int erl_gp_noisy_destroy_f64(erl_gp_noisy_f64 *gp);
int erl_gp_noisy_train_f32(erl_gp_noisy_f32 *gp, int kernel, float scale, long x_dim, long y_dim, long n, const float *x,
                           long ld_x, const float *y, long ld_y, const float *grad, long ld_grad, const float *var_x,
                           const float *var_y, const float *var_grad, const long *grad_flag,
                           int no_gradient_observation);
int erl_gp_noisy_train_f64(erl_gp_noisy_f64 *gp, int kernel, double scale, long x_dim, long y_dim, long n,
                           const double *x, long ld_x, const double *y, long ld_y, const double *grad, long ld_grad,
                           const double *var_x, const double *var_y, const double *var_grad, const long *grad_flag,
                           int no_gradient_observation);
int erl_gp_noisy_get_f32(erl_gp_noisy_f32 *gp, long *rows, int *info, float *k, long ld_k, float *l, long ld_l,
                         float *alpha, long ld_a);
int erl_gp_noisy_get_f64(erl_gp_noisy_f64 *gp, long *rows, int *info, double *k, long ld_k, double *l, long ld_l,
                         double *alpha, long ld_a);
int erl_gp_noisy_test_f32(erl_gp_noisy_f32 *gp, long num_test, const float *x_test, long ld_xt, int predict_gradient,
                          float *mean, float *grad, float *var, float *grad_var, float *cov);
int erl_gp_noisy_test_f64(erl_gp_noisy_f64 *gp, long num_test, const double *x_test, long ld_xt, int predict_gradient,
                          double *mean, double *grad, double *var, double *grad_var, double *cov);

#pragma GCC visibility pop
#endif

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* ERL_GP_B200_H_ */
