// LidarGaussianProcess2D / RangeSensorGaussianProcess3D on top of the batched small-GP kernels.
//
//   partition tables      src/lidar_gp_2d.cpp:238-300, src/range_sensor_gp_3d.cpp:199-259   (host, once)
//   per-partition gather  src/lidar_gp_2d.cpp:379-389, src/range_sensor_gp_3d.cpp:348-356   (GatherKernel)
//   ray -> partition      src/lidar_gp_2d.cpp:69-79, 398-411; src/range_sensor_gp_3d.cpp:366-393
//                         (AssignKernel + device counting sort; replaces the serial per-ray loop with a
//                          heap allocation per ray of the reference's TestResult constructors)
//   mean / variance       batched predict kernel with the scatter index
//   ComputeOcc            src/lidar_gp_2d.cpp:428-459 (batched over positions)
#include "erl_gp_internal.cuh"

#include <algorithm>
#include <limits>

namespace erl_gp {

    template<typename T>
    struct PartitionTable {
        std::vector<long> index_left, index_right;
        std::vector<T> coord_left, coord_right;

        bool bad = false;  // a coordinate index fell outside the frame while the table was built

        long
        Size() const {
            return static_cast<long>(index_left.size());
        }

        void
        Add(long il, long ir, T cl, T cr) {
            index_left.push_back(il);
            index_right.push_back(ir);
            coord_left.push_back(cl);
            coord_right.push_back(cr);
        }
    };

    // coords(i) = coords[i * stride]
    template<typename T>
    static PartitionTable<T>
    MakePartitionTable(const T *coords, long stride, long n, long group_size, long overlap_size, long margin, bool symmetric) {
        PartitionTable<T> t;
        // the reference indexes its coordinate vector unchecked (latent out-of-bounds reads for margin >= n or
        // overlap_size > group_size - overlap_size); here an index outside [0, n) marks the table as bad instead
        auto c = [&](long i) {
            if (i < 0 || i >= n) {
                t.bad = true;
                i = i < 0 ? 0 : n - 1;
            }
            return coords[i * stride];
        };
        const long step = group_size - overlap_size;
        const long num_groups = std::max(1l, n / step) + 1;
        const long gs2 = (n - (num_groups - 2) * step) / 2;
        const long half_overlap = overlap_size / 2;
        if (symmetric) {
            t.Add(0, gs2 + half_overlap, c(margin), c(gs2));
            for (long i = 0; i < num_groups - 2; ++i) {
                const long il = i * step + gs2 - half_overlap;
                const long ir = il + group_size;
                t.Add(il, ir, c(il + half_overlap), c(ir - half_overlap));
            }
            t.Add(n - gs2 - half_overlap, n, c(n - 1 - gs2), c(n - 1 - margin));
            return t;
        }
        for (long i = 0; i < num_groups - 2; ++i) {
            const long il = i * step;
            const long ir = il + group_size;
            t.Add(il, ir, c(il), c(ir - half_overlap));
        }
        long il = (num_groups - 2) * step;
        long ir = il + (n - il + overlap_size) / 2;
        t.Add(il, ir, c(il), c(ir - half_overlap));
        il = il + (n - il - overlap_size) / 2;
        t.Add(il, n, c(il), c(n - 1));
        return t;
    }

    // LidarGaussianProcess2D::PartitionOnHitRays, src/lidar_gp_2d.cpp:302-348: the partitions follow the HIT rays of the current
    // frame (n = number of hit rays; the symmetric rule is not implemented there either, :322-324).  The reference indexes
    // hit_ray_indices with group_size-strided positions that reach n (and, for the last group, with an index that is already an
    // ORIGINAL ray index, :341-343) and reads angles[hit_ray_indices[n - 1] + 1] (:344-347), i.e. past the end whenever the last
    // ray is a hit.  Here every index into hit_ray_indices is clamped to [0, n - 1] and every index into angles to [0, N - 1];
    // wherever the reference stays in bounds the table is the reference's.
    template<typename T>
    static PartitionTable<T>
    MakeHitRayPartitionTable(const T *angles, long num_rays, const uint8_t *hit, long group_size, long overlap_size) {
        PartitionTable<T> t;
        std::vector<long> hit_idx;
        hit_idx.reserve(static_cast<std::size_t>(num_rays));
        for (long i = 0; i < num_rays; ++i) {
            if (hit[i] != 0) { hit_idx.push_back(i); }
        }
        const long n = static_cast<long>(hit_idx.size());
        if (n == 0) { return t; }
        auto h = [&](long i) { return hit_idx[static_cast<std::size_t>(i < 0 ? 0 : (i > n - 1 ? n - 1 : i))]; };
        auto a = [&](long i) { return angles[i < 0 ? 0 : (i > num_rays - 1 ? num_rays - 1 : i)]; };
        const long step = group_size - overlap_size;
        const long num_groups = std::max(1l, n / step) + 1;
        for (long i = 0; i < num_groups - 2; ++i) {
            const long il = h(i * step);
            const long ir = h(i * step + group_size);
            t.Add(il, ir, a(il), a(ir));
        }
        long il = (num_groups - 2) * step;
        long ir = il + (n - il + overlap_size) / 2;
        il = h(il);
        ir = h(ir);
        t.Add(il, ir, a(il), a(ir));
        il = il + (n - il - overlap_size) / 2;  // (the reference mixes the two index spaces here, :341)
        il = h(il);
        ir = h(n - 1) + 1;
        t.Add(il, ir, a(il), a(ir));
        return t;
    }

    template<typename T>
    static bool
    PartitionsFit(const PartitionTable<T> &t, long n, long capacity) {
        if (t.bad) { return false; }
        for (long i = 0; i < t.Size(); ++i) {
            if (t.index_left[i] < 0 || t.index_right[i] > n || t.index_right[i] - t.index_left[i] > capacity || t.index_right[i] < t.index_left[i]) { return false; }
        }
        return true;
    }

    // device copy of a partition table (int32 indices are enough: rays / pixels per axis)
    template<typename T>
    struct DevicePartitions {
        DeviceBuffer<int> il, ir;
        DeviceBuffer<T> cl, cr;
        int count = 0;

        // capacity > t.Size(): the table is padded with empty partitions no coordinate can match (hit-ray partitioning: the
        // number of partitions changes from frame to frame, the batch of partition GPs does not)
        cudaError_t
        Upload(const PartitionTable<T> &t, cudaStream_t stream, long capacity = 0) {
            count = static_cast<int>(std::max(capacity, t.Size()));
            std::vector<int> hil(t.index_left.begin(), t.index_left.end()), hir(t.index_right.begin(), t.index_right.end());
            std::vector<T> hcl(t.coord_left), hcr(t.coord_right);
            hil.resize(count, 0), hir.resize(count, 0);
            hcl.resize(count, std::numeric_limits<T>::infinity()), hcr.resize(count, -std::numeric_limits<T>::infinity());
            cudaError_t err = il.Reserve(count);
            if (err == cudaSuccess) { err = ir.Reserve(count); }
            if (err == cudaSuccess) { err = cl.Reserve(count); }
            if (err == cudaSuccess) { err = cr.Reserve(count); }
            if (err == cudaSuccess) { err = cudaMemcpyAsync(il.ptr, hil.data(), sizeof(int) * count, cudaMemcpyHostToDevice, stream); }
            if (err == cudaSuccess) { err = cudaMemcpyAsync(ir.ptr, hir.data(), sizeof(int) * count, cudaMemcpyHostToDevice, stream); }
            if (err == cudaSuccess) { err = cudaMemcpyAsync(cl.ptr, hcl.data(), sizeof(T) * count, cudaMemcpyHostToDevice, stream); }
            if (err == cudaSuccess) { err = cudaMemcpyAsync(cr.ptr, hcr.data(), sizeof(T) * count, cudaMemcpyHostToDevice, stream); }
            if (err == cudaSuccess) { err = cudaStreamSynchronize(stream); }  // host vectors die here
            return err;
        }
    };

    // ---- gather kernels: one warp per partition GP, order-preserving compaction -------------
    // 2-D lidar: samples are the hit rays of [index_left, index_right) in ray order.
    template<typename T>
    __global__ void
    LidarGatherKernel(
        const int num_parts,
        const int max_n,
        const int *__restrict__ il,
        const int *__restrict__ ir,
        const T *__restrict__ angles,
        const T *__restrict__ ranges,
        const uint8_t *__restrict__ hit,
        const uint8_t *__restrict__ con,
        const int discon_detection,
        const T range_var,
        const T discon_var,
        const int mapping,
        const T mapping_scale,
        int *__restrict__ n_train,
        T *__restrict__ x,
        T *__restrict__ y,
        T *__restrict__ var) {
        const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const int lane = threadIdx.x & 31;
        if (p >= num_parts) { return; }
        int cnt = 0;
        for (int j0 = il[p]; j0 < ir[p]; j0 += 32) {
            const int j = j0 + lane;
            const bool take = j < ir[p] && hit[j] != 0;
            const unsigned ballot = __ballot_sync(0xffffffffu, take);
            if (take) {
                const int slot = cnt + __popc(ballot & ((1u << lane) - 1u));
                if (slot < max_n) {
                    x[static_cast<long>(p) * max_n + slot] = angles[j];
                    y[static_cast<long>(p) * max_n + slot] = MappingMap<T>(mapping, mapping_scale, ranges[j]);
                    var[static_cast<long>(p) * max_n + slot] = (discon_detection && con[j] == 0) ? discon_var : range_var;
                }
            }
            cnt += __popc(ballot);
        }
        if (lane == 0) { n_train[p] = cnt < max_n ? cnt : max_n; }
    }

    // 3-D range sensor: GP (i, j) = row partition i x col partition j, samples gathered
    // col-outer / row-inner (src/range_sensor_gp_3d.cpp:348-356); arrays are rows x cols col-major.
    template<typename T>
    __global__ void
    Range3dGatherKernel(
        const int num_row_parts,
        const int num_col_parts,
        const int max_n,
        const int rows,
        const int *__restrict__ ril,
        const int *__restrict__ rir,
        const int *__restrict__ cil,
        const int *__restrict__ cir,
        const T *__restrict__ frame_coords,
        const T *__restrict__ ranges,
        const uint8_t *__restrict__ hit,
        const T range_var,
        const int mapping,
        const T mapping_scale,
        int *__restrict__ n_train,
        T *__restrict__ x,
        T *__restrict__ y,
        T *__restrict__ var) {
        const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const int lane = threadIdx.x & 31;
        if (g >= num_row_parts * num_col_parts) { return; }
        const int pi = g % num_row_parts, pj = g / num_row_parts;
        const int r0 = ril[pi], r1 = rir[pi], c0 = cil[pj], c1 = cir[pj];
        const int h = r1 - r0;
        const int total = h * (c1 - c0);
        int cnt = 0;
        for (int e0 = 0; e0 < total; e0 += 32) {
            const int e = e0 + lane;
            const int r = r0 + e % h, c = c0 + e / h;
            const long pix = r + static_cast<long>(c) * rows;
            const bool take = e < total && hit[pix] != 0;
            const unsigned ballot = __ballot_sync(0xffffffffu, take);
            if (take) {
                const int slot = cnt + __popc(ballot & ((1u << lane) - 1u));
                if (slot < max_n) {
                    const long dst = static_cast<long>(g) * max_n + slot;
                    x[2 * dst] = frame_coords[2 * pix];
                    x[2 * dst + 1] = frame_coords[2 * pix + 1];
                    y[dst] = MappingMap<T>(mapping, mapping_scale, ranges[pix]);
                    var[dst] = range_var;
                }
            }
            cnt += __popc(ballot);
        }
        if (lane == 0) { n_train[g] = cnt < max_n ? cnt : max_n; }
    }

    // ---- query -> GP assignment ----------------------------------------------------------------
    template<typename T>
    __device__ __forceinline__ int
    SearchPartitionDev(const int count, const T *__restrict__ cl, const T *__restrict__ cr, const T coord, const bool right_closed) {
        if (!isfinite(coord)) { return -1; }
        for (int i = 0; i < count; ++i) {  // first match, as the reference's linear scan
            const bool in = right_closed ? (coord >= cl[i] && coord <= cr[i]) : (coord >= cl[i] && coord < cr[i]);
            if (in) { return i; }
        }
        return -1;
    }

    // mode 0: angles (world or local); mode 1: 2-D positions in the sensor frame (ComputeOcc)
    template<typename T>
    __global__ void
    LidarAssignKernel(
        const long num_q,
        const T *__restrict__ q_in,
        const int mode,
        const int angles_are_local,
        const T r00,
        const T r10,
        const T r01,
        const T r11,
        const int num_parts,
        const T *__restrict__ cl,
        const T *__restrict__ cr,
        const int *__restrict__ info,
        T *__restrict__ q_local,
        T *__restrict__ q_dist,
        int *__restrict__ q_gp,
        int *__restrict__ counts) {
        const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
        if (i >= num_q) { return; }
        T a;
        if (mode == 0) {
            a = q_in[i];
            if (!angles_are_local) {  // DirWorldToFrame: R^T (cos a, sin a), then atan2 — src/lidar_gp_2d.cpp:69-75
                const T dx = cos(a), dy = sin(a);
                a = atan2(r01 * dx + r11 * dy, r00 * dx + r10 * dy);
            }
        } else {
            const T px = q_in[2 * i], py = q_in[2 * i + 1];
            q_dist[i] = sqrt(px * px + py * py);
            a = atan2(py, px);
        }
        q_local[i] = a;
        int g = SearchPartitionDev<T>(num_parts, cl, cr, a, true);
        if (g >= 0 && info[g] != 0) { g = -1; }  // partition GP not trained
        q_gp[i] = g;
        if (g >= 0) { atomicAdd(&counts[g], 1); }
    }

    template<typename T>
    __global__ void
    Range3dAssignKernel(
        const long num_q,
        const T *__restrict__ coords,
        const uint8_t *__restrict__ coords_ok,
        const int num_row_parts,
        const T *__restrict__ rcl,
        const T *__restrict__ rcr,
        const int num_col_parts,
        const T *__restrict__ ccl,
        const T *__restrict__ ccr,
        const int *__restrict__ info,
        int *__restrict__ q_gp,
        int *__restrict__ counts) {
        const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
        if (i >= num_q) { return; }
        int g = -1;
        if (coords_ok == nullptr || coords_ok[i] != 0) {
            const int pr = SearchPartitionDev<T>(num_row_parts, rcl, rcr, coords[2 * i], false);       // [l, r)  :376
            const int pc = pr >= 0 ? SearchPartitionDev<T>(num_col_parts, ccl, ccr, coords[2 * i + 1], true) : -1;  // [l, r]  :388
            if (pr >= 0 && pc >= 0) {
                g = pr + pc * num_row_parts;
                if (info[g] != 0) { g = -1; }
            }
        }
        q_gp[i] = g;
        if (g >= 0) { atomicAdd(&counts[g], 1); }
    }

    // exclusive scan of the per-GP counts (single CTA; num_gps is at most a few thousand)
    __global__ void
    ScanCountsKernel(const int num_gps, const int *__restrict__ counts, long *__restrict__ offsets, int *__restrict__ cursor) {
        __shared__ long s_part[1024];
        const int tid = threadIdx.x;
        const int per = (num_gps + blockDim.x - 1) / blockDim.x;
        const int begin = tid * per;
        const int end = begin + per < num_gps ? begin + per : num_gps;
        long sum = 0;
        for (int i = begin; i < end; ++i) { sum += counts[i]; }
        s_part[tid] = sum;
        __syncthreads();
        if (tid == 0) {
            long run = 0;
            for (unsigned t = 0; t < blockDim.x; ++t) {
                const long v = s_part[t];
                s_part[t] = run;
                run += v;
            }
            offsets[num_gps] = run;
        }
        __syncthreads();
        long run = s_part[tid];
        for (int i = begin; i < end; ++i) {
            offsets[i] = run;
            run += counts[i];
            cursor[i] = 0;
        }
    }

    template<typename T, int XDIM>
    __global__ void
    ScatterQueriesKernel(const long num_q, const T *__restrict__ q_local, const int *__restrict__ q_gp, const long *__restrict__ offsets, int *__restrict__ cursor, T *__restrict__ sorted_x,
                         int *__restrict__ out_index) {
        const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
        if (i >= num_q) { return; }
        const int g = q_gp[i];
        if (g < 0) { return; }
        const long pos = offsets[g] + atomicAdd(&cursor[g], 1);
#pragma unroll
        for (int d = 0; d < XDIM; ++d) { sorted_x[pos * XDIM + d] = q_local[i * XDIM + d]; }
        out_index[pos] = static_cast<int>(i);
    }

    // ComputeOcc epilogue — src/lidar_gp_2d.cpp:447-457
    template<typename T>
    __global__ void
    OccEpilogueKernel(const long num, const T *__restrict__ dist, const uint8_t *__restrict__ valid, const T *__restrict__ variance, const T max_valid_range_var, const T temperature,
                      const int mapping, const T mapping_scale, T *__restrict__ range_pred, T *__restrict__ occ, uint8_t *__restrict__ ok, const T *__restrict__ prefill) {
        const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
        if (i >= num) { return; }
        bool good = valid[i] != 0 && !(variance[i] > max_valid_range_var);
        if (good) {
            const T f = range_pred[i];
            const T a = dist[i] * temperature;
            occ[i] = T(2) / (T(1) + exp(a * (f - MappingMap<T>(mapping, mapping_scale, dist[i])))) - T(1);
            range_pred[i] = MappingInv<T>(mapping, mapping_scale, f);
        } else if (prefill != nullptr) {
            range_pred[i] = prefill[i];  // rejected by the variance gate: the reference returns before it assigns range_pred
        }
        ok[i] = good ? 1 : 0;
    }

    // ---- shared query pipeline state ----------------------------------------------------------
    template<typename T>
    struct QueryWorkspace {
        DeviceBuffer<T> q_in, q_local, q_dist, sorted_x, mean, variance;
        DeviceBuffer<int> q_gp, counts, cursor, out_index;
        DeviceBuffer<long> offsets;
        DeviceBuffer<uint8_t> valid, coords_ok, ok;

        cudaError_t
        Reserve(long num_q, long num_gps, int in_dim, int x_dim) {
            cudaError_t err = q_in.Reserve(num_q * in_dim);
            if (err == cudaSuccess) { err = q_local.Reserve(num_q * x_dim); }
            if (err == cudaSuccess) { err = q_dist.Reserve(num_q); }
            if (err == cudaSuccess) { err = sorted_x.Reserve(num_q * x_dim); }
            if (err == cudaSuccess) { err = mean.Reserve(num_q); }
            if (err == cudaSuccess) { err = variance.Reserve(num_q); }
            if (err == cudaSuccess) { err = q_gp.Reserve(num_q); }
            if (err == cudaSuccess) { err = out_index.Reserve(num_q); }
            if (err == cudaSuccess) { err = counts.Reserve(num_gps); }
            if (err == cudaSuccess) { err = cursor.Reserve(num_gps); }
            if (err == cudaSuccess) { err = offsets.Reserve(num_gps + 1); }
            if (err == cudaSuccess) { err = valid.Reserve(num_q); }
            if (err == cudaSuccess) { err = coords_ok.Reserve(num_q); }
            if (err == cudaSuccess) { err = ok.Reserve(num_q); }
            return err;
        }
    };

    template<typename T>
    struct Lidar2d {
        Context *ctx = nullptr;
        erl_gp_lidar2d_setting setting{};
        long num_rays = 0;
        PartitionTable<T> parts;
        DevicePartitions<T> d_parts;
        std::vector<T> h_angles;        // partition_on_hit_rays: the table is rebuilt from every frame's hit mask
        std::vector<uint8_t> h_hit;
        long capacity = 0;              // partition GPs in the batch (= parts.Size() unless partition_on_hit_rays)
        DeviceBuffer<T> d_angles, d_ranges;
        DeviceBuffer<uint8_t> d_hit, d_con;
        Batch<T> *batch = nullptr;
        T rotation[4] = {1, 0, 0, 1};
        bool trained = false;
        QueryWorkspace<T> ws;

        ~Lidar2d() { delete batch; }
    };

    template<typename T>
    struct Range3d {
        Context *ctx = nullptr;
        erl_gp_range3d_setting setting{};
        long rows = 0, cols = 0;
        PartitionTable<T> row_parts, col_parts;
        DevicePartitions<T> d_row_parts, d_col_parts;
        DeviceBuffer<T> d_frame_coords, d_ranges;
        DeviceBuffer<uint8_t> d_hit;
        Batch<T> *batch = nullptr;
        bool trained = false;
        QueryWorkspace<T> ws;

        ~Range3d() { delete batch; }
    };

    template<typename T>
    static int
    Lidar2dCreate(erl_gp_context *c, const erl_gp_lidar2d_setting *s, const T *angles, long num_rays, Lidar2d<T> **out) {
        Context *ctx = Ctx(c);
        if (ctx == nullptr || s == nullptr || angles == nullptr || out == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        *out = nullptr;
        if (s->group_size <= s->overlap_size || s->overlap_size < 0 || s->margin < 0 || s->margin >= num_rays) {
            return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "lidar2d: group_size must exceed overlap_size and 0 <= margin < num_rays");
        }
        if (num_rays <= s->overlap_size) {  // "no enough samples to perform partition", src/lidar_gp_2d.cpp:177-180
            return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "lidar2d: num_rays=%ld <= overlap_size=%ld", num_rays, s->overlap_size);
        }
        auto *gp = new (std::nothrow) Lidar2d<T>();
        if (gp == nullptr) { return ERL_GP_STATUS_ALLOC_FAILED; }
        gp->ctx = ctx;
        gp->setting = *s;
        gp->num_rays = num_rays;
        if (s->partition_on_hit_rays) {
            // the constructor leaves the table empty (src/lidar_gp_2d.cpp:182); Train() builds it.  At most
            // max(1, num_rays / step) + 1 partitions (:311 with n <= num_rays)
            gp->h_angles.assign(angles, angles + num_rays);
            gp->capacity = std::max(1l, num_rays / (s->group_size - s->overlap_size)) + 1;
        } else {
            gp->parts = MakePartitionTable<T>(angles, 1, num_rays, s->group_size, s->overlap_size, s->margin, s->symmetric_partitions != 0);
            if (!PartitionsFit(gp->parts, num_rays, s->group_size)) {
                delete gp;
                return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "lidar2d: partition table does not fit group_size (margin / overlap out of range)");
            }
            gp->capacity = gp->parts.Size();
        }
        int rc = BatchCreate<T>(c, gp->capacity, s->group_size, 1, s->kernel, static_cast<T>(s->kernel_scale), &gp->batch);
        if (rc != ERL_GP_STATUS_OK) {
            delete gp;
            return rc;
        }
        cudaError_t err = gp->d_parts.Upload(gp->parts, ctx->stream, gp->capacity);
        if (err == cudaSuccess) { err = gp->d_angles.Reserve(num_rays); }
        if (err == cudaSuccess) { err = gp->d_ranges.Reserve(num_rays); }
        if (err == cudaSuccess) { err = gp->d_hit.Reserve(num_rays); }
        if (err == cudaSuccess) { err = gp->d_con.Reserve(num_rays); }
        if (err == cudaSuccess) { err = cudaMemcpyAsync(gp->d_angles.ptr, angles, sizeof(T) * num_rays, cudaMemcpyHostToDevice, ctx->stream); }
        if (err == cudaSuccess) { err = cudaStreamSynchronize(ctx->stream); }
        if (err != cudaSuccess) {
            delete gp;
            return SetError(ctx, ERL_GP_STATUS_CUDA_ERROR, "lidar2d: %s", cudaGetErrorString(err));
        }
        *out = gp;
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    static int
    Lidar2dTrain(Lidar2d<T> *gp, const T *rotation, const T *ranges, const uint8_t *hit, const uint8_t *con) {
        if (gp == nullptr || ranges == nullptr || hit == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        const auto &s = gp->setting;
        gp->trained = false;
        if (rotation != nullptr) { std::memcpy(gp->rotation, rotation, sizeof(gp->rotation)); }
        const long n = gp->num_rays;
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->d_ranges.ptr, ranges, sizeof(T) * n, cudaMemcpyDefault, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->d_hit.ptr, hit, n, cudaMemcpyDefault, ctx->stream));
        if (con != nullptr) {
            ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->d_con.ptr, con, n, cudaMemcpyDefault, ctx->stream));
        } else {
            ERL_GP_CUDA_OK(ctx, cudaMemsetAsync(gp->d_con.ptr, 1, n, ctx->stream));
        }
        if (s.partition_on_hit_rays) {  // src/lidar_gp_2d.cpp:364
            gp->h_hit.resize(static_cast<std::size_t>(n));
            ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->h_hit.data(), hit, n, cudaMemcpyDefault, ctx->stream));  // host or device pointer
            ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
            PartitionTable<T> t = MakeHitRayPartitionTable<T>(gp->h_angles.data(), n, gp->h_hit.data(), s.group_size, s.overlap_size);
            if (t.Size() > 0) {  // "No hit rays are stored": the previous table stays (:306-309)
                gp->parts = std::move(t);
                ERL_GP_CUDA_OK(ctx, gp->d_parts.Upload(gp->parts, ctx->stream, gp->capacity));
            }
        }
        Batch<T> *b = gp->batch;
        const int num_parts = static_cast<int>(gp->capacity);  // padding partitions are empty: their GPs come out untrained
        const int warps_per_block = 4;
        LidarGatherKernel<T><<<static_cast<unsigned>(CeilDiv(num_parts, warps_per_block)), warps_per_block * 32, 0, ctx->stream>>>(
            num_parts, static_cast<int>(b->max_n), gp->d_parts.il.ptr, gp->d_parts.ir.ptr, gp->d_angles.ptr, gp->d_ranges.ptr, gp->d_hit.ptr, gp->d_con.ptr, s.discontinuity_detection,
            static_cast<T>(s.sensor_range_var), static_cast<T>(s.discontinuity_var), s.mapping, static_cast<T>(s.mapping_scale), b->n_train.ptr, b->x.ptr, b->y.ptr, b->var.ptr);
        ctx->launches += 1;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        const int rc = BatchTrainDev<T>(b, 0, 1);  // train iff cnt > 0, src/lidar_gp_2d.cpp:391
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        gp->trained = true;
        return ERL_GP_STATUS_OK;
    }

    // assignment + counting sort + batched predict.  mode 0: angles, mode 1: positions (ComputeOcc).
    template<typename T>
    static int
    Lidar2dQuery(Lidar2d<T> *gp, const T *q_host, long num_q, int mode, int angles_are_local, int mapping, const T *mean_seed, const T *var_seed) {
        Context *ctx = gp->ctx;
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        Batch<T> *b = gp->batch;
        QueryWorkspace<T> &ws = gp->ws;
        const int in_dim = mode == 0 ? 1 : 2;
        ERL_GP_CUDA_OK(ctx, ws.Reserve(num_q, b->num_gps, in_dim, 1));
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(ws.q_in.ptr, q_host, sizeof(T) * num_q * in_dim, cudaMemcpyDefault, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaMemsetAsync(ws.counts.ptr, 0, sizeof(int) * b->num_gps, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaMemsetAsync(ws.valid.ptr, 0, num_q, ctx->stream));
        if (mean_seed != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(ws.mean.ptr, mean_seed, sizeof(T) * num_q, cudaMemcpyDefault, ctx->stream)); }
        if (var_seed != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(ws.variance.ptr, var_seed, sizeof(T) * num_q, cudaMemcpyDefault, ctx->stream)); }
        const unsigned blocks = static_cast<unsigned>(CeilDiv(num_q, 256));
        const T *r = gp->rotation;  // col-major: r[0]=R00 r[1]=R10 r[2]=R01 r[3]=R11
        LidarAssignKernel<T><<<blocks, 256, 0, ctx->stream>>>(num_q, ws.q_in.ptr, mode, angles_are_local, r[0], r[1], r[2], r[3], gp->d_parts.count, gp->d_parts.cl.ptr, gp->d_parts.cr.ptr,
                                                              b->info.ptr, ws.q_local.ptr, ws.q_dist.ptr, ws.q_gp.ptr, ws.counts.ptr);
        ScanCountsKernel<<<1, 1024, 0, ctx->stream>>>(static_cast<int>(b->num_gps), ws.counts.ptr, ws.offsets.ptr, ws.cursor.ptr);
        ScatterQueriesKernel<T, 1><<<blocks, 256, 0, ctx->stream>>>(num_q, ws.q_local.ptr, ws.q_gp.ptr, ws.offsets.ptr, ws.cursor.ptr, ws.sorted_x.ptr, ws.out_index.ptr);
        ctx->launches += 3;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        return BatchPredictDev<T>(b, ws.offsets.ptr, ws.sorted_x.ptr, ws.out_index.ptr, num_q, mapping, static_cast<T>(gp->setting.mapping_scale), ws.mean.ptr, ws.variance.ptr, ws.valid.ptr);
    }

    template<typename T>
    static int
    Lidar2dTest(Lidar2d<T> *gp, const T *angles, long num_test, int angles_are_local, int un_map, T *mean, T *var, uint8_t *valid) {
        if (gp == nullptr || angles == nullptr || num_test < 0) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        if (!gp->trained) { return SetError(ctx, ERL_GP_STATUS_NOT_TRAINED, "lidar2d: Test() before Train()"); }
        if (num_test == 0) { return ERL_GP_STATUS_OK; }
        const int rc = Lidar2dQuery<T>(gp, angles, num_test, 0, angles_are_local, un_map ? gp->setting.mapping : ERL_GP_MAPPING_NONE, mean, var);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        QueryWorkspace<T> &ws = gp->ws;
        if (mean != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(mean, ws.mean.ptr, sizeof(T) * num_test, cudaMemcpyDefault, ctx->stream)); }
        if (var != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(var, ws.variance.ptr, sizeof(T) * num_test, cudaMemcpyDefault, ctx->stream)); }
        if (valid != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(valid, ws.valid.ptr, num_test, cudaMemcpyDefault, ctx->stream)); }
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    static int
    Lidar2dComputeOcc(Lidar2d<T> *gp, const T *pos, long num, T max_valid_range_var, T temperature, T *dist, T *range_pred, T *occ, uint8_t *ok) {
        if (gp == nullptr || pos == nullptr || num < 0 || ok == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        if (!gp->trained) { return SetError(ctx, ERL_GP_STATUS_NOT_TRAINED, "lidar2d: ComputeOcc() before Train()"); }
        if (num == 0) { return ERL_GP_STATUS_OK; }
        int rc = Lidar2dQuery<T>(gp, pos, num, 1, 1, ERL_GP_MAPPING_NONE, range_pred, nullptr);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        QueryWorkspace<T> &ws = gp->ws;
        // occ lives in sorted_x's slot-free twin: reuse q_local (dead after the scatter) as the occ buffer
        T *d_occ = ws.q_local.ptr;
        if (occ != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(d_occ, occ, sizeof(T) * num, cudaMemcpyDefault, ctx->stream)); }
        T *d_prefill = nullptr;
        if (range_pred != nullptr) {  // sorted_x is dead after the predict: keeps the caller's values for the gated-out positions
            d_prefill = ws.sorted_x.ptr;
            ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(d_prefill, range_pred, sizeof(T) * num, cudaMemcpyDefault, ctx->stream));
        }
        OccEpilogueKernel<T><<<static_cast<unsigned>(CeilDiv(num, 256)), 256, 0, ctx->stream>>>(num, ws.q_dist.ptr, ws.valid.ptr, ws.variance.ptr, max_valid_range_var, temperature, gp->setting.mapping,
                                                                                               static_cast<T>(gp->setting.mapping_scale), ws.mean.ptr, d_occ, ws.ok.ptr, d_prefill);
        ctx->launches += 1;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        if (dist != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(dist, ws.q_dist.ptr, sizeof(T) * num, cudaMemcpyDefault, ctx->stream)); }
        if (range_pred != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(range_pred, ws.mean.ptr, sizeof(T) * num, cudaMemcpyDefault, ctx->stream)); }
        if (occ != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(occ, d_occ, sizeof(T) * num, cudaMemcpyDefault, ctx->stream)); }
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(ok, ws.ok.ptr, num, cudaMemcpyDefault, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    static int
    Range3dCreate(erl_gp_context *c, const erl_gp_range3d_setting *s, const T *frame_coords, long rows, long cols, Range3d<T> **out) {
        Context *ctx = Ctx(c);
        if (ctx == nullptr || s == nullptr || frame_coords == nullptr || out == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        *out = nullptr;
        if (s->row_overlap_size % 2 != 0 || s->col_overlap_size % 2 != 0) {  // src/range_sensor_gp_3d.cpp:190-197
            return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "range3d: row_overlap_size / col_overlap_size must be even");
        }
        if (s->row_margin < 0 || s->row_margin >= rows || s->col_margin < 0 || s->col_margin >= cols || s->row_overlap_size < 0 || s->col_overlap_size < 0) {
            return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "range3d: margins must lie inside the frame and overlaps must be >= 0");
        }
        if (s->row_group_size <= s->row_overlap_size || s->col_group_size <= s->col_overlap_size || rows <= s->row_overlap_size || cols <= s->col_overlap_size) {
            return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "range3d: group sizes must exceed overlap sizes and fit the frame");
        }
        auto *gp = new (std::nothrow) Range3d<T>();
        if (gp == nullptr) { return ERL_GP_STATUS_ALLOC_FAILED; }
        gp->ctx = ctx;
        gp->setting = *s;
        gp->rows = rows;
        gp->cols = cols;
        // row coordinate = component 0 of frame_coords(r, 0); col coordinate = component 1 of frame_coords(0, c)
        gp->row_parts = MakePartitionTable<T>(frame_coords, 2, rows, s->row_group_size, s->row_overlap_size, s->row_margin, true);
        gp->col_parts = MakePartitionTable<T>(frame_coords + 1, 2 * rows, cols, s->col_group_size, s->col_overlap_size, s->col_margin, true);
        if (!PartitionsFit(gp->row_parts, rows, s->row_group_size) || !PartitionsFit(gp->col_parts, cols, s->col_group_size)) {
            delete gp;
            return SetError(ctx, ERL_GP_STATUS_INVALID_ARGUMENT, "range3d: partition table does not fit the group sizes");
        }
        const long num_gps = gp->row_parts.Size() * gp->col_parts.Size();
        int rc = BatchCreate<T>(c, num_gps, s->row_group_size * s->col_group_size, 2, s->kernel, static_cast<T>(s->kernel_scale), &gp->batch);
        if (rc != ERL_GP_STATUS_OK) {
            delete gp;
            return rc;
        }
        const long pixels = rows * cols;
        cudaError_t err = gp->d_row_parts.Upload(gp->row_parts, ctx->stream);
        if (err == cudaSuccess) { err = gp->d_col_parts.Upload(gp->col_parts, ctx->stream); }
        if (err == cudaSuccess) { err = gp->d_frame_coords.Reserve(2 * pixels); }
        if (err == cudaSuccess) { err = gp->d_ranges.Reserve(pixels); }
        if (err == cudaSuccess) { err = gp->d_hit.Reserve(pixels); }
        if (err == cudaSuccess) { err = cudaMemcpyAsync(gp->d_frame_coords.ptr, frame_coords, sizeof(T) * 2 * pixels, cudaMemcpyHostToDevice, ctx->stream); }
        if (err == cudaSuccess) { err = cudaStreamSynchronize(ctx->stream); }
        if (err != cudaSuccess) {
            delete gp;
            return SetError(ctx, ERL_GP_STATUS_CUDA_ERROR, "range3d: %s", cudaGetErrorString(err));
        }
        *out = gp;
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    static int
    Range3dTrain(Range3d<T> *gp, const T *ranges, const uint8_t *hit) {
        if (gp == nullptr || ranges == nullptr || hit == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        const auto &s = gp->setting;
        gp->trained = false;
        const long pixels = gp->rows * gp->cols;
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->d_ranges.ptr, ranges, sizeof(T) * pixels, cudaMemcpyDefault, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(gp->d_hit.ptr, hit, pixels, cudaMemcpyDefault, ctx->stream));
        Batch<T> *b = gp->batch;
        const int warps_per_block = 4;
        Range3dGatherKernel<T><<<static_cast<unsigned>(CeilDiv(b->num_gps, warps_per_block)), warps_per_block * 32, 0, ctx->stream>>>(
            static_cast<int>(gp->row_parts.Size()), static_cast<int>(gp->col_parts.Size()), static_cast<int>(b->max_n), static_cast<int>(gp->rows), gp->d_row_parts.il.ptr,
            gp->d_row_parts.ir.ptr, gp->d_col_parts.il.ptr, gp->d_col_parts.ir.ptr, gp->d_frame_coords.ptr, gp->d_ranges.ptr, gp->d_hit.ptr, static_cast<T>(s.sensor_range_var), s.mapping,
            static_cast<T>(s.mapping_scale), b->n_train.ptr, b->x.ptr, b->y.ptr, b->var.ptr);
        ctx->launches += 1;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        const int rc = BatchTrainDev<T>(b, s.min_num_samples_per_group, 1);  // train iff cnt > min, src/range_sensor_gp_3d.cpp:358
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        gp->trained = true;
        return ERL_GP_STATUS_OK;
    }

    // Batched RangeSensorGaussianProcess3D::ComputeOcc (src/range_sensor_gp_3d.cpp:409-439).  ComputeFrameCoords /
    // CoordsIsInFrame belong to erl_geometry and stay on the caller's side: coords (2 x num), coords_ok (their combined bool)
    // and dist (the distance ComputeFrameCoords returns) are the inputs.
    template<typename T>
    static int
    Range3dTest(Range3d<T> *gp, const T *coords, const uint8_t *coords_ok, long num_test, int un_map, T *mean, T *var, uint8_t *valid);

    template<typename T>
    static int
    Range3dComputeOcc(Range3d<T> *gp, const T *coords, const uint8_t *coords_ok, const T *dist, long num, T max_valid_range_var, T temperature, T *range_pred, T *occ, uint8_t *ok) {
        if (gp == nullptr || coords == nullptr || dist == nullptr || num < 0 || ok == nullptr || range_pred == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        if (!gp->trained) { return SetError(ctx, ERL_GP_STATUS_NOT_TRAINED, "range3d: ComputeOcc() before Train()"); }
        if (num == 0) { return ERL_GP_STATUS_OK; }
        // the caller's range_pred values come back for every rejected position: keep them on the device first
        QueryWorkspace<T> &ws = gp->ws;
        ERL_GP_CUDA_OK(ctx, ws.Reserve(num, gp->batch->num_gps, 2, 2));
        ERL_GP_CUDA_OK(ctx, ws.q_local.Reserve(2 * num));
        T *d_prefill = ws.q_local.ptr + num;  // q_local is unused by the 3-D query: [0, num) = occ, [num, 2 num) = prefill
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(d_prefill, range_pred, sizeof(T) * num, cudaMemcpyDefault, ctx->stream));
        // mapped mean + variance of every position (results stay in the workspace; the host copy of mean is re-written below)
        std::vector<T> var_host(static_cast<size_t>(num));
        int rc = Range3dTest<T>(gp, coords, coords_ok, num, 0, range_pred, var_host.data(), nullptr);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(ws.q_dist.ptr, dist, sizeof(T) * num, cudaMemcpyDefault, ctx->stream));
        T *d_occ = ws.q_local.ptr;
        if (occ != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(d_occ, occ, sizeof(T) * num, cudaMemcpyDefault, ctx->stream)); }
        OccEpilogueKernel<T><<<static_cast<unsigned>(CeilDiv(num, 256)), 256, 0, ctx->stream>>>(num, ws.q_dist.ptr, ws.valid.ptr, ws.variance.ptr, max_valid_range_var, temperature, gp->setting.mapping,
                                                                                               static_cast<T>(gp->setting.mapping_scale), ws.mean.ptr, d_occ, ws.ok.ptr, d_prefill);
        ctx->launches += 1;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(range_pred, ws.mean.ptr, sizeof(T) * num, cudaMemcpyDefault, ctx->stream));
        if (occ != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(occ, d_occ, sizeof(T) * num, cudaMemcpyDefault, ctx->stream)); }
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(ok, ws.ok.ptr, num, cudaMemcpyDefault, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    static int
    Range3dTest(Range3d<T> *gp, const T *coords, const uint8_t *coords_ok, long num_test, int un_map, T *mean, T *var, uint8_t *valid) {
        if (gp == nullptr || coords == nullptr || num_test < 0) { return ERL_GP_STATUS_INVALID_ARGUMENT; }
        Context *ctx = gp->ctx;
        ERL_GP_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        if (!gp->trained) { return SetError(ctx, ERL_GP_STATUS_NOT_TRAINED, "range3d: Test() before Train()"); }
        if (num_test == 0) { return ERL_GP_STATUS_OK; }
        Batch<T> *b = gp->batch;
        QueryWorkspace<T> &ws = gp->ws;
        ERL_GP_CUDA_OK(ctx, ws.Reserve(num_test, b->num_gps, 2, 2));
        ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(ws.q_in.ptr, coords, sizeof(T) * 2 * num_test, cudaMemcpyDefault, ctx->stream));
        if (coords_ok != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(ws.coords_ok.ptr, coords_ok, num_test, cudaMemcpyDefault, ctx->stream)); }
        ERL_GP_CUDA_OK(ctx, cudaMemsetAsync(ws.counts.ptr, 0, sizeof(int) * b->num_gps, ctx->stream));
        ERL_GP_CUDA_OK(ctx, cudaMemsetAsync(ws.valid.ptr, 0, num_test, ctx->stream));
        if (mean != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(ws.mean.ptr, mean, sizeof(T) * num_test, cudaMemcpyDefault, ctx->stream)); }
        if (var != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(ws.variance.ptr, var, sizeof(T) * num_test, cudaMemcpyDefault, ctx->stream)); }
        const unsigned blocks = static_cast<unsigned>(CeilDiv(num_test, 256));
        Range3dAssignKernel<T><<<blocks, 256, 0, ctx->stream>>>(num_test, ws.q_in.ptr, coords_ok != nullptr ? ws.coords_ok.ptr : nullptr, gp->d_row_parts.count, gp->d_row_parts.cl.ptr,
                                                                gp->d_row_parts.cr.ptr, gp->d_col_parts.count, gp->d_col_parts.cl.ptr, gp->d_col_parts.cr.ptr, b->info.ptr, ws.q_gp.ptr,
                                                                ws.counts.ptr);
        ScanCountsKernel<<<1, 1024, 0, ctx->stream>>>(static_cast<int>(b->num_gps), ws.counts.ptr, ws.offsets.ptr, ws.cursor.ptr);
        ScatterQueriesKernel<T, 2><<<blocks, 256, 0, ctx->stream>>>(num_test, ws.q_in.ptr, ws.q_gp.ptr, ws.offsets.ptr, ws.cursor.ptr, ws.sorted_x.ptr, ws.out_index.ptr);
        ctx->launches += 3;
        ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        const int rc = BatchPredictDev<T>(b, ws.offsets.ptr, ws.sorted_x.ptr, ws.out_index.ptr, num_test, un_map ? gp->setting.mapping : ERL_GP_MAPPING_NONE,
                                          static_cast<T>(gp->setting.mapping_scale), mean != nullptr ? ws.mean.ptr : nullptr, var != nullptr ? ws.variance.ptr : nullptr, ws.valid.ptr);
        if (rc != ERL_GP_STATUS_OK) { return rc; }
        if (mean != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(mean, ws.mean.ptr, sizeof(T) * num_test, cudaMemcpyDefault, ctx->stream)); }
        if (var != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(var, ws.variance.ptr, sizeof(T) * num_test, cudaMemcpyDefault, ctx->stream)); }
        if (valid != nullptr) { ERL_GP_CUDA_OK(ctx, cudaMemcpyAsync(valid, ws.valid.ptr, num_test, cudaMemcpyDefault, ctx->stream)); }
        ERL_GP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return ERL_GP_STATUS_OK;
    }

    template<typename T>
    static int
    CopyPartitions(const PartitionTable<T> &t, long *il, long *ir, T *cl, T *cr) {
        for (long i = 0; i < t.Size(); ++i) {
            if (il != nullptr) { il[i] = t.index_left[i]; }
            if (ir != nullptr) { ir[i] = t.index_right[i]; }
            if (cl != nullptr) { cl[i] = t.coord_left[i]; }
            if (cr != nullptr) { cr[i] = t.coord_right[i]; }
        }
        return ERL_GP_STATUS_OK;
    }

}  // namespace erl_gp

using namespace erl_gp;

struct erl_gp_lidar2d_f32 : Lidar2d<float> {};
struct erl_gp_lidar2d_f64 : Lidar2d<double> {};
struct erl_gp_range3d_f32 : Range3d<float> {};
struct erl_gp_range3d_f64 : Range3d<double> {};

extern "C" {

#define ERL_GP_DEFINE_SENSOR(T, SFX)                                                                                                                                            \
    int erl_gp_lidar2d_create_##SFX(erl_gp_context *ctx, const erl_gp_lidar2d_setting *setting, const T *angles, long num_rays, erl_gp_lidar2d_##SFX **gp) {                    \
        return Lidar2dCreate<T>(ctx, setting, angles, num_rays, reinterpret_cast<Lidar2d<T> **>(gp));                                                                           \
    }                                                                                                                                                                           \
    int erl_gp_lidar2d_destroy_##SFX(erl_gp_lidar2d_##SFX *gp) {                                                                                                                \
        if (gp != nullptr) {                                                                                                                                                    \
            cudaSetDevice(gp->ctx->device);                                                                                                                                     \
            cudaStreamSynchronize(gp->ctx->stream);                                                                                                                             \
            delete static_cast<Lidar2d<T> *>(gp);                                                                                                                               \
        }                                                                                                                                                                       \
        return ERL_GP_STATUS_OK;                                                                                                                                                \
    }                                                                                                                                                                           \
    int erl_gp_lidar2d_num_partitions_##SFX(erl_gp_lidar2d_##SFX *gp, long *num) {                                                                                              \
        if (gp == nullptr || num == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }                                                                                         \
        *num = gp->parts.Size();                                                                                                                                                \
        return ERL_GP_STATUS_OK;                                                                                                                                                \
    }                                                                                                                                                                           \
    int erl_gp_lidar2d_partitions_##SFX(erl_gp_lidar2d_##SFX *gp, long *index_left, long *index_right, T *coord_left, T *coord_right) {                                         \
        if (gp == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }                                                                                                           \
        return CopyPartitions<T>(gp->parts, index_left, index_right, coord_left, coord_right);                                                                                  \
    }                                                                                                                                                                           \
    int erl_gp_lidar2d_train_##SFX(erl_gp_lidar2d_##SFX *gp, const T *rotation, const T *ranges, const uint8_t *mask_hit, const uint8_t *mask_continuous) {                     \
        return Lidar2dTrain<T>(gp, rotation, ranges, mask_hit, mask_continuous);                                                                                                \
    }                                                                                                                                                                           \
    int erl_gp_lidar2d_test_##SFX(erl_gp_lidar2d_##SFX *gp, const T *angles, long num_test, int angles_are_local, int un_map, T *mean, T *var, uint8_t *valid) {                \
        return Lidar2dTest<T>(gp, angles, num_test, angles_are_local, un_map, mean, var, valid);                                                                                \
    }                                                                                                                                                                           \
    int erl_gp_lidar2d_get_gp_##SFX(erl_gp_lidar2d_##SFX *gp, long p, int *info, long *n, T *l, long ld_l, T *alpha) {                                                          \
        if (gp == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }                                                                                                           \
        return BatchGetGp<T>(gp->batch, p, info, n, l, ld_l, alpha);                                                                                                            \
    }                                                                                                                                                                           \
    int erl_gp_lidar2d_compute_occ_##SFX(erl_gp_lidar2d_##SFX *gp, const T *pos_local, long num, T max_valid_range_var, T occ_test_temperature, T *dist, T *range_pred, T *occ, \
                                         uint8_t *ok) {                                                                                                                         \
        return Lidar2dComputeOcc<T>(gp, pos_local, num, max_valid_range_var, occ_test_temperature, dist, range_pred, occ, ok);                                                  \
    }                                                                                                                                                                           \
    int erl_gp_range3d_create_##SFX(erl_gp_context *ctx, const erl_gp_range3d_setting *setting, const T *frame_coords, long rows, long cols, erl_gp_range3d_##SFX **gp) {       \
        return Range3dCreate<T>(ctx, setting, frame_coords, rows, cols, reinterpret_cast<Range3d<T> **>(gp));                                                                   \
    }                                                                                                                                                                           \
    int erl_gp_range3d_destroy_##SFX(erl_gp_range3d_##SFX *gp) {                                                                                                                \
        if (gp != nullptr) {                                                                                                                                                    \
            cudaSetDevice(gp->ctx->device);                                                                                                                                     \
            cudaStreamSynchronize(gp->ctx->stream);                                                                                                                             \
            delete static_cast<Range3d<T> *>(gp);                                                                                                                               \
        }                                                                                                                                                                       \
        return ERL_GP_STATUS_OK;                                                                                                                                                \
    }                                                                                                                                                                           \
    int erl_gp_range3d_grid_##SFX(erl_gp_range3d_##SFX *gp, long *num_row_partitions, long *num_col_partitions) {                                                               \
        if (gp == nullptr || num_row_partitions == nullptr || num_col_partitions == nullptr) { return ERL_GP_STATUS_INVALID_ARGUMENT; }                                         \
        *num_row_partitions = gp->row_parts.Size();                                                                                                                             \
        *num_col_partitions = gp->col_parts.Size();                                                                                                                             \
        return ERL_GP_STATUS_OK;                                                                                                                                                \
    }                                                                                                                                                                           \
    int erl_gp_range3d_partitions_##SFX(erl_gp_range3d_##SFX *gp, int axis, long *index_left, long *index_right, T *coord_left, T *coord_right) {                               \
        if (gp == nullptr || (axis != 0 && axis != 1)) { return ERL_GP_STATUS_INVALID_ARGUMENT; }                                                                               \
        return CopyPartitions<T>(axis == 0 ? gp->row_parts : gp->col_parts, index_left, index_right, coord_left, coord_right);                                                  \
    }                                                                                                                                                                           \
    int erl_gp_range3d_train_##SFX(erl_gp_range3d_##SFX *gp, const T *ranges, const uint8_t *mask_hit) { return Range3dTrain<T>(gp, ranges, mask_hit); }                        \
    int erl_gp_range3d_test_##SFX(erl_gp_range3d_##SFX *gp, const T *coords, const uint8_t *coords_ok, long num_test, int un_map, T *mean, T *var, uint8_t *valid) {            \
        return Range3dTest<T>(gp, coords, coords_ok, num_test, un_map, mean, var, valid);                                                                                       \
    }                                                                                                                                                                           \
    int erl_gp_range3d_compute_occ_##SFX(erl_gp_range3d_##SFX *gp, const T *coords, const uint8_t *coords_ok, const T *dist, long num, T max_valid_range_var,                   \
                                         T occ_test_temperature, T *range_pred, T *occ, uint8_t *ok) {                                                                          \
        return Range3dComputeOcc<T>(gp, coords, coords_ok, dist, num, max_valid_range_var, occ_test_temperature, range_pred, occ, ok);                                          \
    }                                                                                                                                                                           \
    int erl_gp_range3d_get_gp_##SFX(erl_gp_range3d_##SFX *gp, long row_part, long col_part, int *info, long *n, T *l, long ld_l, T *alpha) {                                    \
        if (gp == nullptr || row_part < 0 || col_part < 0 || row_part >= gp->row_parts.Size() || col_part >= gp->col_parts.Size()) { return ERL_GP_STATUS_INVALID_ARGUMENT; }   \
        return BatchGetGp<T>(gp->batch, row_part + col_part * gp->row_parts.Size(), info, n, l, ld_l, alpha);                                                                   \
    }

ERL_GP_DEFINE_SENSOR(float, f32)
ERL_GP_DEFINE_SENSOR(double, f64)

}  // extern "C"
