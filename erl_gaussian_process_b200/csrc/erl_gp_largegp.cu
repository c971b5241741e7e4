// Partition GPs beyond the shared-memory kernels' capacity (n > 256 in float, n > 192 in double).
//
// The reference sets a partition GP's capacity to row_group_size * col_group_size with no upper bound
// (src/range_sensor_gp_3d.cpp:213-214; the nearest reachable grid to BASELINE's "32 x 24" is 33 x 25 partitions of
// n >= 400 samples) and trains / tests each of them with VanillaGaussianProcess::Solve / TestResult
// (src/vanilla_gp.cpp:492-505, 106-150).  A factor of that size does not fit one CTA's shared memory, so here L lives
// where the API materialises it anyway - the batch's L buffer in HBM (col-major, ld = max_n; at these sizes it stays
// in the 126 MB L2 while its CTA works on it) - and one CTA per GP runs a blocked right-looking Cholesky on it:
//   fill      K(x_i, x_j) + noise diagonal straight into the lower triangle of L (strict upper = 0, as matrixL())
//   per 32-column block: diagonal block factorised by one warp in shared memory, panel solve thread-per-row,
//             trailing update in 64 x 64 tiles (two 64 x 32 panel slices staged in shared memory, 4 x 4 outputs / thread)
//   alpha     blocked forward / backward substitution (32 x 32 triangular solves by one warp, the rest by all threads)
// Predict (one CTA per tile of TQ queries of one GP): the n x TQ tile of X = Ktest lives in shared memory; per block
// the 32 x TQ triangular solve is done thread-per-query, the update of the rows below by all threads with the rows of
// L streamed once from L2 (coalesced along the rows); ||v||^2 is accumulated as the blocks are finalised.
// Same BatchParams / modes / info conventions as the one-CTA-per-GP kernels (erl_gp_batched.cuh).
#include "erl_gp_internal.cuh"

namespace erl_gp {
    namespace largegp {

        constexpr int kThreads = 256;
        constexpr int kB = 32;       // block edge of the factorisation / substitution
        constexpr int kLdD = kB + 1; // padded leading dimension of the diagonal block in shared memory
        constexpr int kTile = 64;    // trailing-update tile

        template<typename T>
        __device__ __forceinline__ T
        PointDist2(const T *__restrict__ a, const T *__restrict__ b, const int x_dim) {
            T r2 = 0;
            for (int d = 0; d < x_dim; ++d) {
                const T diff = a[d] - b[d];
                r2 += diff * diff;
            }
            return r2;
        }

        // diagonal block [c0, c0 + nb) of L -> shared memory D (identity padding beyond nb)
        template<typename T>
        __device__ __forceinline__ void
        LoadDiag(const T *gl, const long ld, const int c0, const int nb, T *__restrict__ dblk) {
            for (int e = threadIdx.x; e < kB * kB; e += kThreads) {
                const int r = e & (kB - 1), c = e >> 5;
                T v = r == c ? T(1) : T(0);
                if (r < nb && c < nb && r >= c) { v = gl[(c0 + r) + static_cast<long>(c0 + c) * ld]; }
                dblk[r * kLdD + c] = v;
            }
        }

        template<typename T>
        __global__ void __launch_bounds__(kThreads)
        TrainKernel(const BatchParams<T> p, const int x_dim) {
            extern __shared__ __align__(16) unsigned char smem_raw[];
            T *dblk = reinterpret_cast<T *>(smem_raw);   // [kB][kLdD]
            T *pa = dblk + kB * kLdD;                     // [kTile][kB + 1] panel slice (rows of the tile)
            T *pb = pa + kTile * (kB + 1);                // [kTile][kB + 1] panel slice (cols of the tile)
            T *zs = pb + kTile * (kB + 1);                // [max_n] right-hand side / alpha
            T *red = zs + p.max_n;                        // [kB] column sums of the backward substitution
            __shared__ int s_fail;

            const int g = blockIdx.x;
            const int tid = threadIdx.x;
            const int warp = tid >> 5, lane = tid & 31;
            const int n = p.n_train[g];
            if (n <= p.min_train || n <= 0) {  // `cnt > min_num_samples_per_group` / `cnt > 0` gate of the callers
                if (tid == 0) { p.info[g] = -1; }
                return;
            }
            const long ld = p.max_n;
            T *gl = p.l + static_cast<long>(g) * ld * ld;
            const T *gx = p.x + static_cast<long>(g) * ld * x_dim;
            const T *gy = p.y + static_cast<long>(g) * ld;
            const T *gv = p.var + static_cast<long>(g) * ld;
            if (tid == 0) { s_fail = 0; }

            // ---- Gram matrix into the lower triangle (Covariance::ComputeKtrain: K[i][i] = 1 + var[i]) ----
            for (int r = tid; r < n; r += kThreads) {
                T xr[3] = {0, 0, 0};
                for (int d = 0; d < x_dim; ++d) { xr[d] = gx[static_cast<long>(r) * x_dim + d]; }
                const T diag = T(1) + gv[r];
                for (int c = 0; c < n; ++c) {
                    T v = T(0);
                    if (r > c) {
                        v = p.cov(PointDist2<T>(xr, gx + static_cast<long>(c) * x_dim, x_dim));
                    } else if (r == c) {
                        v = diag;
                    }
                    gl[r + static_cast<long>(c) * ld] = v;
                }
                zs[r] = gy[r];
            }
            __syncthreads();

            // ---- blocked right-looking Cholesky ----
            for (int c0 = 0; c0 < n; c0 += kB) {
                const int nb = n - c0 < kB ? n - c0 : kB;
                LoadDiag<T>(gl, ld, c0, nb, dblk);
                __syncthreads();
                if (warp == 0) {  // lane = row of the diagonal block
                    for (int c = 0; c < nb; ++c) {
                        const T d = dblk[c * kLdD + c];
                        __syncwarp();  // every lane has read the pivot before lane c overwrites it
                        if (!(d > T(0))) {
                            if (lane == 0 && s_fail == 0) { s_fail = c0 + c + 1; }  // LLT failed at this column (Eigen: info() != Success)
                            break;
                        }
                        const T piv = sqrt(d);
                        T lrc = T(0);
                        if (lane >= c) {
                            lrc = lane == c ? piv : dblk[lane * kLdD + c] / piv;
                            dblk[lane * kLdD + c] = lrc;
                        }
                        __syncwarp();
                        if (lane > c) {
                            for (int j = c + 1; j <= lane; ++j) { dblk[lane * kLdD + j] -= lrc * dblk[j * kLdD + c]; }
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
                if (s_fail != 0) { break; }
                for (int e = tid; e < nb * nb; e += kThreads) {
                    const int r = e % nb, c = e / nb;
                    if (r >= c) { gl[(c0 + r) + static_cast<long>(c0 + c) * ld] = dblk[r * kLdD + c]; }
                }
                const int m0 = c0 + nb;  // first row / column of the trailing matrix
                if (m0 >= n) { break; }
                // panel solve: row r of L21 = A21 L11^-T, one thread per row
                for (int r = m0 + tid; r < n; r += kThreads) {
                    T x[kB];
#pragma unroll
                    for (int j = 0; j < kB; ++j) { x[j] = j < nb ? gl[r + static_cast<long>(c0 + j) * ld] : T(0); }
#pragma unroll
                    for (int j = 0; j < kB; ++j) {
                        T s = x[j];
#pragma unroll
                        for (int k = 0; k < j; ++k) { s -= x[k] * dblk[j * kLdD + k]; }
                        x[j] = s / dblk[j * kLdD + j];
                    }
#pragma unroll
                    for (int j = 0; j < kB; ++j) {
                        if (j < nb) { gl[r + static_cast<long>(c0 + j) * ld] = x[j]; }
                    }
                }
                __syncthreads();
                // trailing update A22 -= L21 L21^T (lower triangle), 64 x 64 tiles
                const int m = n - m0;
                const int nt = (m + kTile - 1) / kTile;
                const int ty = tid & 15, tx = tid >> 4;  // 16 x 16 threads, 4 x 4 outputs each; rows along the lanes (coalesced read-modify-write)
                for (int ti = 0; ti < nt; ++ti) {
                    for (int tj = 0; tj <= ti; ++tj) {
                        for (int e = tid; e < kTile * kB; e += kThreads) {
                            const int rr = e & (kTile - 1), k = e >> 6;
                            const int ra = m0 + ti * kTile + rr, rb = m0 + tj * kTile + rr;
                            pa[rr * (kB + 1) + k] = (ra < n && k < nb) ? gl[ra + static_cast<long>(c0 + k) * ld] : T(0);
                            pb[rr * (kB + 1) + k] = (rb < n && k < nb) ? gl[rb + static_cast<long>(c0 + k) * ld] : T(0);
                        }
                        __syncthreads();
                        T acc[4][4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) { acc[i][j] = T(0); }
                        }
#pragma unroll 8
                        for (int k = 0; k < kB; ++k) {
                            T a[4], b[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                a[i] = pa[(ty + 16 * i) * (kB + 1) + k];
                                b[i] = pb[(tx + 16 * i) * (kB + 1) + k];
                            }
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) { acc[i][j] += a[i] * b[j]; }
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int r = m0 + ti * kTile + ty + 16 * i, c = m0 + tj * kTile + tx + 16 * j;
                                if (r < n && c <= r) { gl[r + static_cast<long>(c) * ld] -= acc[i][j]; }
                            }
                        }
                        __syncthreads();
                    }
                }
            }
            __syncthreads();
            if (s_fail != 0) {
                if (tid == 0) { p.info[g] = s_fail; }
                return;
            }

            // ---- alpha = L^-T L^-1 y ----
            for (int c0 = 0; c0 < n; c0 += kB) {  // forward: z = L^-1 y
                const int nb = n - c0 < kB ? n - c0 : kB;
                LoadDiag<T>(gl, ld, c0, nb, dblk);
                __syncthreads();
                if (warp == 0) {
                    T z = lane < nb ? zs[c0 + lane] : T(0);
                    for (int c = 0; c < nb; ++c) {
                        const T zc = __shfl_sync(0xffffffffu, z, c) / dblk[c * kLdD + c];
                        if (lane == c) { z = zc; }
                        if (lane > c) { z -= dblk[lane * kLdD + c] * zc; }
                    }
                    if (lane < nb) { zs[c0 + lane] = z; }
                }
                __syncthreads();
                for (int r = c0 + nb + tid; r < n; r += kThreads) {
                    T s = zs[r];
                    for (int j = 0; j < nb; ++j) { s -= gl[r + static_cast<long>(c0 + j) * ld] * zs[c0 + j]; }
                    zs[r] = s;
                }
                __syncthreads();
            }
            const int last = ((n - 1) / kB) * kB;
            for (int c0 = last; c0 >= 0; c0 -= kB) {  // backward: alpha = L^-T z
                const int nb = n - c0 < kB ? n - c0 : kB;
                LoadDiag<T>(gl, ld, c0, nb, dblk);
                for (int j = warp; j < nb; j += kThreads / 32) {  // red[j] = sum_{r >= c0 + nb} L[r][c0 + j] alpha[r]
                    T s = T(0);
                    for (int r = c0 + nb + lane; r < n; r += 32) { s += gl[r + static_cast<long>(c0 + j) * ld] * zs[r]; }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); }
                    if (lane == 0) { red[j] = s; }
                }
                __syncthreads();
                if (warp == 0) {
                    T a = lane < nb ? zs[c0 + lane] - red[lane] : T(0);
                    for (int c = nb - 1; c >= 0; --c) {
                        const T ac = __shfl_sync(0xffffffffu, a, c) / dblk[c * kLdD + c];
                        if (lane == c) { a = ac; }
                        if (lane < c) { a -= dblk[c * kLdD + lane] * ac; }  // L[c][lane]
                    }
                    if (lane < nb) { zs[c0 + lane] = a; }
                }
                __syncthreads();
            }
            T *ga = p.alpha + static_cast<long>(g) * ld;
            for (int e = tid; e < n; e += kThreads) { ga[e] = zs[e]; }
            if (tid == 0) { p.info[g] = 0; }
        }

        template<typename T>
        static size_t
        TrainSmemBytes(const long max_n) {
            return sizeof(T) * (kB * kLdD + 2 * kTile * (kB + 1) + max_n + kB);
        }

        // queries of GP g in tiles of tq; mark_invalid: fused train + predict semantics (valid = 0 for an untrained / failed GP)
        template<typename T>
        __global__ void __launch_bounds__(kThreads)
        PredictKernel(const BatchParams<T> p, const int x_dim, const int tq, const int ldx, const int mark_invalid) {
            extern __shared__ __align__(16) unsigned char smem_raw[];
            T *xs = reinterpret_cast<T *>(smem_raw);  // [tq][ldx]: X = Ktest tile, row index fastest
            T *dblk = xs + static_cast<long>(tq) * ldx;  // [kB][kLdD]
            T *vb = dblk + kB * kLdD;                     // [kB][tq] solved block
            T *al = vb + kB * tq;                         // [max_n]
            T *mean = al + p.max_n;                       // [tq]
            T *ss = mean + tq;                            // [tq]

            const int g = blockIdx.x;
            const int tid = threadIdx.x;
            const int warp = tid >> 5, lane = tid & 31;
            const long q0 = p.q_offsets[g], q1 = p.q_offsets[g + 1];
            if (q1 <= q0) { return; }
            if (p.info[g] != 0) {
                if (mark_invalid && p.valid != nullptr) {
                    for (long q = q0 + static_cast<long>(blockIdx.y) * kThreads + tid; q < q1; q += static_cast<long>(gridDim.y) * kThreads) {
                        p.valid[p.q_out_index != nullptr ? p.q_out_index[q] : q] = 0;
                    }
                }
                return;  // predict-only mode: outputs of an untrained GP stay untouched
            }
            const int n = p.n_train[g];
            const long ld = p.max_n;
            const T *gl = p.l + static_cast<long>(g) * ld * ld;
            const T *gx = p.x + static_cast<long>(g) * ld * x_dim;
            const T *ga = p.alpha + static_cast<long>(g) * ld;
            for (int e = tid; e < n; e += kThreads) { al[e] = ga[e]; }

            const int tqc = tq >> 2;  // queries per thread in the update (4 query chunks x 64 row lanes)
            for (long qb = q0 + static_cast<long>(blockIdx.y) * tq; qb < q1; qb += static_cast<long>(gridDim.y) * tq) {
                const int nq = static_cast<int>(q1 - qb < tq ? q1 - qb : tq);
                __syncthreads();
                // Ktest tile (Covariance::ComputeKtest, no noise term)
                for (int q = warp; q < tq; q += kThreads / 32) {
                    T xq[3] = {0, 0, 0};
                    if (q < nq) {
                        for (int d = 0; d < x_dim; ++d) { xq[d] = p.q_x[(qb + q) * x_dim + d]; }
                    }
                    T m = T(0);
                    for (int r = lane; r < n; r += 32) {
                        const T k = q < nq ? p.cov(PointDist2<T>(xq, gx + static_cast<long>(r) * x_dim, x_dim)) : T(0);
                        xs[static_cast<long>(q) * ldx + r] = k;
                        m += k * al[r];
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { m += __shfl_xor_sync(0xffffffffu, m, o); }
                    if (lane == 0) {
                        mean[q] = m;
                        ss[q] = T(0);
                    }
                }
                for (int c0 = 0; c0 < n; c0 += kB) {
                    const int nb = n - c0 < kB ? n - c0 : kB;
                    __syncthreads();
                    LoadDiag<T>(gl, ld, c0, nb, dblk);
                    __syncthreads();
                    if (tid < tq) {  // 32 x 32 forward substitution, one thread per query
                        T v[kB];
                        T s2 = T(0);
#pragma unroll
                        for (int j = 0; j < kB; ++j) { v[j] = j < nb ? xs[static_cast<long>(tid) * ldx + c0 + j] : T(0); }
#pragma unroll
                        for (int j = 0; j < kB; ++j) {
                            T s = v[j];
#pragma unroll
                            for (int k = 0; k < j; ++k) { s -= dblk[j * kLdD + k] * v[k]; }
                            v[j] = s / dblk[j * kLdD + j];
                            s2 += v[j] * v[j];
                            vb[j * tq + tid] = v[j];
                        }
                        ss[tid] += s2;
                    }
                    __syncthreads();
                    const int rr = tid & 63, qc = tid >> 6;
                    for (int r = c0 + nb + rr; r < n; r += 64) {
                        T lrow[kB];
#pragma unroll
                        for (int k = 0; k < kB; ++k) { lrow[k] = k < nb ? gl[r + static_cast<long>(c0 + k) * ld] : T(0); }
                        for (int q = qc * tqc; q < (qc + 1) * tqc; ++q) {
                            T acc = T(0);
#pragma unroll
                            for (int k = 0; k < kB; ++k) { acc += lrow[k] * vb[k * tq + q]; }
                            xs[static_cast<long>(q) * ldx + r] -= acc;
                        }
                    }
                }
                __syncthreads();
                if (tid < nq) {
                    const long src = qb + tid;
                    const long dst = p.q_out_index != nullptr ? p.q_out_index[src] : src;
                    if (p.mean != nullptr) {
                        T f = mean[tid];
                        if (p.mapping != ERL_GP_MAPPING_NONE) { f = MappingInv<T>(p.mapping, p.mapping_scale, f); }
                        p.mean[dst] = f;
                    }
                    if (p.variance != nullptr) { p.variance[dst] = T(1) - ss[tid]; }  // literal prior 1.0f, src/vanilla_gp.cpp:121
                    if (p.valid != nullptr) { p.valid[dst] = 1; }
                }
            }
        }

        template<typename T>
        static size_t
        PredictSmemBytes(const long max_n, const int tq, const int ldx) {
            return sizeof(T) * (static_cast<size_t>(tq) * ldx + kB * kLdD + kB * tq + max_n + 2 * tq);
        }

    }  // namespace largegp

    long
    LargeGpMaxN() {
        return 2048;
    }

    template<typename T>
    int
    LaunchLargeGp(Context *ctx, const BatchParams<T> &params, const int x_dim, const int mode, const int tiles_per_gp) {
        using namespace largegp;
        if (params.max_n > LargeGpMaxN()) { return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "batch: max_n=%d exceeds %ld", params.max_n, LargeGpMaxN()); }
        if ((mode & kBatchTrain) != 0) {
            const size_t bytes = TrainSmemBytes<T>(params.max_n);
            auto kernel = TrainKernel<T>;
            ERL_GP_CUDA_OK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
            kernel<<<static_cast<unsigned>(params.num_gps), kThreads, bytes, ctx->stream>>>(params, x_dim);
            ctx->launches += 1;
            ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        }
        if ((mode & kBatchPredict) != 0) {
            const int ldx = (params.max_n + 31) / 32 * 32 + 1;  // odd multiple-of-32 + 1: the per-query columns of the solve start in different banks
            int tq = 64;
            while (tq > 8 && PredictSmemBytes<T>(params.max_n, tq, ldx) > 200 * 1024) { tq >>= 1; }
            const size_t bytes = PredictSmemBytes<T>(params.max_n, tq, ldx);
            if (static_cast<int>(bytes) > ctx->max_smem_optin) {
                return SetError(ctx, ERL_GP_STATUS_UNSUPPORTED, "large-GP predict needs %zu B of shared memory, device allows %d", bytes, ctx->max_smem_optin);
            }
            auto kernel = PredictKernel<T>;
            ERL_GP_CUDA_OK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
            // fused mode: one CTA per GP would serialise its query tiles; spread them like the predict-only callers do
            int tiles = tiles_per_gp < 1 ? 1 : tiles_per_gp;
            if (mode == kBatchTrainPredict) {
                const long want = CeilDiv(2L * ctx->sm_count, params.num_gps);
                tiles = static_cast<int>(want < 1 ? 1 : (want > 64 ? 64 : want));
            }
            const dim3 grid(static_cast<unsigned>(params.num_gps), static_cast<unsigned>(tiles));
            kernel<<<grid, kThreads, bytes, ctx->stream>>>(params, x_dim, tq, ldx, mode == kBatchTrainPredict ? 1 : 0);
            ctx->launches += 1;
            ERL_GP_CUDA_OK(ctx, cudaGetLastError());
        }
        return ERL_GP_STATUS_OK;
    }

    template int LaunchLargeGp<float>(Context *, const BatchParams<float> &, int, int, int);
    template int LaunchLargeGp<double>(Context *, const BatchParams<double> &, int, int, int);

}  // namespace erl_gp
