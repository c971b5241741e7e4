// FP64 tensor-path helpers (mma.sync.m8n8k4.f64 -> SASS DMMA) shared by the dense GEMM and the fused predict kernel.
// Tile convention: a 128 x 128 C tile, 256 threads; warp w owns rows 64 (w & 1) .. +63 and columns 32 (w >> 1) .. +31;
//   acc[mi][ni][e] = C[64 wm + 8 mi + lane / 4][32 wn + 8 ni + 2 (lane % 4) + e]
// Operand slabs live in shared memory as [k][outer] with row stride kMmaLd doubles.
#pragma once

namespace erl_gp {

    constexpr int kMmaLd = 132;
    constexpr int kMmaBk = 16;

    __device__ __forceinline__ void
    Dmma884(double (&c)[2], const double a, const double b) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
    }

    // acc += A * B over one 16-deep slab: at[kk][row], bt[kk][col]
    __device__ __forceinline__ void
    SlabMma(double (&acc)[8][4][2], const double *__restrict__ at, const double *__restrict__ bt, const int wm, const int wn, const int lane) {
        const int kq = lane & 3;
        const int g = lane >> 2;
#pragma unroll
        for (int k4 = 0; k4 < kMmaBk / 4; ++k4) {
            const double *ap = at + (4 * k4 + kq) * kMmaLd + 64 * wm + g;
            const double *bp = bt + (4 * k4 + kq) * kMmaLd + 32 * wn + g;
            double a[8], b[4];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) { a[mi] = ap[8 * mi]; }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) { b[ni] = bp[8 * ni]; }
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) {
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) { Dmma884(acc[mi][ni], a[mi], b[ni]); }
            }
        }
    }

}  // namespace erl_gp
