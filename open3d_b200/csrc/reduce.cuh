// reduce.cuh — the 29(+1)-scalar Gauss-Newton reduction shared by the ICP kernels (icp.cu) and RGB-D
// odometry (odometry.cu): per-thread f32 partials -> f64 warp tree -> per-warp shared-memory slots ->
// per-block partials -> last block (atomic ticket) sums in block order; plus the f64 6x6 solve and
// pose -> transformation that the reference runs on the host (kernel/TransformationConverter.cpp).
#pragma once

#include "common.cuh"

namespace o3db {

static constexpr int kThreads = 256;
static constexpr int kNumSums = 30;   // 29 reference slots + sum of dist^2
static constexpr int kSumStride = 32;
#ifndef ICP_FLUSH_EVERY
#define ICP_FLUSH_EVERY 32
#endif
static constexpr int kFlushEvery = ICP_FLUSH_EVERY;   // f32 terms per thread before the f64 tree (error <= kFlushEvery * 2^-24 of sum|term|)

// ------------------------------------------------- 29(+1)-scalar reduction

// Per-thread f32 partials (at most kFlushEvery terms each) -> f64 warp tree ->
// per-warp f64 slots in shared memory.  Deterministic for a fixed launch shape.
__device__ __forceinline__ void flush_acc(float (&acc)[kNumSums], double (*s_warp)[kSumStride]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) {
        const double v = warp_sum((double)acc[k]);
        if (lane == 0) s_warp[w][k] += v;
        acc[k] = 0.f;
    }
}

// Block epilogue: per-warp slots -> block partial -> (last block) grand total in
// block-index order.  Returns true in the last block, with s_final[] filled.
__device__ __forceinline__ bool block_reduce_to_global(double (*s_warp)[kSumStride], double* __restrict__ partials,
                                                       unsigned* ticket, double* s_final) {
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x < kNumSums) {
        double v = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) v += s_warp[w][threadIdx.x];
        partials[(size_t)blockIdx.x * kSumStride + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    if (threadIdx.x < kNumSums) {
        double v = 0;
        for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(&partials[(size_t)b * kSumStride + threadIdx.x]);
        s_final[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) *ticket = 0;
    __syncthreads();
    return true;
}

// --------------------------------------------------------- 6x6 solve (f64)

// TransformationConverter.cpp:189-226: LU with partial pivoting (LAPACK dgesv semantics).
__device__ __host__ inline bool solve6x6(const double* A, double* x) {
    double M[6][7];
    for (int j = 0; j < 6; ++j) {
        const int base = (j * (j + 1)) / 2;
        for (int k = 0; k <= j; ++k) {
            M[j][k] = A[base + k];
            M[k][j] = A[base + k];
        }
        M[j][6] = -A[21 + j];
    }
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        double best = fabs(M[c][c]);
        for (int r = c + 1; r < 6; ++r)
            if (fabs(M[r][c]) > best) {
                best = fabs(M[r][c]);
                piv = r;
            }
        if (!(best > 0.0)) return false;
        if (piv != c)
            for (int k = 0; k < 7; ++k) {
                const double t = M[c][k];
                M[c][k] = M[piv][k];
                M[piv][k] = t;
            }
        for (int r = c + 1; r < 6; ++r) {
            const double f = M[r][c] / M[c][c];
            for (int k = c; k < 7; ++k) M[r][k] -= f * M[c][k];
        }
    }
    for (int r = 5; r >= 0; --r) {
        double s = M[r][6];
        for (int k = r + 1; k < 6; ++k) s -= M[r][k] * x[k];
        x[r] = s / M[r][r];
    }
    return true;
}

// TransformationConverterImpl.h:22-42 + TransformationConverter.cpp:81-104.
__device__ __host__ inline void pose_to_T(const double* p, double* T) {
    for (int i = 0; i < 16; ++i) T[i] = 0.0;
    T[15] = 1.0;
    const double ca = cos(p[0]), sa = sin(p[0]), cb = cos(p[1]), sb = sin(p[1]), cg = cos(p[2]), sg = sin(p[2]);
    T[0] = cg * cb;
    T[1] = -1 * sg * ca + cg * sb * sa;
    T[2] = sg * sa + cg * sb * ca;
    T[4] = sg * cb;
    T[5] = cg * ca + sg * sb * sa;
    T[6] = -1 * cg * sa + sg * sb * ca;
    T[8] = -1 * sb;
    T[9] = cb * sa;
    T[10] = cb * ca;
    T[3] = p[3];
    T[7] = p[4];
    T[11] = p[5];
}

}  // namespace o3db
