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

// Transposed warp reduction (the fused ICP kernels): every lane brings one query's 30 terms, lane l leaves
// with the sum over the warp of term l.  Five exchange steps; in step `half` a lane keeps one half of its
// remaining terms and trades the other half with lane ^ half, so a term costs one shuffle in total
// (16 + 8 + 4 + 2 + 1 = 31 shuffles for 32 slots) instead of five, and the running totals of a warp live in
// ONE register per lane instead of 30: the search keeps the register file.  The association is a fixed
// binary tree over the lanes: deterministic.
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32]) {
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            const float keep = up ? v[k + half] : v[k];
            const float send = up ? v[k] : v[k + half];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

#ifndef O3DB_FENCE_ACQREL
#define O3DB_FENCE_ACQREL 0
#endif
// release of the block partial before the ticket / acquire of everybody's partials after it
__device__ __forceinline__ void reduce_fence() {
#if O3DB_FENCE_ACQREL
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
#else
    __threadfence();
#endif
}

// Block epilogue: per-warp slots -> block partial -> (last block) grand total in
// block-index order.  Returns true in the last block, with s_final[] filled.
template <int THREADS = kThreads>
__device__ __forceinline__ bool block_reduce_to_global(double (*s_warp)[kSumStride], double* __restrict__ partials,
                                                       unsigned* ticket, double* s_final,
                                                       long long* stamps = nullptr) {
    __shared__ bool s_last;
    __syncthreads();
    if (stamps && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(stamps[0]));
    if (threadIdx.x < kNumSums) {
        double v = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) v += s_warp[w][threadIdx.x];
        partials[(size_t)blockIdx.x * kSumStride + threadIdx.x] = v;
    }
    reduce_fence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return false;
    if (stamps && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(stamps[1]));
    reduce_fence();
    // Grand total by the WHOLE block (round 1 had 30 threads walk all per-block partials one dependent
    // L2 round trip at a time: ~130k cycles for 444 blocks, a third of the kernel): warp w sums the blocks
    // b = w, w + 8, ... for all 30 columns (lane = column, 256-byte coalesced rows, 8 loads in flight),
    // then thread k adds the 8 per-warp totals in warp order.  The association is fixed by the launch
    // shape, so the result is still deterministic for a given grid.
    __shared__ double s_part[THREADS / 32][kSumStride];
    {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        constexpr int kW = THREADS / 32, kU = 8;
        double v = 0;
        for (unsigned b = w; b < gridDim.x; b += kU * kW) {   // kU loads in flight, also in the last (partial) batch
            double t[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const unsigned bb = b + u * kW;
                t[u] = bb < gridDim.x ? __ldcg(&partials[(size_t)bb * kSumStride + lane]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) v += t[u];
        }
        s_part[w][lane] = v;
    }
    __syncthreads();
    if (threadIdx.x < kNumSums) {
        double v = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) v += s_part[w][threadIdx.x];
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

// The same factorisation by one warp (the fused loops run it in the serial tail of every iteration, where
// a single thread walking a 6 x 7 array in local memory cost ~5 us): lane c < 7 owns column c of the augmented
// matrix in registers; per elimination step the pivot choice and the five multipliers are computed by the
// owner of the pivot column and broadcast, every lane updates its own column.  Same operations on the same
// operands as solve6x6 up to the back substitution, which runs column-oriented (as LAPACK's dtrsv does) from
// shared memory on lane 0.  Must be called by all 32 lanes of a warp; `Ms` is 42 doubles of shared memory.
// Returns false (on every lane) for a singular system.
__device__ __forceinline__ bool solve6x6_warp(const double* __restrict__ A, double* __restrict__ Ms, double* x) {
    const int lane = threadIdx.x & 31;
    const int c_own = lane < 7 ? lane : 6;
    double col[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        const int hi = r > c_own ? r : c_own, lo = r > c_own ? c_own : r;
        col[r] = c_own == 6 ? -A[21 + r] : A[(hi * (hi + 1)) / 2 + lo];
    }
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        double best = fabs(col[c]);
#pragma unroll
        for (int r = c + 1; r < 6; ++r)
            if (fabs(col[r]) > best) {
                best = fabs(col[r]);
                piv = r;
            }
        piv = __shfl_sync(0xffffffffu, piv, c);
        best = __shfl_sync(0xffffffffu, best, c);
        if (!(best > 0.0)) ok = false;
#pragma unroll
        for (int r = c + 1; r < 6; ++r)
            if (piv == r) {
                const double t = col[c];
                col[c] = col[r];
                col[r] = t;
            }
#pragma unroll
        for (int r = c + 1; r < 6; ++r) {
            const double f = __shfl_sync(0xffffffffu, col[r] / col[c], c);
            col[r] -= f * col[c];
        }
    }
    if (lane < 7) {
#pragma unroll
        for (int r = 0; r < 6; ++r) Ms[r * 7 + lane] = col[r];
    }
    __syncwarp();
    if (lane == 0 && ok) {
        double b[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) b[r] = Ms[r * 7 + 6];
#pragma unroll
        for (int k = 5; k >= 0; --k) {
            const double xk = b[k] / Ms[k * 7 + k];
            x[k] = xk;
#pragma unroll
            for (int r = 0; r < k; ++r) b[r] -= Ms[r * 7 + k] * xk;
        }
    }
    __syncwarp();
    return ok;
}

// TransformationConverterImpl.h:22-42 + TransformationConverter.cpp:81-104.  The trigonometric values come in as
// arguments so that a warp can evaluate the six sin / cos calls on six lanes (odometry.cu) and still run the very
// same expressions as the single-thread path.
__device__ __host__ inline void pose_to_T_trig(const double* p, double ca, double sa, double cb, double sb, double cg,
                                               double sg, double* T) {
    for (int i = 0; i < 16; ++i) T[i] = 0.0;
    T[15] = 1.0;
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

__device__ __host__ inline void pose_to_T(const double* p, double* T) {
    pose_to_T_trig(p, cos(p[0]), sin(p[0]), cos(p[1]), sin(p[1]), cos(p[2]), sin(p[2]), T);
}

}  // namespace o3db
