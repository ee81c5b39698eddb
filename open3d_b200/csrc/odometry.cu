// odometry.cu — RGB-D odometry, PointToPlane method, for sm_100a: what
// slam::Model::TrackFrameToModel runs between RayCast and Integrate in the dense-SLAM loop.
//
// Reference: t/pipelines/odometry/RGBDOdometry.cpp:56-206 (RGBDOdometryMultiScale / ...PointToPlane driver),
// :432-459 (ComputeOdometryResultPointToPlane), t/pipelines/kernel/RGBDOdometryCUDA.cu:37-125 + RGBDOdometryJacobianImpl.h
// (per-pixel Jacobian, Huber terms, 29-float BlockReduce + atomics, host DecodeAndSolve6x6), and the depth-pyramid
// kernels of t/geometry/kernel/ImageImpl.h:86-315 (ClipTransform, PyrDownDepth, CreateVertexMap, CreateNormalMap).
//
// Here the whole multi-scale loop is device resident, like the ICP loop of icp.cu: one kernel per iteration does the
// per-pixel projection + Jacobian, the 29-scalar reduction of reduce.cuh (f32 partials -> f64 tree, deterministic) and,
// in the last block, the f64 6x6 solve, T <- dT * T, the result bookkeeping and the reference's relative convergence
// test; a level that has converged turns its remaining launches into no-ops.  No host synchronisation until the end
// (upstream: one cuda::Synchronize + 29-float D2H + host LU per iteration).
//
// Image::FilterBilateral is NPP upstream (closed source): the kernel here evaluates NPP's documented
// definition of nppiFilterBilateralGaussBorder (see o3db_image_filter_bilateral in the header) — parity unpinned.
//
// All pixel-selecting arithmetic (projection, roundf, residual gate) is evaluated without FMA contraction in the
// reference's source order, as everywhere else in this library.
#include <climits>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "reduce.cuh"
#include "vbg.cuh"

namespace o3db {

static constexpr int kOT = 256;
#ifndef ODO_BLOCKS_PER_SM
#define ODO_BLOCKS_PER_SM 4   // blocks per SM of the iteration kernel; fewer = shorter serial tail in the last block (tunable)
#endif

// ------------------------------------------------------------ image kernels

template <typename src_t>
__global__ void clip_transform_kernel(const src_t* __restrict__ src, int64_t n, float scale, float min_value,
                                      float max_value, float clip_fill, float* __restrict__ dst) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float out = dvd((float)src[i], scale);   // ImageImpl.h:112-116
    out = out <= min_value ? clip_fill : out;
    out = out >= max_value ? clip_fill : out;
    dst[i] = out;
}

// ImageImpl.h:122-198
__global__ void pyr_down_depth_kernel(const float* __restrict__ src, int rows, int cols, float depth_diff,
                                      float invalid_fill, float* __restrict__ dst) {
    const int rows_down = rows / 2, cols_down = cols / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_down * cols_down) return;
    const int y = i / cols_down, x = i % cols_down;
    const int y_src = 2 * y, x_src = 2 * x;
    const float v_center = src[(size_t)y_src * cols + x_src];
    if (v_center == invalid_fill) {
        dst[i] = invalid_fill;
        return;
    }
    const float gweights[3] = {0.375f, 0.25f, 0.0625f};
    const int x_min = max(0, x_src - 2), y_min = max(0, y_src - 2);
    const int x_max = min(cols - 1, x_src + 2), y_max = min(rows - 1, y_src + 2);
    float v_sum = 0.f, w_sum = 0.f;
    for (int yk = y_min; yk <= y_max; ++yk)
        for (int xk = x_min; xk <= x_max; ++xk) {
            const float v = src[(size_t)yk * cols + xk];
            const int dy = abs(yk - y_src), dx = abs(xk - x_src);
            if (v != invalid_fill && fabsf(sub(v, v_center)) < depth_diff) {
                const float w = mul(gweights[dx], gweights[dy]);
                v_sum = add(v_sum, mul(w, v));
                w_sum = add(w_sum, w);
            }
        }
    dst[i] = w_sum == 0 ? invalid_fill : dvd(v_sum, w_sum);
}

__device__ __forceinline__ bool is_invalid(float v, float invalid_fill) {   // ImageImpl.h:227-231
    if (isinf(invalid_fill)) return isinf(v);
    if (isnan(invalid_fill)) return isnan(v);
    return v == invalid_fill;
}

// ImageImpl.h:200-248
__global__ void create_vertex_map_kernel(const float* __restrict__ depth, int rows, int cols, Cam ti, float invalid_fill,
                                         float* __restrict__ vertex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int y = i / cols, x = i % cols;
    const float d = depth[i];
    float vx = invalid_fill, vy = invalid_fill, vz = invalid_fill;
    if (!is_invalid(d, invalid_fill)) unproject(ti, (float)x, (float)y, d, vx, vy, vz);
    vertex[3 * (size_t)i] = vx;
    vertex[3 * (size_t)i + 1] = vy;
    vertex[3 * (size_t)i + 2] = vz;
}

// ImageImpl.h:249-315
__global__ void create_normal_map_kernel(const float* __restrict__ vertex, int rows, int cols, float invalid_fill,
                                         float* __restrict__ normal) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int y = i / cols, x = i % cols;
    float n0 = invalid_fill, n1 = invalid_fill, n2 = invalid_fill;
    if (y < rows - 1 && x < cols - 1) {
        const float* v00 = vertex + 3 * (size_t)i;
        const float* v10 = v00 + 3;
        const float* v01 = v00 + 3 * (size_t)cols;
        const bool bad = (v00[0] == invalid_fill && v00[1] == invalid_fill && v00[2] == invalid_fill) ||
                         (v01[0] == invalid_fill && v01[1] == invalid_fill && v01[2] == invalid_fill) ||
                         (v10[0] == invalid_fill && v10[1] == invalid_fill && v10[2] == invalid_fill);
        if (!bad) {
            const float dx0 = sub(v01[0], v00[0]), dy0 = sub(v01[1], v00[1]), dz0 = sub(v01[2], v00[2]);
            const float dx1 = sub(v10[0], v00[0]), dy1 = sub(v10[1], v00[1]), dz1 = sub(v10[2], v00[2]);
            n0 = sub(mul(dy0, dz1), mul(dz0, dy1));
            n1 = sub(mul(dz0, dx1), mul(dx0, dz1));
            n2 = sub(mul(dx0, dy1), mul(dy0, dx1));
            float norm = __fsqrt_rn(add(add(mul(n0, n0), mul(n1, n1)), mul(n2, n2)));
            norm = fmaxf(norm, 1e-5f);
            n0 = dvd(n0, norm);
            n1 = dvd(n1, norm);
            n2 = dvd(n2, norm);
        }
    }
    normal[3 * (size_t)i] = n0;
    normal[3 * (size_t)i + 1] = n1;
    normal[3 * (size_t)i + 2] = n2;
}

// nppiFilterBilateralGaussBorder_32f_C1R (NPPImage.cpp:319-376) by its documented definition; replicated border.
__global__ void filter_bilateral_kernel(const float* __restrict__ src, int rows, int cols, int radius, float val2,
                                        float pos2, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int y = i / cols, x = i % cols;
    const float vc = src[i];
    float v_sum = 0.f, w_sum = 0.f;
    for (int dy = -radius; dy <= radius; ++dy)
        for (int dx = -radius; dx <= radius; ++dx) {
            const int yy = min(max(y + dy, 0), rows - 1), xx = min(max(x + dx, 0), cols - 1);
            const float v = src[(size_t)yy * cols + xx];
            const float dv = sub(v, vc);
            const float w = mul(expf(dvd(-((float)(dx * dx + dy * dy)), pos2)), expf(dvd(-mul(dv, dv), val2)));
            v_sum = add(v_sum, mul(w, v));
            w_sum = add(w_sum, w);
        }
    dst[i] = dvd(v_sum, w_sum);
}

// ------------------------------------------------------ per-pixel Jacobian

// RGBDOdometryJacobianImpl.h:29-37.  Sign() takes an int (GeometryMacros.h:92): the residual is truncated first.
__device__ __forceinline__ float huber_deriv(float r, float delta) {
    const float abs_r = fabsf(r);
    const int ir = (int)r;
    return abs_r < delta ? r : mul(delta, (float)((ir > 0) ? 1 : ((ir < 0) ? -1 : 0)));
}
__device__ __forceinline__ float huber_loss(float r, float delta) {
    const float abs_r = fabsf(r);
    // `0.5 * r * r` / `delta * abs_r - 0.5 * delta * delta`: double where the literal forces it, as written
    return abs_r < delta ? (float)__dmul_rn(__dmul_rn(0.5, (double)r), (double)r)
                         : (float)__dsub_rn((double)mul(delta, abs_r), __dmul_rn(__dmul_rn(0.5, (double)delta), (double)delta));
}

// RGBDOdometryJacobianImpl.h:106-160
__device__ __forceinline__ bool jacobian_p2plane(int x, int y, float trunc, const float* __restrict__ sv_map,
                                                 const float* __restrict__ tv_map, const float* __restrict__ tn_map,
                                                 int rows, int cols, const Cam& ti, float (&J)[6], float& r) {
    const float* sv = sv_map + 3 * ((size_t)y * cols + x);
    const float s0 = sv[0];
    if (isnan(s0)) return false;
    float p0, p1, p2, u, v;
    rigid(ti, s0, sv[1], sv[2], p0, p1, p2);
    project(ti, p0, p1, p2, u, v);
    u = roundf(u);
    v = roundf(v);
    if (p2 < 0 || !in_boundary(u, v, rows, cols)) return false;
    const int ui = (int)u, vi = (int)v;
    const float* tv = tv_map + 3 * ((size_t)vi * cols + ui);
    const float* tn = tn_map + 3 * ((size_t)vi * cols + ui);
    const float t0 = tv[0], n0 = tn[0];
    if (isnan(t0) || isnan(n0)) return false;
    const float t1 = tv[1], t2 = tv[2], n1 = tn[1], n2 = tn[2];
    r = add(add(mul(sub(p0, t0), n0), mul(sub(p1, t1), n1)), mul(sub(p2, t2), n2));
    if (fabsf(r) > trunc) return false;
    J[0] = add(mul(-p2, n1), mul(p1, n2));
    J[1] = sub(mul(p2, n0), mul(p0, n2));
    J[2] = add(mul(-p1, n0), mul(p0, n1));
    J[3] = n0;
    J[4] = n1;
    J[5] = n2;
    return true;
}

__device__ __forceinline__ void accumulate_odometry(float (&acc)[kNumSums], const float (&J)[6], float r, float delta) {
    const float d_huber = huber_deriv(r, delta), r_huber = huber_loss(r, delta);
    int s = 0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
        for (int k = 0; k <= j; ++k) acc[s++] += J[j] * J[k];   // NOT Huber-weighted, as upstream (:68-70)
        acc[21 + j] += J[j] * d_huber;
    }
    acc[27] += r_huber;
    acc[28] += 1.0f;
}

// ----------------------------------------------------------- fused loop

static constexpr int kMaxLevels = 8;

struct OdoState {
    double T[16];                 // source -> target, updated every iteration
    double sums[kSumStride];
    double res_rmse, res_fitness; // OdometryResult::inlier_rmse_ / fitness_ (RGBDOdometry.cpp:165, 190-191)
    int level_done[kMaxLevels];
    int status;                   // 0 ok, 1 singular 6x6, 2 inlier_count <= 0
    int executed;
    unsigned ticket;
};

struct OdoArgs {
    const float* sv;
    const float* tv;
    const float* tn;
    int rows, cols, level;
    float fx, fy, cx, cy;         // this level's intrinsics (f32, TransformIndexer)
    float trunc, huber_delta;
    double rel_rmse, rel_fitness;
    double* partials;
    OdoState* st;
    double* per_iter;             // optional device log: (inlier_rmse, fitness) per executed iteration
    int standalone;               // 1: ComputeOdometryResultPointToPlane seam (T is not updated, delta -> st->sums[..])
    double* delta_out;            // standalone: 16 doubles (delta transformation)
};

__device__ void odometry_finalize(const OdoArgs& a, const double* s_final) {
    OdoState* st = a.st;
    double s[29], pose[6];
    // the 29 sums reach DecodeAndSolve6x6 as a Float32 tensor (RGBDOdometryCUDA.cu:112-124)
    for (int k = 0; k < 29; ++k) s[k] = (double)(float)s_final[k];
    const int count = (int)s[28];
    if (!solve6x6(s, pose)) {     // TransformationConverter.cpp:215-225
        st->status = 1;
        return;
    }
    if (count <= 0) {             // RGBDOdometry.cpp:449-452
        st->status = 2;
        return;
    }
    double dT[16];
    pose_to_T(pose, dT);
    const double d_rmse = (double)((float)s[27] / (float)count);   // float inlier_residual / int (:455)
    const double d_fit = (double)count / (double)((int64_t)a.rows * a.cols);
    if (a.standalone) {
        for (int i = 0; i < 16; ++i) a.delta_out[i] = dT[i];
        st->res_rmse = d_rmse;
        st->res_fitness = d_fit;
        return;
    }
    double R[16];
    for (int i = 0; i < 4; ++i)   // :175-176 result.transformation_ = delta.transformation_.Matmul(result.transformation_)
        for (int j = 0; j < 4; ++j) {
            double v = 0;
            for (int k = 0; k < 4; ++k) v += dT[i * 4 + k] * st->T[k * 4 + j];
            R[i * 4 + j] = v;
        }
    for (int i = 0; i < 16; ++i) st->T[i] = R[i];
    if (a.per_iter) {
        a.per_iter[2 * st->executed] = d_rmse;
        a.per_iter[2 * st->executed + 1] = d_fit;
    }
    st->executed += 1;
    if (fabs(st->res_fitness - d_fit) / st->res_fitness < a.rel_fitness &&
        fabs(st->res_rmse - d_rmse) / st->res_rmse < a.rel_rmse) {   // :181-189 early exit
        st->level_done[a.level] = 1;
        return;
    }
    st->res_rmse = d_rmse;
    st->res_fitness = d_fit;
}

__global__ void __launch_bounds__(kThreads) odometry_iteration_kernel(OdoArgs a) {
    __shared__ double s_warp[kThreads / 32][kSumStride];
    __shared__ double s_final[kSumStride];
    __shared__ Cam s_cam;
    __shared__ int s_skip;
    if (threadIdx.x == 0) s_skip = (*(volatile int*)&a.st->level_done[a.level]) | (*(volatile int*)&a.st->status);
    if (threadIdx.x < 12) s_cam.e[threadIdx.x / 4][threadIdx.x % 4] = (float)a.st->T[threadIdx.x];   // TransformIndexer: f32
    if (threadIdx.x == 12) {
        s_cam.fx = a.fx;
        s_cam.fy = a.fy;
        s_cam.cx = a.cx;
        s_cam.cy = a.cy;
        s_cam.scale = 1.0f;
    }
    for (int k = threadIdx.x; k < (kThreads / 32) * kSumStride; k += kThreads) (&s_warp[0][0])[k] = 0.0;
    __syncthreads();
    if (s_skip) return;
    const Cam ti = s_cam;
    float acc[kNumSums];
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) acc[k] = 0.f;
    const int n = a.rows * a.cols;
    int since = 0;
    for (int base = blockIdx.x * kThreads; base < n; base += gridDim.x * kThreads) {
        const int i = base + threadIdx.x;
        if (i < n) {
            float J[6], r;
            if (jacobian_p2plane(i % a.cols, i / a.cols, a.trunc, a.sv, a.tv, a.tn, a.rows, a.cols, ti, J, r))
                accumulate_odometry(acc, J, r, a.huber_delta);
        }
        if (++since == kFlushEvery) {
            flush_acc(acc, s_warp);
            since = 0;
        }
    }
    flush_acc(acc, s_warp);
    if (!block_reduce_to_global(s_warp, a.partials, &a.st->ticket, s_final)) return;
    if (threadIdx.x < 29) a.st->sums[threadIdx.x] = s_final[threadIdx.x];
    if (threadIdx.x == 0) odometry_finalize(a, s_final);
}

// ------------------------------------------------------------- host side

static int launch_image_1d(int64_t n) { return (int)ceil_div(n, kOT); }

static int check_image(const char* who, const void* src, const void* dst, int rows, int cols) {
    O3DB_REQUIRE(rows > 0 && cols > 0 && (int64_t)rows * cols < INT_MAX / 4, "%s: invalid shape (%d, %d)", who, rows, cols);
    O3DB_REQUIRE(src != nullptr && dst != nullptr, "%s: null image", who);
    return O3DB_OK;
}

static int clip_transform(const void* src, int dtype, int rows, int cols, float scale, float min_value, float max_value,
                          float clip_fill, float* dst, cudaStream_t st) {
    const int64_t n = (int64_t)rows * cols;
    if (dtype == O3DB_DEPTH_U16)
        clip_transform_kernel<uint16_t><<<launch_image_1d(n), kOT, 0, st>>>((const uint16_t*)src, n, scale, min_value,
                                                                          max_value, clip_fill, dst);
    else
        clip_transform_kernel<float><<<launch_image_1d(n), kOT, 0, st>>>((const float*)src, n, scale, min_value, max_value,
                                                                       clip_fill, dst);
    O3DB_LAUNCH_CHECK();
    return O3DB_OK;
}

static Cam image_cam(const double* K) {
    const double eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    return make_cam(K, eye, 1.0f);
}

struct Level {
    int rows = 0, cols = 0;
    double K[9];
    float *sv = nullptr, *tv = nullptr, *tn = nullptr;
};

}  // namespace o3db

using namespace o3db;

extern "C" {

int o3db_image_clip_transform(const void* src_dev, int depth_dtype, int rows, int cols, float scale, float min_value,
                              float max_value, float clip_fill, float* dst_dev, void* stream) {
    int rc = check_image("ClipTransform", src_dev, dst_dev, rows, cols);
    if (rc) return rc;
    O3DB_REQUIRE(depth_dtype == O3DB_DEPTH_U16 || depth_dtype == O3DB_DEPTH_F32, "ClipTransform: dtype must be UInt16 or Float32");
    O3DB_REQUIRE(!(scale < 0 || min_value < 0 || max_value < 0),
                 "Expected positive scale, min_value, and max_value, but got %g, %g, and %g", scale, min_value, max_value);
    return clip_transform(src_dev, depth_dtype, rows, cols, scale, min_value, max_value, clip_fill, dst_dev, (cudaStream_t)stream);
}

int o3db_image_pyr_down_depth(const float* src_dev, int rows, int cols, float diff_threshold, float invalid_fill,
                              float* dst_dev, void* stream) {
    int rc = check_image("PyrDownDepth", src_dev, dst_dev, rows, cols);
    if (rc) return rc;
    if (rows / 2 == 0 || cols / 2 == 0) return O3DB_OK;
    pyr_down_depth_kernel<<<launch_image_1d((int64_t)(rows / 2) * (cols / 2)), kOT, 0, (cudaStream_t)stream>>>(
            src_dev, rows, cols, diff_threshold, invalid_fill, dst_dev);
    O3DB_LAUNCH_CHECK();
    return O3DB_OK;
}

int o3db_image_create_vertex_map(const float* depth_dev, int rows, int cols, const double K[9], float invalid_fill,
                                 float* vertex_dev, void* stream) {
    int rc = check_image("CreateVertexMap", depth_dev, vertex_dev, rows, cols);
    if (rc) return rc;
    O3DB_REQUIRE(K != nullptr, "CreateVertexMap: null intrinsics");
    create_vertex_map_kernel<<<launch_image_1d((int64_t)rows * cols), kOT, 0, (cudaStream_t)stream>>>(
            depth_dev, rows, cols, image_cam(K), invalid_fill, vertex_dev);
    O3DB_LAUNCH_CHECK();
    return O3DB_OK;
}

int o3db_image_create_normal_map(const float* vertex_dev, int rows, int cols, float invalid_fill, float* normal_dev,
                                 void* stream) {
    int rc = check_image("CreateNormalMap", vertex_dev, normal_dev, rows, cols);
    if (rc) return rc;
    create_normal_map_kernel<<<launch_image_1d((int64_t)rows * cols), kOT, 0, (cudaStream_t)stream>>>(
            vertex_dev, rows, cols, invalid_fill, normal_dev);
    O3DB_LAUNCH_CHECK();
    return O3DB_OK;
}

int o3db_image_filter_bilateral(const float* src_dev, int rows, int cols, int kernel_size, float value_sigma,
                                float dist_sigma, float* dst_dev, void* stream) {
    int rc = check_image("FilterBilateral", src_dev, dst_dev, rows, cols);
    if (rc) return rc;
    O3DB_REQUIRE(kernel_size >= 3, "Kernel size must be >= 3, but got %d.", kernel_size);   // Image.cpp:251-253
    filter_bilateral_kernel<<<launch_image_1d((int64_t)rows * cols), kOT, 0, (cudaStream_t)stream>>>(
            src_dev, rows, cols, kernel_size / 2, 2.0f * (value_sigma * value_sigma), 2.0f * (dist_sigma * dist_sigma),
            dst_dev);
    O3DB_LAUNCH_CHECK();
    return O3DB_OK;
}

}  // extern "C"

namespace o3db {

struct OdoScratch {
    OdoState* st = nullptr;
    OdoState* h_st = nullptr;
    double* partials = nullptr;
    double* per_iter = nullptr;
    double* delta = nullptr;
    int blocks = 1;
};

static void odo_scratch_free(OdoScratch* s, cudaStream_t st) {
    if (s->h_st) cudaStreamSynchronize(st);   // the pinned block may still feed / receive an async copy
    if (s->st) cudaFreeAsync(s->st, st);
    if (s->partials) cudaFreeAsync(s->partials, st);
    if (s->per_iter) cudaFreeAsync(s->per_iter, st);
    if (s->delta) cudaFreeAsync(s->delta, st);
    if (s->h_st) pinned_release(s->h_st);
    *s = OdoScratch{};
}

static int odo_scratch_alloc(OdoScratch* s, int64_t pixels, int log_entries, const double* T, cudaStream_t st) {
    configure_memory_pool();
    s->blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(pixels, kThreads), (int64_t)num_sms() * ODO_BLOCKS_PER_SM));
    static_assert(sizeof(OdoState) <= 4096, "OdoState must fit a pinned block");
    O3DB_CUDA_CHECK(cudaMallocAsync(&s->st, sizeof(OdoState), st));
    O3DB_CUDA_CHECK(cudaMallocAsync(&s->partials, (size_t)s->blocks * kSumStride * sizeof(double), st));
    O3DB_CUDA_CHECK(cudaMallocAsync(&s->per_iter, (size_t)std::max(1, log_entries) * 2 * sizeof(double), st));
    O3DB_CUDA_CHECK(cudaMallocAsync(&s->delta, 16 * sizeof(double), st));
    s->h_st = (OdoState*)pinned_acquire(sizeof(OdoState));
    if (!s->h_st) {
        set_last_error("pinned host allocation failed");
        return O3DB_ERR_CUDA;
    }
    OdoState h{};
    for (int i = 0; i < 16; ++i) h.T[i] = T[i];
    h.res_rmse = 0.0;       // RGBDOdometry.cpp:165 OdometryResult(trans, /*prev rmse*/ 0.0, /*prev fitness*/ 1.0)
    h.res_fitness = 1.0;
    memcpy(s->h_st, &h, sizeof(h));
    O3DB_CUDA_CHECK(cudaMemcpyAsync(s->st, s->h_st, sizeof(OdoState), cudaMemcpyHostToDevice, st));
    O3DB_CUDA_CHECK(cudaMemsetAsync(s->partials, 0, (size_t)s->blocks * kSumStride * sizeof(double), st));
    return O3DB_OK;
}

static int odo_status_to_rc(int status) {
    if (status == 1) {
        set_last_error("Singular 6x6 linear system detected, tracking failed.");
        return O3DB_ERR_SINGULAR;
    }
    if (status == 2) {
        set_last_error("Invalid inlier_count value 0, must be > 0.");   // RGBDOdometry.cpp:449-452
        return O3DB_ERR_NO_INLIERS;
    }
    return O3DB_OK;
}

}  // namespace o3db

extern "C" {

int o3db_compute_odometry_result_point_to_plane(const float* source_vertex_map_dev, const float* target_vertex_map_dev,
                                                const float* target_normal_map_dev, int rows, int cols,
                                                const double K[9], const double init_source_to_target[16],
                                                float depth_outlier_trunc, float depth_huber_delta,
                                                double delta_transformation_host[16], double* inlier_rmse_host,
                                                double* fitness_host, double* sums29_host, void* stream) {
    O3DB_REQUIRE(source_vertex_map_dev && target_vertex_map_dev && target_normal_map_dev,
                 "o3db_compute_odometry_result_point_to_plane: null map");
    O3DB_REQUIRE(rows > 0 && cols > 0 && K && init_source_to_target && delta_transformation_host,
                 "o3db_compute_odometry_result_point_to_plane: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    OdoScratch s;
    int rc = odo_scratch_alloc(&s, (int64_t)rows * cols, 1, init_source_to_target, st);
    if (rc) {
        odo_scratch_free(&s, st);
        return rc;
    }
    OdoArgs a{};
    a.sv = source_vertex_map_dev;
    a.tv = target_vertex_map_dev;
    a.tn = target_normal_map_dev;
    a.rows = rows;
    a.cols = cols;
    a.level = 0;
    a.fx = (float)K[0];
    a.fy = (float)K[4];
    a.cx = (float)K[2];
    a.cy = (float)K[5];
    a.trunc = depth_outlier_trunc;
    a.huber_delta = depth_huber_delta;
    a.partials = s.partials;
    a.st = s.st;
    a.standalone = 1;
    a.delta_out = s.delta;
    odometry_iteration_kernel<<<s.blocks, kThreads, 0, st>>>(a);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.h_st, s.st, sizeof(OdoState), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(delta_transformation_host, s.delta, 16 * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        set_last_error("o3db_compute_odometry_result_point_to_plane: %s", cudaGetErrorString(e));
        odo_scratch_free(&s, st);
        return O3DB_ERR_CUDA;
    }
    rc = odo_status_to_rc(s.h_st->status);
    if (inlier_rmse_host) *inlier_rmse_host = s.h_st->res_rmse;
    if (fitness_host) *fitness_host = s.h_st->res_fitness;
    if (sums29_host) memcpy(sums29_host, s.h_st->sums, 29 * sizeof(double));
    odo_scratch_free(&s, st);
    return rc;
}

int o3db_rgbd_odometry_multi_scale_point_to_plane(const void* source_depth_dev, int source_dtype,
                                                  const void* target_depth_dev, int target_dtype, int rows, int cols,
                                                  const double K[9], const double init_source_to_target[16],
                                                  float depth_scale, float depth_max,
                                                  const o3db_odometry_criteria* criteria, int num_levels,
                                                  float depth_outlier_trunc, float depth_huber_delta,
                                                  o3db_odometry_result* result_host, double* per_iteration_host,
                                                  void* stream) {
    O3DB_REQUIRE(result_host != nullptr, "o3db_rgbd_odometry_multi_scale_point_to_plane: null result");
    O3DB_REQUIRE(source_depth_dev && target_depth_dev && K && init_source_to_target && criteria,
                 "o3db_rgbd_odometry_multi_scale_point_to_plane: null argument");
    O3DB_REQUIRE(num_levels >= 1 && num_levels <= kMaxLevels, "o3db_rgbd_odometry_multi_scale_point_to_plane: 1..%d levels",
                 kMaxLevels);
    O3DB_REQUIRE(rows > 0 && cols > 0 && (rows >> (num_levels - 1)) > 1 && (cols >> (num_levels - 1)) > 1,
                 "o3db_rgbd_odometry_multi_scale_point_to_plane: image %dx%d too small for %d levels", cols, rows, num_levels);
    for (int d : {source_dtype, target_dtype})
        O3DB_REQUIRE(d == O3DB_DEPTH_U16 || d == O3DB_DEPTH_F32, "depth images must be UInt16 or Float32");
    cudaStream_t st = (cudaStream_t)stream;
    configure_memory_pool();
    const float nanf_ = nanf("");
    int total_iters = 0;
    for (int i = 0; i < num_levels; ++i) {
        O3DB_REQUIRE(criteria[i].max_iteration >= 0, "max_iteration must be non-negative");
        total_iters += criteria[i].max_iteration;
    }
    // one allocation for the whole pyramid
    std::vector<Level> lv(num_levels);
    size_t floats = 0;
    {
        int r = rows, c = cols;
        for (int i = 0; i < num_levels; ++i, r /= 2, c /= 2) floats += (size_t)r * c * 9;   // sv, tv, tn
    }
    const size_t full = (size_t)rows * cols;
    float* pool = nullptr;
    O3DB_CUDA_CHECK(cudaMallocAsync(&pool, (floats + full * 4 + full * 3) * sizeof(float), st));
    float* src_d = pool + floats;          // current source / target depth, ping-pong halves, smoothed target, its vertices
    float* tgt_d = src_d + full;
    float* tmp = tgt_d + full;
    float* smooth = tmp + full;
    float* tsm = smooth + full;            // [full * 3]
    OdoScratch s;
    int rc = O3DB_OK;
#define ODO_TRY(expr)                  \
    do {                               \
        if (rc == O3DB_OK) rc = (expr); \
    } while (0)
    // RGBDOdometry.cpp:84-88 ClipTransform(depth_scale, 0, depth_max, NAN)
    ODO_TRY(clip_transform(source_depth_dev, source_dtype, rows, cols, depth_scale, 0.0f, depth_max, nanf_, src_d, st));
    ODO_TRY(clip_transform(target_depth_dev, target_dtype, rows, cols, depth_scale, 0.0f, depth_max, nanf_, tgt_d, st));
    double Kp[9];
    memcpy(Kp, K, sizeof(Kp));
    {
        float* cursor = pool;
        int r = rows, c = cols;
        for (int i = 0; i < num_levels && rc == O3DB_OK; ++i) {   // :132-163
            Level& L = lv[num_levels - 1 - i];
            L.rows = r;
            L.cols = c;
            memcpy(L.K, Kp, sizeof(Kp));
            L.sv = cursor;
            L.tv = cursor + (size_t)r * c * 3;
            L.tn = cursor + (size_t)r * c * 6;
            cursor += (size_t)r * c * 9;
            ODO_TRY(o3db_image_create_vertex_map(src_d, r, c, Kp, nanf_, L.sv, st));
            ODO_TRY(o3db_image_create_vertex_map(tgt_d, r, c, Kp, nanf_, L.tv, st));
            ODO_TRY(o3db_image_filter_bilateral(tgt_d, r, c, 5, 5.0f, 10.0f, smooth, st));
            ODO_TRY(o3db_image_create_vertex_map(smooth, r, c, Kp, nanf_, tsm, st));
            ODO_TRY(o3db_image_create_normal_map(tsm, r, c, nanf_, L.tn, st));
            if (i != num_levels - 1) {
                ODO_TRY(o3db_image_pyr_down_depth(src_d, r, c, depth_outlier_trunc * 2, nanf_, tmp, st));
                std::swap(src_d, tmp);
                ODO_TRY(o3db_image_pyr_down_depth(tgt_d, r, c, depth_outlier_trunc * 2, nanf_, tmp, st));
                std::swap(tgt_d, tmp);
                r /= 2;
                c /= 2;
                for (int k = 0; k < 9; ++k) Kp[k] /= 2;   // :159-160
                Kp[8] = 1;
            }
        }
    }
    ODO_TRY(odo_scratch_alloc(&s, full, total_iters, init_source_to_target, st));
    for (int i = 0; i < num_levels && rc == O3DB_OK; ++i) {
        OdoArgs a{};
        a.sv = lv[i].sv;
        a.tv = lv[i].tv;
        a.tn = lv[i].tn;
        a.rows = lv[i].rows;
        a.cols = lv[i].cols;
        a.level = i;
        a.fx = (float)lv[i].K[0];
        a.fy = (float)lv[i].K[4];
        a.cx = (float)lv[i].K[2];
        a.cy = (float)lv[i].K[5];
        a.trunc = depth_outlier_trunc;
        a.huber_delta = depth_huber_delta;
        a.rel_rmse = criteria[i].relative_rmse;
        a.rel_fitness = criteria[i].relative_fitness;
        a.partials = s.partials;
        a.st = s.st;
        a.per_iter = s.per_iter;
        const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div((int64_t)a.rows * a.cols, kThreads), s.blocks));
        for (int it = 0; it < criteria[i].max_iteration; ++it) {
            odometry_iteration_kernel<<<blocks, kThreads, 0, st>>>(a);
            count_launch();
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_last_error("odometry_iteration_kernel launch failed: %s", cudaGetErrorString(e));
            rc = O3DB_ERR_CUDA;
        }
    }
    if (rc == O3DB_OK) {
        cudaError_t e = cudaMemcpyAsync(s.h_st, s.st, sizeof(OdoState), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess && per_iteration_host && s.h_st->executed > 0) {
            e = cudaMemcpyAsync(per_iteration_host, s.per_iter, (size_t)s.h_st->executed * 2 * sizeof(double),
                                cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
        if (e != cudaSuccess) {
            set_last_error("o3db_rgbd_odometry_multi_scale_point_to_plane: %s", cudaGetErrorString(e));
            rc = O3DB_ERR_CUDA;
        } else {
            memcpy(result_host->transformation, s.h_st->T, sizeof(s.h_st->T));
            result_host->inlier_rmse = s.h_st->res_rmse;
            result_host->fitness = s.h_st->res_fitness;
            result_host->iterations = s.h_st->executed;
            result_host->status = odo_status_to_rc(s.h_st->status);
            rc = result_host->status;
        }
    }
#undef ODO_TRY
    odo_scratch_free(&s, st);
    cudaFreeAsync(pool, st);
    return rc;
}

}  // extern "C"
