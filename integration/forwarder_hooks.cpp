// integration/forwarder_hooks.cpp — TEST INFRASTRUCTURE for the stand-alone build of o3d_forwarders.cpp: extern "C"
// entry points that wrap raw device pointers in the ref-shim's stub core::Tensor and call the forwarders THROUGH THE
// REFERENCE'S OWN DECLARATIONS (the headers included below are the reference's), so tests/test_forwarders_gpu.py can
// drive them from ctypes.  Returns 0, or -1 with the message in fwd_last_error().
#define BUILD_CUDA_MODULE
#include <cuda_runtime.h>

#include <cstring>
#include <stdexcept>
#include <string>

#include "open3d/core/Tensor.h"
#include "open3d/core/nns/FixedRadiusIndex.h"
#include "open3d/t/geometry/kernel/Transform.h"
#include "open3d/t/geometry/kernel/VoxelBlockGrid.h"
#include "open3d/t/pipelines/kernel/RegistrationImpl.h"

using open3d::core::Device;
using open3d::core::Tensor;
namespace core = open3d::core;

static thread_local std::string g_err;
static const Device kCuda("CUDA:0");

template <typename F>
static int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

extern "C" {

const char* fwd_last_error() { return g_err.c_str(); }

int fwd_compute_pose_point_to_plane(float* src, float* tgt, float* nrm, int64_t* corr, int64_t n, int robust_method,
                                    double scale, double shape, double pose_out[6], float* residual, int* inlier_count) {
    return guarded([&] {
        namespace reg = open3d::t::pipelines::registration;
        Tensor s(src, {n, 3}, core::Float32, kCuda), t(tgt, {n, 3}, core::Float32, kCuda), nn(nrm, {n, 3}, core::Float32, kCuda),
                c(corr, {n}, core::Int64, kCuda), pose;
        reg::RobustKernel k((reg::RobustKernelMethod)robust_method, scale, shape);
        open3d::t::pipelines::kernel::ComputePosePointToPlaneCUDA(s, t, nn, c, pose, *residual, *inlier_count, core::Float32, kCuda, k);
        if (!pose.GetDevice().IsCPU() || pose.NumElements() != 6) throw std::runtime_error("pose must be a {6} host tensor");
        memcpy(pose_out, pose.GetDataPtr<double>(), 6 * sizeof(double));
    });
}

int fwd_compute_pose_colored_icp(float* src, float* scol, float* tgt, float* nrm, float* tcol, float* tgrad, int64_t* corr,
                                 int64_t n, int64_t m, double lambda_geometric, double pose_out[6], float* residual,
                                 int* inlier_count) {
    return guarded([&] {
        namespace reg = open3d::t::pipelines::registration;
        Tensor s(src, {n, 3}, core::Float32, kCuda), sc(scol, {n, 3}, core::Float32, kCuda), t(tgt, {m, 3}, core::Float32, kCuda),
                nn(nrm, {m, 3}, core::Float32, kCuda), tc(tcol, {m, 3}, core::Float32, kCuda), tg(tgrad, {m, 3}, core::Float32, kCuda),
                c(corr, {n}, core::Int64, kCuda), pose;
        reg::RobustKernel k;
        open3d::t::pipelines::kernel::ComputePoseColoredICPCUDA(s, sc, t, nn, tc, tg, c, pose, *residual, *inlier_count, core::Float32,
                                                               kCuda, k, lambda_geometric);
        memcpy(pose_out, pose.GetDataPtr<double>(), 6 * sizeof(double));
    });
}

int fwd_transform_points(float* T_dev_f32_4x4, float* points, float* normals, int64_t n) {
    return guarded([&] {
        Tensor T(T_dev_f32_4x4, {4, 4}, core::Float32, kCuda), p(points, {n, 3}, core::Float32, kCuda), q(normals, {n, 3}, core::Float32, kCuda);
        open3d::t::geometry::kernel::transform::TransformPointsCUDA(T, p);
        if (normals) open3d::t::geometry::kernel::transform::TransformNormalsCUDA(T, q);
    });
}

// DepthTouchCUDA: returns the number of blocks, copies up to max_blocks keys into block_coords_out (device)
int64_t fwd_depth_touch(void* depth, int depth_is_u16, int rows, int cols, double K[9], double E[16], int resolution,
                        float voxel_size, float sdf_trunc, float depth_scale, float depth_max, int32_t* block_coords_out,
                        int64_t max_blocks) {
    int64_t n = -1;
    const int rc = guarded([&] {
        Tensor d(depth, {rows, cols}, depth_is_u16 ? core::UInt16 : core::Float32, kCuda), k(K, {3, 3}, core::Float64), e(E, {4, 4}, core::Float64), out;
        std::shared_ptr<core::HashMap> hm;
        open3d::t::geometry::kernel::voxel_grid::DepthTouchCUDA(hm, d, k, e, out, resolution, voxel_size, sdf_trunc, depth_scale, depth_max, 4);
        n = out.GetLength();
        if (!out.GetDevice().IsCUDA() || n > max_blocks) throw std::runtime_error("unexpected voxel_block_coords tensor");
        if (cudaMemcpy(block_coords_out, out.GetDataPtr(), (size_t)n * 12, cudaMemcpyDeviceToDevice) != cudaSuccess)
            throw std::runtime_error("copy failed");
    });
    return rc ? -1 : n;
}

// IntegrateCUDA<...>: value_f32 selects the Float32 weight / colour layout, depth_is_u16 the input types
int fwd_integrate(void* depth, void* color, int depth_is_u16, int rows, int cols, int32_t* block_indices, int64_t num_blocks,
                  int32_t* block_keys, int64_t capacity, float* tsdf, void* weight, void* color_buf, int value_f32, double dK[9],
                  double cK[9], double E[16], int resolution, float voxel_size, float sdf_trunc, float depth_scale, float depth_max) {
    return guarded([&] {
        namespace vg = open3d::t::geometry::kernel::voxel_grid;
        const int64_t r3 = (int64_t)resolution * resolution * resolution;
        Tensor d(depth, {rows, cols}, depth_is_u16 ? core::UInt16 : core::Float32, kCuda);
        Tensor c = color ? Tensor(color, {rows, cols, 3}, depth_is_u16 ? core::UInt8 : core::Float32, kCuda) : Tensor();
        Tensor bi(block_indices, {num_blocks}, core::Int32, kCuda), bk(block_keys, {capacity, 3}, core::Int32, kCuda);
        open3d::t::geometry::TensorMap values("tsdf");
        values["tsdf"] = Tensor(tsdf, {capacity, r3}, core::Float32, kCuda);
        values["weight"] = Tensor(weight, {capacity, r3}, value_f32 ? core::Float32 : core::UInt16, kCuda);
        if (color_buf) values["color"] = Tensor(color_buf, {capacity, r3, 3}, value_f32 ? core::Float32 : core::UInt16, kCuda);
        Tensor k1(dK, {3, 3}, core::Float64), k2(cK, {3, 3}, core::Float64), e(E, {4, 4}, core::Float64);
        if (depth_is_u16 && !value_f32) vg::IntegrateCUDA<uint16_t, uint8_t, float, uint16_t, uint16_t>(d, c, bi, bk, values, k1, k2, e, resolution, voxel_size, sdf_trunc, depth_scale, depth_max);
        else if (depth_is_u16) vg::IntegrateCUDA<uint16_t, uint8_t, float, float, float>(d, c, bi, bk, values, k1, k2, e, resolution, voxel_size, sdf_trunc, depth_scale, depth_max);
        else if (!value_f32) vg::IntegrateCUDA<float, float, float, uint16_t, uint16_t>(d, c, bi, bk, values, k1, k2, e, resolution, voxel_size, sdf_trunc, depth_scale, depth_max);
        else vg::IntegrateCUDA<float, float, float, float, float>(d, c, bi, bk, values, k1, k2, e, resolution, voxel_size, sdf_trunc, depth_scale, depth_max);
        if (cudaDeviceSynchronize() != cudaSuccess) throw std::runtime_error("device fault in IntegrateCUDA");
    });
}

// BuildSpatialHashTableCUDA<float> + HybridSearchCUDA<float,int32_t>
int fwd_hash_table_and_hybrid_search(float* points, int64_t m, float* queries, int64_t nq, double radius, int max_knn,
                                     uint32_t hash_table_size, uint32_t* table_index_out, uint32_t* cell_splits_out,
                                     int32_t* idx_out, float* dist_out, int32_t* cnt_out) {
    return guarded([&] {
        namespace nns = open3d::core::nns;
        int64_t prs[2] = {0, m}, qrs[2] = {0, nq};
        uint32_t hts[2] = {0, hash_table_size};
        Tensor p(points, {m, 3}, core::Float32, kCuda), q(queries, {nq, 3}, core::Float32, kCuda);
        Tensor points_row_splits(prs, {2}, core::Int64), queries_row_splits(qrs, {2}, core::Int64), hash_table_splits(hts, {2}, core::UInt32);
        Tensor index(table_index_out, {m}, core::UInt32, kCuda), splits(cell_splits_out, {(int64_t)hash_table_size + 1}, core::UInt32, kCuda);
        nns::BuildSpatialHashTableCUDA<float>(p, radius, points_row_splits, hash_table_splits, index, splits);
        Tensor ni, nc, nd;
        nns::HybridSearchCUDA<float, int32_t>(p, q, radius, max_knn, points_row_splits, queries_row_splits, hash_table_splits, index, splits,
                                              nns::L2, ni, nc, nd);
        if (ni.GetShape(0) != nq || ni.GetShape(1) != max_knn || nc.GetLength() != nq) throw std::runtime_error("bad output shapes");
        cudaMemcpy(idx_out, ni.GetDataPtr(), (size_t)nq * max_knn * 4, cudaMemcpyDeviceToDevice);
        cudaMemcpy(dist_out, nd.GetDataPtr(), (size_t)nq * max_knn * 4, cudaMemcpyDeviceToDevice);
        cudaMemcpy(cnt_out, nc.GetDataPtr(), (size_t)nq * 4, cudaMemcpyDeviceToDevice);
    });
}

}  // extern "C"
