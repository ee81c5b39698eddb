// ref_shim_reg.cpp — TEST INFRASTRUCTURE.  Compiles the reference's own CPU reduction kernels, unmodified, from where
// they lie under /root/reference: t/pipelines/kernel/RegistrationCPU.cpp (ComputePosePointToPlaneCPU,
// ComputePoseColoredICPCPU: Jacobian + robust weight + 29-slot accumulation as whole functions) and
// t/pipelines/kernel/RGBDOdometryCPU.cpp (ComputeOdometryResultPointToPlaneCPU), plus the header-inline
// EstimatePointWiseColorGradientKernel of t/geometry/kernel/PointCloudImpl.h.  tbb::parallel_reduce is a stub that
// runs the body once over the whole range (the serial order); DecodeAndSolve6x6 — which upstream implements with
// LAPACK — is replaced by a probe that hands the 29 reduced scalars back to the test.
#include <cmath>
#include <cstdint>
#include <cstring>

using std::abs;
using std::exp;
using std::max;
using std::min;
using std::pow;

#include "open3d/t/pipelines/kernel/RegistrationCPU.cpp"   // from -I /root/reference/cpp
#include "open3d/t/pipelines/kernel/RGBDOdometryCPU.cpp"
#include "open3d/t/geometry/kernel/PointCloudImpl.h"        // EstimatePointWiseColorGradientKernel (:1066-1165)

namespace {
thread_local double g_sums[29];
}

namespace open3d {
namespace t {
namespace pipelines {
namespace kernel {
// probe: TransformationConverter.cpp:189-226 receives exactly this tensor
void DecodeAndSolve6x6(const core::Tensor& A, core::Tensor&, float& inlier_residual, int& inlier_count) {
    for (int i = 0; i < 29; ++i)
        g_sums[i] = A.GetDtype() == core::Float32 ? (double)A.GetDataPtr<float>()[i] : A.GetDataPtr<double>()[i];
    inlier_residual = (float)g_sums[27];
    inlier_count = (int)g_sums[28];
}
}  // namespace kernel
}  // namespace pipelines
}  // namespace t
}  // namespace open3d

namespace o3c = open3d::core;
namespace o3k = open3d::t::pipelines::kernel;
namespace o3r = open3d::t::pipelines::registration;

extern "C" {

// ComputePosePointToPlaneCPU (RegistrationCPU.cpp:93-122), Float32 clouds
void ref_pose_p2plane_sums_f32(const float* src, const float* tgt, const float* nrm, const int64_t* corr, int64_t n,
                               int64_t m, int method, double scale, double shape, double sums29[29]) {
    o3c::Tensor s((void*)src, {n, 3}, o3c::Float32), t((void*)tgt, {m, 3}, o3c::Float32), nn((void*)nrm, {m, 3}, o3c::Float32);
    o3c::Tensor c((void*)corr, {n}, o3c::Int64), pose;
    float residual;
    int count;
    o3r::RobustKernel k(static_cast<o3r::RobustKernelMethod>(method), scale, shape);
    o3k::ComputePosePointToPlaneCPU(s, t, nn, c, pose, residual, count, o3c::Float32, o3c::Device(), k);
    memcpy(sums29, g_sums, sizeof(g_sums));
}

// ComputePoseColoredICPCPU (RegistrationCPU.cpp:124-218)
void ref_pose_colored_sums_f32(const float* src, const float* src_c, const float* tgt, const float* nrm,
                               const float* tgt_c, const float* tgt_g, const int64_t* corr, int64_t n, int64_t m,
                               int method, double scale, double shape, double lambda_geometric, double sums29[29]) {
    o3c::Tensor s((void*)src, {n, 3}, o3c::Float32), sc((void*)src_c, {n, 3}, o3c::Float32);
    o3c::Tensor t((void*)tgt, {m, 3}, o3c::Float32), nn((void*)nrm, {m, 3}, o3c::Float32);
    o3c::Tensor tc((void*)tgt_c, {m, 3}, o3c::Float32), tg((void*)tgt_g, {m, 3}, o3c::Float32);
    o3c::Tensor c((void*)corr, {n}, o3c::Int64), pose;
    float residual;
    int count;
    o3r::RobustKernel k(static_cast<o3r::RobustKernelMethod>(method), scale, shape);
    o3k::ComputePoseColoredICPCPU(s, sc, t, nn, tc, tg, c, pose, residual, count, o3c::Float32, o3c::Device(), k,
                                  lambda_geometric);
    memcpy(sums29, g_sums, sizeof(g_sums));
}

// odometry::ComputeOdometryResultPointToPlaneCPU (RGBDOdometryCPU.cpp:286-362)
void ref_odometry_p2plane_sums(const float* source_vertex, const float* target_vertex, const float* target_normal,
                               int rows, int cols, const double K[9], const double T[16], float depth_outlier_trunc,
                               float depth_huber_delta, double sums29[29]) {
    o3c::Tensor sv((void*)source_vertex, {rows, cols, 3}, o3c::Float32), tv((void*)target_vertex, {rows, cols, 3}, o3c::Float32);
    o3c::Tensor tn((void*)target_normal, {rows, cols, 3}, o3c::Float32);
    o3c::Tensor Kt((void*)K, {3, 3}, o3c::Float64), Tt((void*)T, {4, 4}, o3c::Float64), delta;
    float residual;
    int count;
    o3k::odometry::ComputeOdometryResultPointToPlaneCPU(sv, tv, tn, Kt, Tt, delta, residual, count, depth_outlier_trunc,
                                                        depth_huber_delta);
    memcpy(sums29, g_sums, sizeof(g_sums));
}

// EstimatePointWiseColorGradientKernel<float> (t/geometry/kernel/PointCloudImpl.h:1066-1165) for point i with its
// neighbour list (indices[0] = the point itself, as the hybrid search returns it).  Ends in the reference's own
// solve_svd3x3<float>.
void ref_color_gradient_point_f32(const float* points, const float* normals, const float* colors, int64_t i,
                                  const int32_t* indices, int32_t count, float* gradients) {
    const int32_t off = (int32_t)(3 * i);
    open3d::t::geometry::kernel::pointcloud::EstimatePointWiseColorGradientKernel<float>(points, normals, colors, off,
                                                                                          indices, count, gradients);
}

// ComputeInformationMatrixCPU (RegistrationCPU.cpp:655-735), Float32 clouds: 6x6 Float64 GTG over the matched target points
void ref_information_matrix_f32(const float* tgt, int64_t m, const int64_t* corr, int64_t n, double info36[36]) {
    o3c::Tensor t((void*)tgt, {m, 3}, o3c::Float32), c((void*)corr, {n}, o3c::Int64), info(info36, {6, 6}, o3c::Float64);
    open3d::t::pipelines::kernel::ComputeInformationMatrixCPU(t, c, info, o3c::Float32, o3c::Device());
}

}  // extern "C"
